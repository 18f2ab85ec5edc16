"""The C-ABI library loads and exports every symbol include/sionna_b200.h declares (no compute calls: runs without
a GPU); host-side argument validation returns error codes instead of crashing."""
import ctypes as C
import os
import re
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sionna_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound(sb_lib):
    from sionna_b200 import _lib
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(sb_lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in sionna_b200/_lib.py"
    for n in _lib.SIGNATURES:
        assert n in names, f"{n} bound but not declared in the header"


def test_error_codes_without_gpu(sb_lib):
    h = C.c_void_p()
    cn = np.array([0, 0, 5], np.int32)          # CN index 5 out of range
    vn = np.array([0, 1, 2], np.int32)
    rc = sb_lib.sb_ldpc_graph_create(C.byref(h), 2, 3, 3, cn.ctypes.data, vn.ctypes.data, None, 0, None, 0, None, 0, 0)
    assert rc == -1 and b"out of range" in sb_lib.sb_last_error()
    cn = np.array([0, 0, 1], np.int32)
    vn = np.array([0, 0, 2], np.int32)          # duplicate edge (0,0)
    rc = sb_lib.sb_ldpc_graph_create(C.byref(h), 2, 3, 3, cn.ctypes.data, vn.ctypes.data, None, 0, None, 0, None, 0, 0)
    assert rc == -1 and b"duplicate" in sb_lib.sb_last_error()
    assert sb_lib.sb_version() >= 100


def test_blocks_fail_loudly_without_cuda():
    """No CPU fallback: calling a block without a CUDA device raises."""
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    from sionna_b200.phy.mapping import Mapper
    dec = LDPC5GDecoder(LDPC5GEncoder(50, 100))
    with pytest.raises(RuntimeError):
        dec(np.zeros((2, 100), np.float32))
    with pytest.raises(RuntimeError):
        Mapper("qam", 2)(np.zeros((2, 4), np.float32))


def test_empty_batches_are_accepted_without_a_device(sb_lib):
    """Every batch-taking entry point returns SB_OK for an empty batch before touching pointers or the device."""
    assert sb_lib.sb_awgn(None, None, 1, None, 0, 0, 0, None) == 0
    assert sb_lib.sb_binary_source(None, 0, 0, 0, None) == 0
    assert sb_lib.sb_ldpc5g_encode(None, None, 0, None, None) == 0
    assert sb_lib.sb_ldpc_decode(None, None, 0, 20, 0, 0, 0.0, 20.0, 1, None, None, None, None, 0, None) == 0
    assert sb_lib.sb_count_errors(None, None, 0, 8, None, None) == 0
    assert sb_lib.sb_pusch_precode(None, None, None, 0, 1, 1, 2, 12, None) == 0
