"""GPU parity: sb_ldpc_decode (through the LDPCBPDecoder / LDPC5GDecoder host classes and the C-ABI) against the
CPU oracle on identical seeded inputs.

Bar: BIT-EXACT (np.array_equal on soft outputs and decoder state) against the oracle in kernel-math mode
(math_mode=1: same sb_math.h functions, order="kernel": same summation order). Against the oracle in libm mode
(glibc expf/logf, reference list orders) hard decisions must agree on >= 99.9 % of bits and soft outputs of
single iterations within rtol 1e-4 -- the north-star tolerance -- see test_vs_libm_oracle.
"""
import numpy as np
import pytest
import torch

from oracle import ldpc as O

pytestmark = pytest.mark.gpu
RULES = ["boxplus-phi", "boxplus", "minsum", "offset-minsum"]


@pytest.fixture(params=["qc", "generic"])
def kernel_path(request, monkeypatch):
    """5G graphs run either the index-free QC kernel (default) or, with SB_LDPC_DISABLE_QC=1, the generic one."""
    monkeypatch.setenv("SB_LDPC_DISABLE_QC", "0" if request.param == "qc" else "1")
    return request.param


def _noisy_llr(c, ebno_db, rate, rng):
    """BPSK over AWGN: logits log p(1)/p(0) for codeword bits c."""
    no = 1.0 / (10 ** (ebno_db / 10) * rate)
    x = 2.0 * c - 1.0                     # bit 1 -> +1 so that logit = 4 y / no > 0 for bit 1
    y = x + rng.normal(size=c.shape) * np.sqrt(no / 2)
    return (4 * y / no).astype(np.float32)


def _example_pcm(i):
    import os
    p = os.path.join(os.path.dirname(O.__file__), "..", "sionna_b200", "phy", "fec", "ldpc", "codes", "example_pcms.npz")
    with np.load(p) as d:
        return d[f"pcm{i}"].astype(np.float64)


@pytest.mark.parametrize("rule", RULES + ["identity"])
@pytest.mark.parametrize("pcm_id", [0, 1, 2, 3, 4])
def test_generic_pcm_bit_exact(cuda_device, rule, pcm_id):
    from sionna_b200.phy.fec.ldpc import LDPCBPDecoder
    pcm = _example_pcm(pcm_id)
    rng = np.random.default_rng(100 + pcm_id)
    n = pcm.shape[1]
    llr = (rng.normal(size=(37, n)) * 3 + 1.0).astype(np.float32)
    llr[0] = 0.0                                    # all-erasure row
    llr[1, ::3] = 0.0
    llr[2] = 50.0 * np.sign(llr[2])                 # beyond llr_max
    for hard in (True, False):
        dec = LDPCBPDecoder(pcm, cn_update=rule, hard_out=hard, num_iter=7, return_state=True)
        x, st = dec(torch.from_numpy(llr).to(cuda_device))
        xr, str_ = O.bp_decode(pcm, llr, num_iter=7, cn_update=rule, hard_out=hard, return_state=True,
                               math_mode=1, order="kernel")
        assert np.array_equal(x.cpu().numpy(), xr), f"{rule} pcm{pcm_id} hard={hard}"
        assert np.array_equal(st.cpu().numpy(), str_)
    if rule != "identity":
        assert np.all(x.cpu().numpy()[0] == 0.0)     # all-erasure in -> exactly 0 out (test_ldpc_decoding.py:277-290)


@pytest.mark.parametrize("k,n", [(5032, 9216), (5032, 9600)])
def test_qc_phi_with_reduced_log_table_copies(cuda_device, k, n):
    """Graphs that leave < 16 KB of shared memory next to the messages run the boxplus-phi kernel with 8 copies
    ((5032, 9216): the code block of a 16-PRB PUSCH slot) or a single copy ((5032, 9600)) of the log table: same bits."""
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    rng = np.random.default_rng(n)
    enc_r = O.LDPC5GEncoderRef(k, n)
    c = enc_r(rng.integers(0, 2, (12, k)))
    llr = _noisy_llr(c, 1.2, k / n, rng)
    dec = LDPC5GDecoder(LDPC5GEncoder(k, n), hard_out=False, return_infobits=False, num_iter=8)
    assert dec._graph.is_qc()
    x = dec(torch.from_numpy(llr).to(cuda_device)).cpu().numpy()
    ref = O.LDPC5GDecoderRef(enc_r, hard_out=False, return_infobits=False, num_iter=8)
    assert np.array_equal(x, ref(llr, math_mode=1, order="kernel"))


@pytest.mark.parametrize("rule", RULES)
@pytest.mark.parametrize("k,n", [(64, 128), (100, 200), (562, 871), (1024, 2048), (1500, 2000), (4224, 8448)])
def test_5g_bit_exact(cuda_device, kernel_path, rule, k, n):
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    rng = np.random.default_rng(k + n)
    enc_r = O.LDPC5GEncoderRef(k, n)
    bs = 24 if n > 4000 else 40
    u = rng.integers(0, 2, (bs, k))
    c = enc_r(u)
    llr = _noisy_llr(c, 1.5, k / n, rng)
    enc = LDPC5GEncoder(k, n)
    for hard, info in ((True, True), (False, False)):
        dec = LDPC5GDecoder(enc, cn_update=rule, hard_out=hard, return_infobits=info, num_iter=10, return_state=True)
        assert dec.on_chip and dec._graph.is_qc() == (kernel_path == "qc")
        x, st = dec(torch.from_numpy(llr).to(cuda_device))
        ref = O.LDPC5GDecoderRef(enc_r, cn_update=rule, hard_out=hard, return_infobits=info, num_iter=10,
                                 return_state=True)
        xr, str_ = ref(llr, math_mode=1, order="kernel")
        assert np.array_equal(x.cpu().numpy(), xr)
        assert np.array_equal(st.cpu().numpy(), str_)


def test_5g_interleaver_and_no_pruning(cuda_device, kernel_path):
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    rng = np.random.default_rng(7)
    k, n, m = 300, 720, 6
    enc_r = O.LDPC5GEncoderRef(k, n, num_bits_per_symbol=m)
    c = enc_r(rng.integers(0, 2, (16, k)))
    llr = _noisy_llr(c, 3.0, k / n, rng)
    enc = LDPC5GEncoder(k, n, num_bits_per_symbol=m)
    for prune in (True, False):
        for info in (True, False):
            dec = LDPC5GDecoder(enc, hard_out=False, return_infobits=info, prune_pcm=prune, num_iter=5)
            x = dec(torch.from_numpy(llr).to(cuda_device)).cpu().numpy()
            ref = O.LDPC5GDecoderRef(enc_r, hard_out=False, return_infobits=info, prune_pcm=prune, num_iter=5)
            assert np.array_equal(x, ref(llr, math_mode=1, order="kernel"))


@pytest.mark.parametrize("rule", ["boxplus-phi", "minsum"])
def test_layered_schedule_bit_exact(cuda_device, rule):
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    rng = np.random.default_rng(11)
    k, n = 200, 400
    enc_r = O.LDPC5GEncoderRef(k, n)
    c = enc_r(rng.integers(0, 2, (20, k)))
    llr = _noisy_llr(c, 2.0, k / n, rng)
    dec = LDPC5GDecoder(LDPC5GEncoder(k, n), cn_update=rule, cn_schedule="layered", hard_out=False, num_iter=4,
                        return_state=True)
    x, st = dec(torch.from_numpy(llr).to(cuda_device))
    ref = O.LDPC5GDecoderRef(enc_r, cn_update=rule, cn_schedule="layered", hard_out=False, num_iter=4,
                             return_state=True)
    xr, sr = ref(llr, math_mode=1, order="kernel")
    assert np.array_equal(x.cpu().numpy(), xr)
    assert np.array_equal(st.cpu().numpy(), sr)


def test_large_graph_global_workspace_path(cuda_device):
    """k=8448, n=23000 does not fit in shared memory: same kernel, messages in the L2-resident workspace."""
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    rng = np.random.default_rng(5)
    k, n = 8448, 23000
    enc_r = O.LDPC5GEncoderRef(k, n)
    c = enc_r(rng.integers(0, 2, (6, k)))
    llr = _noisy_llr(c, 1.0, k / n, rng)
    dec = LDPC5GDecoder(LDPC5GEncoder(k, n), hard_out=False, return_infobits=False, num_iter=3)
    assert not dec.on_chip
    x = dec(torch.from_numpy(llr).to(cuda_device)).cpu().numpy()
    ref = O.LDPC5GDecoderRef(enc_r, hard_out=False, return_infobits=False, num_iter=3)
    assert np.array_equal(x, ref(llr, math_mode=1, order="kernel"))


def test_state_handover_and_multidim(cuda_device, kernel_path):
    """1 x N iterations == N x 1 iteration with state hand-over (test_ldpc_decoding.py:875-911); [..., n] batches."""
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    rng = np.random.default_rng(3)
    k, n = 120, 300
    enc = LDPC5GEncoder(k, n)
    llr = torch.from_numpy((rng.normal(size=(2, 3, 5, n)) * 2).astype(np.float32)).to(cuda_device)
    dec = LDPC5GDecoder(enc, hard_out=False, num_iter=5, return_state=True)
    x5, s5 = dec(llr)
    dec1 = LDPC5GDecoder(enc, hard_out=False, num_iter=1, return_state=True)
    x, s = dec1(llr)
    for _ in range(4):
        x, s = dec1(llr, msg_v2c=s)
    assert x5.shape == (2, 3, 5, k)
    assert torch.equal(x5, x) and torch.equal(s5, s)
    # num_iter given at call time, bound on outputs (test_ldpc_decoding.py:978-997)
    x2, s2 = dec1(llr, num_iter=5)
    assert torch.equal(x2, x5)
    assert float(s5.abs().max()) <= 20.0 and float(x5.abs().max()) <= 20.0


def test_vs_libm_oracle(cuda_device, kernel_path):
    """Against the oracle's glibc-libm / reference-order mode (the stand-in for TF's own libm): after ONE iteration
    soft outputs agree to rtol 1e-4 (north-star LLR tolerance) for every rule; after 20 iterations at 2 dB the
    decoded bits agree."""
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    rng = np.random.default_rng(21)
    k, n = 1024, 2048
    enc_r = O.LDPC5GEncoderRef(k, n)
    u = rng.integers(0, 2, (64, k))
    llr = _noisy_llr(enc_r(u), 2.5, k / n, rng)
    enc = LDPC5GEncoder(k, n)
    for rule in RULES:
        dec = LDPC5GDecoder(enc, cn_update=rule, hard_out=False, return_infobits=False, num_iter=1)
        x = dec(torch.from_numpy(llr).to(cuda_device)).cpu().numpy()
        xr = O.LDPC5GDecoderRef(enc_r, cn_update=rule, hard_out=False, return_infobits=False, num_iter=1)(llr)
        np.testing.assert_allclose(x, xr, rtol=1e-4, atol=1e-5)
        dec = LDPC5GDecoder(enc, cn_update=rule, num_iter=20)
        ub = dec(torch.from_numpy(llr).to(cuda_device)).cpu().numpy()
        ur = O.LDPC5GDecoderRef(enc_r, cn_update=rule, num_iter=20)(llr)
        assert np.mean(ub != ur) < 1e-3
        assert np.mean(ub != u) < 1e-2


def test_qc_zero_iterations_and_large_llr_max(cuda_device):
    """num_iter = 0 returns the clipped channel logits; llr_max so large that min-sum must take the generic kernel's
    literal 1e5-sentinel path still matches the oracle."""
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    rng = np.random.default_rng(9)
    k, n = 400, 800
    enc, enc_r = LDPC5GEncoder(k, n), O.LDPC5GEncoderRef(k, n)
    llr = (rng.normal(size=(12, n)) * 8).astype(np.float32)
    x = LDPC5GDecoder(enc, hard_out=False, return_infobits=False, num_iter=0)(torch.from_numpy(llr).to(cuda_device))
    assert np.array_equal(x.cpu().numpy(), np.clip(llr, -20, 20))
    big = (rng.normal(size=(12, n)) * 4e4).astype(np.float32)
    for rule in ("minsum", "offset-minsum"):
        dec = LDPC5GDecoder(enc, cn_update=rule, hard_out=False, num_iter=3, llr_max=90000.0)
        ref = O.LDPC5GDecoderRef(enc_r, cn_update=rule, hard_out=False, num_iter=3, llr_max=90000.0)
        assert np.array_equal(dec(torch.from_numpy(big).to(cuda_device)).cpu().numpy(),
                              ref(big, math_mode=1, order="kernel"))


def test_device_phi_scalar_and_packed_equal_oracle(cuda_device):
    """phi() on the device (scalar sb_math.h path and packed FFMA2 path of sb_math2.cuh) == CPU oracle, bit for bit."""
    import ctypes
    from sionna_b200 import _lib
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(0, 20, 200000), 10 ** rng.uniform(-9, 1.3, 200000), [0, 8.5e-8, 16.635532, 40, 1, 10]])
    x = x.astype(np.float32)[: (len(x) // 2) * 2]
    xd = torch.from_numpy(x).to(cuda_device)
    o1, o2 = torch.empty_like(xd), torch.empty_like(xd)
    _lib.check(_lib.lib().sb_debug_phi(_lib.ptr(xd), _lib.ptr(o1), _lib.ptr(o2), len(x), _lib.current_stream()), "sb_debug_phi")
    ref = np.array([O.phi(v, 1) for v in x[:5000]], np.float32)
    assert np.array_equal(o1.cpu().numpy()[:5000], ref)
    assert torch.equal(o1, o2)


# ---- reference summation order: bit-exact against an oracle that shares NO code with the product -------------------
@pytest.mark.parametrize("rule", ["minsum", "offset-minsum"])
@pytest.mark.parametrize("k,n,ebno", [(1024, 2048, 0.5), (4224, 8448, 1.0)])
def test_reference_order_minsum_bit_exact_vs_pure_libm_oracle(cuda_device, rule, k, n, ebno):
    """sum_order="reference": every node reduction runs in the reference's own list order (the np.argsort results of
    decoding.py:286, 329). (offset-)min-sum uses no transcendental function, so the CUDA result must equal the oracle
    built WITHOUT the product's sb_math.h (oracle/_build/libsbo_libm.so, math_mode 0, order "reference") bit for bit -
    soft outputs, hard decisions and the msg_v2c state - also BELOW the waterfall where most codewords do not
    converge and a different summation order changes ~10 % of the hard decisions."""
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    rng = np.random.default_rng(k + 17)
    enc_r = O.LDPC5GEncoderRef(k, n)
    u = rng.integers(0, 2, (48, k))
    llr = _noisy_llr(enc_r(u), ebno, k / n, rng)
    enc = LDPC5GEncoder(k, n)
    dec = LDPC5GDecoder(enc, cn_update=rule, hard_out=False, return_infobits=False, num_iter=20, return_state=True,
                        sum_order="reference")
    assert not dec._graph.is_qc()
    x, st = dec(torch.from_numpy(llr).to(cuda_device))
    ref = O.LDPC5GDecoderRef(enc_r, cn_update=rule, hard_out=False, return_infobits=False, num_iter=20,
                             return_state=True)
    xr, sr = ref(llr, math_mode=0, order="reference", pure=True)
    assert np.array_equal(x.cpu().numpy(), xr)
    assert np.array_equal(st.cpu().numpy(), sr)
    # the default (ascending) order differs on these non-converged inputs: the orders are not interchangeable
    xa = LDPC5GDecoder(enc, cn_update=rule, hard_out=False, return_infobits=False, num_iter=20)(
        torch.from_numpy(llr).to(cuda_device)).cpu().numpy()
    assert not np.array_equal(xa, xr)
    hb = LDPC5GDecoder(enc, cn_update=rule, num_iter=20, sum_order="reference")(torch.from_numpy(llr).to(cuda_device))
    hr = O.LDPC5GDecoderRef(enc_r, cn_update=rule, num_iter=20)(llr, math_mode=0, order="reference", pure=True)
    assert np.array_equal(hb.cpu().numpy(), hr)


@pytest.mark.parametrize("pcm_id", [2, 3, 4])
def test_reference_order_generic_pcm(cuda_device, pcm_id):
    """Same for generic parity-check matrices (BCH(127,106), (3,6)-LDPC, 802.11n), all rules: min-sum vs the pure-libm
    oracle; the transcendental rules vs the kernel-math oracle in reference order."""
    from sionna_b200.phy.fec.ldpc import LDPCBPDecoder
    pcm = _example_pcm(pcm_id)
    rng = np.random.default_rng(200 + pcm_id)
    llr = (rng.normal(size=(33, pcm.shape[1])) * 2 + 0.5).astype(np.float32)
    for rule in RULES:
        dec = LDPCBPDecoder(pcm, cn_update=rule, hard_out=False, num_iter=12, return_state=True, sum_order="reference")
        x, st = dec(torch.from_numpy(llr).to(cuda_device))
        pure = rule in ("minsum", "offset-minsum")
        xr, sr = O.bp_decode(pcm, llr, num_iter=12, cn_update=rule, hard_out=False, return_state=True,
                             math_mode=0 if pure else 1, order="reference", pure=pure)
        assert np.array_equal(x.cpu().numpy(), xr), rule
        assert np.array_equal(st.cpu().numpy(), sr), rule


def test_reference_order_layered(cuda_device):
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    rng = np.random.default_rng(12)
    k, n = 200, 400
    enc_r = O.LDPC5GEncoderRef(k, n)
    llr = _noisy_llr(enc_r(rng.integers(0, 2, (20, k))), 1.0, k / n, rng)
    dec = LDPC5GDecoder(LDPC5GEncoder(k, n), cn_update="minsum", cn_schedule="layered", hard_out=False, num_iter=6,
                        sum_order="reference")
    x = dec(torch.from_numpy(llr).to(cuda_device)).cpu().numpy()
    ref = O.LDPC5GDecoderRef(enc_r, cn_update="minsum", cn_schedule="layered", hard_out=False, num_iter=6)
    assert np.array_equal(x, ref(llr, math_mode=0, order="reference", pure=True))


@pytest.mark.parametrize("rule", ["boxplus-phi", "minsum"])
def test_full_batch_4096_bit_exact_vs_oracle(cuda_device, rule):
    """configs[1] at its FULL size (batch 4096, n = 8448, 20 iterations) with converged and non-converged codewords in
    the same launch (half the batch far below, half above the decoding threshold): soft outputs of the QC kernel == oracle (kernel math, kernel order) for every one
    of the 4096 x 8448 values; for min-sum additionally sum_order="reference" == the oracle build without sb_math.h."""
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    from bench import host_cores
    k, n, bs = 4224, 8448, 4096
    rng = np.random.default_rng(4096)
    enc_r = O.LDPC5GEncoderRef(k, n)
    u = rng.integers(0, 2, (bs, k))
    c = enc_r(u)
    llr = np.concatenate([_noisy_llr(c[:bs // 2], 0.5, k / n, rng), _noisy_llr(c[bs // 2:], 4.0, k / n, rng)])
    enc = LDPC5GEncoder(k, n)
    threads = host_cores()[0]
    x = LDPC5GDecoder(enc, cn_update=rule, hard_out=False, return_infobits=False, num_iter=20)(
        torch.from_numpy(llr).to(cuda_device)).cpu().numpy()
    ref = O.LDPC5GDecoderRef(enc_r, cn_update=rule, hard_out=False, return_infobits=False, num_iter=20)
    xr = ref(llr, math_mode=1, order="kernel", num_threads=threads)
    assert np.array_equal(x, xr)
    blk = ((x > 0) != (c > 0)).any(axis=1).mean()           # transmitted codeword positions
    assert 0.0 < blk < 1.0                                  # both converged and non-converged codewords are present
    if rule == "minsum":
        x2 = LDPC5GDecoder(enc, cn_update=rule, hard_out=False, return_infobits=False, num_iter=20, sum_order="reference")(
            torch.from_numpy(llr).to(cuda_device)).cpu().numpy()
        assert np.array_equal(x2, ref(llr, math_mode=0, order="reference", num_threads=threads, pure=True))


@pytest.mark.parametrize("rule", RULES)
def test_early_termination_equals_fixed_iteration_decodes(cuda_device, rule):
    """early_stop=True: a codeword stops once its hard decisions satisfy every check. Its output must equal the
    plain decoder run with num_iter = the reported iteration count, bit for bit; non-converging codewords run num_iter
    iterations; at high SNR the average iteration count is a fraction of num_iter while BLER is unchanged."""
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    rng = np.random.default_rng(404)
    k, n, it = 1024, 2048, 20
    enc_r = O.LDPC5GEncoderRef(k, n)
    u = rng.integers(0, 2, (96, k))
    c = enc_r(u)
    llr = np.concatenate([_noisy_llr(c[:32], 0.0, k / n, rng), _noisy_llr(c[32:64], 2.2, k / n, rng), _noisy_llr(c[64:], 5.0, k / n, rng)])
    x = torch.from_numpy(llr).to(cuda_device)
    enc = LDPC5GEncoder(k, n)
    dec = LDPC5GDecoder(enc, cn_update=rule, hard_out=False, return_infobits=False, num_iter=it, early_stop=True)
    y = dec(x).cpu().numpy()
    iters = dec.num_iter_run.cpu().numpy()
    assert iters.shape == (96,) and iters.min() >= 2 and iters.max() <= it
    assert np.all(iters[:32] == it)                                # far below the threshold: never converges
    assert iters[64:].mean() < 6 and iters[32:64].mean() < it      # high SNR: a few iterations
    for v in np.unique(iters):
        sel = np.nonzero(iters == v)[0]
        ref = LDPC5GDecoder(enc, cn_update=rule, hard_out=False, return_infobits=False, num_iter=int(v))(x[sel]).cpu().numpy()
        assert np.array_equal(y[sel], ref), (rule, v)
    full = LDPC5GDecoder(enc, cn_update=rule, hard_out=True, num_iter=it)(x).cpu().numpy()
    early = LDPC5GDecoder(enc, cn_update=rule, hard_out=True, num_iter=it, early_stop=True)(x).cpu().numpy()
    assert np.array_equal((full[32:] != u[32:]).any(1), (early[32:] != u[32:]).any(1))     # same block errors where it matters
    with pytest.raises(Exception):
        LDPC5GDecoder(enc, early_stop=True, sum_order="reference")(x)                      # generic kernel: unsupported
