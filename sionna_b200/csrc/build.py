"""Build libsionna_b200.so in-tree with nvcc for sm_100a (no torch extension machinery needed:
the boundary is a plain C-ABI loaded with ctypes).

    python -m sionna_b200.csrc.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "..", "libsionna_b200.so")
SOURCES = ["common.cu", "ldpc_bp.cu", "ldpc_bp_qc.cu", "ldpc_bp_flat.cu", "ldpc_enc.cu", "phy_kernels.cu", "ofdm_mimo.cu", "channel.cu", "frontend.cu"]
HEADERS = ["sb_common.h", "sb_math.h", "sb_math2.cuh", "sb_logtab.h", "rng.cuh", "ldpc_graph.h", "ldpc_rules.cuh", "lmmse_diag.cuh", "demap_qam.cuh",
           os.path.join("..", "..", "include", "sionna_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # never contract a*b+c implicitly: parity with the CPU oracle is bit-exact
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Xcompiler", "-ffp-contract=off",
    "--shared", "-Xptxas", "-v",
]


def _stale(lib, deps):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source into ``sionna_b200/libsionna_b200.so``; returns the library path."""
    srcs = [os.path.join(HERE, s) for s in SOURCES]
    deps = srcs + [os.path.join(HERE, h) for h in HEADERS] + [os.path.abspath(__file__)]
    lib = os.path.abspath(LIB)
    if not force and not _stale(lib, deps):
        return lib
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tmp = f"{lib}.tmp.{os.getpid()}"                        # atomic publish: concurrent builders (one per rank) cannot
    cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp] + srcs          # expose a half-written library
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed building libsionna_b200.so")
    os.replace(tmp, lib)
    with open(os.path.join(HERE, "..", "build_ptxas.log"), "w") as f:
        f.write(res.stdout + res.stderr)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
