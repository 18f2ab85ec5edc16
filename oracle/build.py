"""Compile the C part of the oracle into oracle/_build/libsbo.so (gcc, OpenMP). Test infrastructure only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "libsbo.so")
OUT_LIBM = os.path.join(HERE, "_build", "libsbo_libm.so")   # ldpc_bp_ref.c alone, without the product's sb_math.h
SRCS = ["ldpc_bp_ref.c", "mapping_ref.c"]


def build(force=False):
    srcs = [os.path.join(HERE, s) for s in SRCS]
    deps = srcs + [os.path.join(HERE, "..", "sionna_b200", "csrc", h) for h in ("sb_math.h", "sb_logtab.h")]
    if not force and os.path.exists(OUT) and os.path.exists(OUT_LIBM) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # -ffp-contract=off: fp32 operations stay separately rounded like the reference's TF ops (and like the
    # CUDA kernels, built with -fmad=false); -mfma only makes the explicit fmaf() calls of sb_math.h fast.
    tmp = f"{OUT}.tmp.{os.getpid()}"                        # atomic publish (several ranks may build at once)
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-mfma",
           "-fno-fast-math", "-o", tmp] + srcs + ["-lm"]
    subprocess.run(cmd, check=True)
    os.replace(tmp, OUT)
    tmp = f"{OUT_LIBM}.tmp.{os.getpid()}"
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
           "-DSBO_PURE_LIBM", "-o", tmp, srcs[0], "-lm"]
    subprocess.run(cmd, check=True)
    os.replace(tmp, OUT_LIBM)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
