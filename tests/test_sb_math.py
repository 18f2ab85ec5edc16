"""sb_math.h accuracy bounds (the deterministic exp / log / tanh / atanh and the table-driven log inside phi): runs
tools/check_math.c over every 997th float (all exponents, ~4.3 M samples per function; the exhaustive run takes ~100 s and
is done by hand) and asserts the bounds the header states, plus the exact values the phi clipping constants rely on."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sb_math_bounds_sampled(tmp_path):
    exe = str(tmp_path / "check_math")
    subprocess.run(["gcc", "-O2", "-std=gnu11", "-ffp-contract=off", "-mfma", "-fopenmp", "-o", exe,
                    os.path.join(ROOT, "tools", "check_math.c"), "-lm"], check=True)
    out = subprocess.run([exe, "997"], check=True, capture_output=True, text=True).stdout
    val = {k: float(v) for k, v in re.findall(r"(\w+) max ulp err ([0-9.]+)", out)}
    # exhaustive maxima: 0.9876 / 0.8623 / 1.5090 / 2.7681 ulp and 0.9769 for the table log
    assert val["expf"] <= 0.99 and val["logf"] <= 0.87 and val["tanhf"] <= 1.51 and val["atanhf"] <= 2.77
    tab = float(re.search(r"logf_tab max err ([0-9.]+)", out).group(1))
    assert tab <= 0.98
    assert "phi<0 count 0" in out and re.search(r"phi non-monotone steps (\d+)", out)
    assert "exp(8.5e-8)=0x1.000002p+0" in out                       # 1 + 2^-23: the lower clip survives the exp
    assert re.search(r"phi\(16\.635532\)=0\b", out) and re.search(r"phi\(40\)=0\b", out)     # exact zero at the upper clip
    assert "phi(2*phi(0))=0" in out
    lg = re.search(r"log\(2\^24\)=(\S+) log\(2\^24-1\)=(\S+)", out)
    assert lg.group(1) == lg.group(2)
