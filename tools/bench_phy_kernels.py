#!/usr/bin/env python
"""Roofline table of the non-LDPC hot-path kernels (SURVEY.md 8 row a / BASELINE.json configs[0], [2], [3]) on one B200.

  python tools/bench_phy_kernels.py [--out profiles/r01_phy_kernels.json] [--only NAME]

Each entry times ONE public block call (CUDA events, 20 iterations after 3 warm-ups, inputs larger than L2 or rotated
between two buffers) and reports achieved GB/s = algorithmic bytes (compulsory reads + writes of the call, stated per
entry) / time against MEASURED_PEAKS.json's HBM bandwidth. `--only NAME` runs a single entry once (for ncu captures).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    import numpy as np
    import torch
    import __graft_entry__ as ge
    ge.build()
    from sionna_b200.phy import config
    from sionna_b200.phy.mapping import Mapper, Demapper, BinarySource
    from sionna_b200.phy.ofdm import (ResourceGrid, ResourceGridMapper, OFDMModulator, OFDMDemodulator,
                                      LSChannelEstimator, LMMSEEqualizer, LinearDetector)
    from sionna_b200.phy.mimo import StreamManagement
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder
    from sionna_b200.phy.utils import complex_normal
    config.seed = 7
    dev = config.device
    peak = 6576.1
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f).get("hbm_gbs", peak))
    except OSError:
        pass
    entries = []

    def add(name, fn, alg_bytes, note):
        entries.append((name, fn, alg_bytes, note))

    # ---- configs[0]/[2]: mapping -----------------------------------------------------------------------------------
    nsym = 2048 * 14 * 76 * 8
    bits6 = BinarySource()([nsym * 6])
    m64, d64 = Mapper("qam", 6), Demapper("app", "qam", 6)
    x64 = m64(bits6)
    no = torch.full((1,), 0.05, device=dev)
    add("mapper_64qam", lambda: m64(bits6), nsym * (6 * 4 + 8), "bits fp32 in, complex64 out; %d symbols" % nsym)
    add("demapper_app_64qam", lambda: d64(x64, 0.05), nsym * (8 + 6 * 4), "complex64 in, 6 fp32 LLRs out per symbol")
    dml = Demapper("maxlog", "qam", 6)
    add("demapper_maxlog_64qam", lambda: dml(x64, 0.05), nsym * (8 + 6 * 4), "same, max-log")
    # ---- configs[2]: OFDM 14 x 76 (+6 CP), and a 4096-point 5G carrier ----------------------------------------------
    for fft, cp, batch in ((76, 6, 2048 * 16), (4096, 288, 256)):
        xg = complex_normal([batch, 1, 1, 14, fft])
        mod, dem = OFDMModulator(cp), OFDMDemodulator(fft, 0, cp)
        xt = mod(xg)
        add(f"ofdm_modulate_{fft}", lambda mod=mod, xg=xg: mod(xg), batch * 14 * 8 * (fft + fft + cp),
            f"batch {batch} x 14 symbols, complex64 grid in, time samples (+CP) out")
        add(f"ofdm_demodulate_{fft}", lambda dem=dem, xt=xt: dem(xt), batch * 14 * 8 * (fft + cp + fft),
            f"batch {batch} x 14 symbols, time samples in, grid out")
    # ---- configs[3]: 4 streams x 16 rx antennas, 14 x 76 grid, batch 1024 ------------------------------------------
    rg = ResourceGrid(14, 76, 15e3, num_tx=1, num_streams_per_tx=4, cyclic_prefix_length=6, num_guard_carriers=(5, 6),
                      dc_null=True, pilot_pattern="kronecker", pilot_ofdm_symbol_indices=[2, 11])
    sm = StreamManagement(np.array([[1]]), 4)
    b = 1024
    f_eff = rg.num_effective_subcarriers
    y = complex_normal([b, 1, 16, 14, 76])
    h = complex_normal([b, 1, 16, 1, 4, 14, f_eff])
    ev = torch.zeros((), device=dev)
    eq = LMMSEEqualizer(rg, sm)
    n_re = b * 14 * f_eff
    add("ofdm_lmmse_4x16", lambda: eq(y, h, ev, 0.05), n_re * (16 * 8 + 64 * 8) + b * 4 * rg.num_data_symbols * 12,
        "per RE: y 16 + H 64 complex64 in; x_hat complex64 + no_eff fp32 out per data symbol and stream")
    det = LinearDetector("lmmse", "bit", "maxlog", rg, sm, "qam", 4)
    add("linear_detector_4x16_16qam", lambda: det(y, h, ev, 0.05),
        n_re * (16 * 8 + 64 * 8) + b * 4 * rg.num_data_symbols * 16, "fused equaliser + max-log demapper, LLRs out")
    est = LSChannelEstimator(rg, "nn")
    add("ls_estimator_nn_4x16", lambda: est(y, 0.05), b * 16 * 14 * 76 * 8 + b * 16 * 4 * 14 * f_eff * 12,
        "y in; h_hat complex64 + err_var fp32 out over the whole grid")
    estl = LSChannelEstimator(rg, "lin")
    add("ls_estimator_lin_4x16", lambda: estl(y, 0.05), b * 16 * 14 * 76 * 8 + b * 16 * 4 * 14 * f_eff * 12, "same, linear")
    # ---- encoder ------------------------------------------------------------------------------------------------------
    enc = LDPC5GEncoder(4224, 8448)
    u = BinarySource()([4096 * 4, 4224])
    add("ldpc5g_encode_4224_8448", lambda: enc(u), 4096 * 4 * 4 * (4224 + 8448), "batch 16384, fp32 bits in/out")

    results = []
    for name, fn, alg, note in entries:
        if args.only and name != args.only:
            continue
        if args.only:
            fn()
            torch.cuda.synchronize()
            continue
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        gbs = alg / ms * 1e-6
        results.append({"kernel_path": name, "ms": ms, "alg_bytes": alg, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak,
                        "note": note})
        print(f"{name:32s} {ms:9.4f} ms {alg / 1e6:10.1f} MB {gbs:9.1f} GB/s  {100 * gbs / peak:5.1f} % of {peak:.0f}", flush=True)
    if args.out and results:
        with open(os.path.join(ROOT, args.out) if not os.path.isabs(args.out) else args.out, "w") as f:
            json.dump({"device": torch.cuda.get_device_name(0), "hbm_peak_gbs": peak,
                       "timing": "CUDA events, mean of 20 calls after 3 warm-ups, includes launch + host glue of the block",
                       "results": results}, f, indent=1)


if __name__ == "__main__":
    main()
