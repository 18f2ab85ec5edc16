#!/usr/bin/env python
"""BER/BLER sweep of BASELINE.json configs[1] on one B200, with the CPU oracle decoded beside it on identical inputs.

  python tools/ber_sweep.py --out profiles/r01_ber_sweep.json [--max-batches 50] [--oracle-cw 256]

Per Eb/N0 point (0..5 dB): BinarySource -> LDPC5GEncoder(4224, 8448) -> QPSK Mapper -> AWGN -> Demapper(app) ->
LDPC5GDecoder(20 it, boxplus-phi), batch 4096, through sim_ber (stops at 1000 block errors or --max-batches). The
first --oracle-cw codewords of the first batch are also decoded by the oracle (libm math = stand-in for the reference's
TensorFlow CPU kernels, and kernel math) from the SAME channel LLRs; their bit errors are listed next to the GPU's on
those codewords (north-star bar: BER within 1e-6 absolute on identical inputs).
Test infrastructure: uses oracle/ as the checker only.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "ber_sweep.json"))
    ap.add_argument("--max-batches", type=int, default=50)
    ap.add_argument("--oracle-cw", type=int, default=256)
    ap.add_argument("--cn-update", default="boxplus-phi")
    args = ap.parse_args()
    import numpy as np
    import torch
    import __graft_entry__ as ge
    ge.build()
    from sionna_b200.phy import config
    from sionna_b200.phy.mapping import BinarySource, Mapper, Demapper
    from sionna_b200.phy.channel import AWGN
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    from sionna_b200.phy.utils import ebnodb2no, sim_ber
    from oracle import ldpc as O

    k, n, batch = 4224, 8448, 4096
    config.seed = 1234
    src, enc = BinarySource(), LDPC5GEncoder(k, n)
    mapper, demapper, awgn = Mapper("qam", 2), Demapper("app", "qam", 2), AWGN()
    dec = LDPC5GDecoder(enc, num_iter=20, cn_update=args.cn_update, hard_out=True)
    keep = {}

    def mc_fun(batch_size, ebno_db):
        no = ebnodb2no(ebno_db, 2, k / n)
        b = src([batch_size, k])
        llr = demapper(awgn(mapper(enc(b)), no), no)
        key = float(ebno_db)
        if key not in keep:                                  # first batch of this SNR point: kept for the oracle
            keep[key] = (b[:args.oracle_cw].clone(), llr[:args.oracle_cw].clone())
        return b, dec(llr)

    ebnos = [0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 4.0, 5.0]
    t0 = time.time()
    ber, bler = sim_ber(mc_fun, ebnos, batch, args.max_batches, num_target_block_errors=1000, early_stop=False,
                        verbose=True)
    torch.cuda.synchronize()
    gpu_s = time.time() - t0
    enc_r = O.LDPC5GEncoderRef(k, n)
    points = []
    for i, e in enumerate(ebnos):
        b, llr = keep[float(e)]
        b_np, llr_np = b.cpu().numpy(), llr.cpu().numpy()
        u_gpu = dec(llr).cpu().numpy()
        row = {"ebno_db": e, "ber": float(ber[i]), "bler": float(bler[i]), "oracle_codewords": int(b_np.shape[0]),
               "bit_errors_same_inputs": {"gpu": int((u_gpu != b_np).sum())}}
        for mode, name in ((0, "oracle_libm"), (1, "oracle_kernel_math")):
            ref = O.LDPC5GDecoderRef(enc_r, cn_update=args.cn_update, num_iter=20)
            u = ref(llr_np, math_mode=mode, num_threads=os.cpu_count())
            row["bit_errors_same_inputs"][name] = int((u != b_np).sum())
            row["bits_differing_from_gpu_" + name] = int((u != u_gpu).sum())
        nb = b_np.size
        row["abs_ber_diff_gpu_vs_libm_same_inputs"] = abs(row["bit_errors_same_inputs"]["gpu"] -
                                                          row["bit_errors_same_inputs"]["oracle_libm"]) / nb
        points.append(row)
        print(row, flush=True)
    out = {"config": "configs[1]: LDPC5G k=4224 n=8448, QPSK, AWGN, 20 it %s, batch 4096" % args.cn_update,
           "device": torch.cuda.get_device_name(0), "gpu_seconds_total": gpu_s, "max_batches": args.max_batches,
           "points": points}
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
