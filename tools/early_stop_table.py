#!/usr/bin/env python
"""Opt-in early termination: iterations run, BLER and decode time against the fixed 20-iteration
decode on identical inputs, configs[1] code (k = 4224, n = 8448), batch 4096, QPSK / AWGN.

  python tools/early_stop_table.py --out profiles/r02_early_stop.json
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
    ap.add_argument("--ebno-dbs", default="1.0,1.5,2.0,2.5,3.0,4.0")
    ap.add_argument("--batch", type=int, default=4096)
    args = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    ge.build()
    from sionna_b200.phy import config
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    from sionna_b200.phy.mapping import Mapper, Demapper, BinarySource
    from sionna_b200.phy.channel import AWGN
    from sionna_b200.phy.utils import ebnodb2no
    config.seed = 7
    k, n = 4224, 8448
    enc = LDPC5GEncoder(k, n)
    dec = LDPC5GDecoder(enc, num_iter=20)
    dec_es = LDPC5GDecoder(enc, num_iter=20, early_stop=True)
    src, mp, dm, ch = BinarySource(), Mapper("qam", 2), Demapper("app", "qam", 2), AWGN()

    def timed(fn, reps=5):
        for _ in range(3):                                     # the first calls of a process carry one-time initialisation
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return out, e0.elapsed_time(e1) / reps

    rows = []
    for db in [float(v) for v in args.ebno_dbs.split(",")]:
        no = float(ebnodb2no(db, 2, k / n))
        b = src([args.batch, k])
        llr = dm(ch(mp(enc(b)), no), no)
        u20, ms20 = timed(lambda: dec(llr))
        ues, mses = timed(lambda: dec_es(llr))
        it = dec_es.num_iter_run.float()
        bl20 = int((u20 != b).any(-1).sum())
        bles = int((ues != b).any(-1).sum())
        rows.append({"ebno_db": db, "block_errors_fixed20": bl20, "block_errors_early_stop": bles, "blocks": args.batch,
                     "mean_iterations": float(it.mean()), "max_iterations": int(it.max()),
                     "frac_stopped_before_20": float((it < 20).float().mean()),
                     "ms_fixed20": ms20, "ms_early_stop": mses})
        print(json.dumps(rows[-1]))
    if args.out:
        json.dump({"config": "LDPC5G k=4224 n=8448, boxplus-phi, 20 iterations max, batch %d, QPSK / AWGN" % args.batch,
                   "device": torch.cuda.get_device_name(0), "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
