#!/usr/bin/env python
"""BER/BLER sweep of BASELINE.json configs[1] on one B200 for all four check-node rules, with the CPU oracle decoded
beside the GPU on IDENTICAL channel LLRs, plus the measured soft-output deviation after 20 iterations.

  python tools/ber_sweep.py --out profiles/r02_ber_sweep.json [--rules boxplus-phi,boxplus,minsum,offset-minsum]
                            [--max-batches 30] [--oracle-max-cw 4096]

Per rule and Eb/N0 point: BinarySource -> LDPC5GEncoder(4224, 8448) -> QPSK Mapper -> AWGN -> Demapper(app) ->
LDPC5GDecoder(20 it), batch 4096.
  * `gpu`: the product's default path (QC kernel, ascending summation order) through sim_ber until 200 block errors or
    --max-batches batches.
  * paired comparison on the first M codewords (M grows in chunks of 512 until the ORACLE has seen >= 100 block errors,
    at most --oracle-max-cw): oracle = libm math + the reference's own list orders (for the (offset-)min-sum rules the
    oracle build that shares no code with the product, oracle/_build/libsbo_libm.so). Reported: bit / block errors of
    both sides on those codewords, the paired BER difference with a 95 % confidence interval (per-codeword differences),
    the number of differing hard decisions, and the same for the product's sum_order="reference" path (generic kernel),
    which must be bit-identical to the oracle for min-sum / offset-min-sum.
  * `soft_deviation`: at 1.0 and 2.5 dB, 64 codewords, soft outputs after 20 iterations: share of entries whose relative
    difference to the libm oracle exceeds 1e-4, and the maximum, for the default path and for sum_order="reference".
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
    ap.add_argument("--rules", default="boxplus-phi,boxplus,minsum,offset-minsum")
    ap.add_argument("--ebno-dbs", default="0,0.5,1.0,1.25,1.5,2.0,3.0")
    ap.add_argument("--max-batches", type=int, default=30)
    ap.add_argument("--oracle-max-cw", type=int, default=4096)
    ap.add_argument("--paired-up-to-db", type=float, default=1.5)
    args = ap.parse_args()
    import numpy as np
    import torch
    import __graft_entry__ as ge
    ge.build()
    from bench import host_cores
    from sionna_b200.phy import config
    from sionna_b200.phy.mapping import BinarySource, Mapper, Demapper
    from sionna_b200.phy.channel import AWGN
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    from sionna_b200.phy.utils import ebnodb2no, sim_ber
    from oracle import ldpc as O

    k, n, batch = 4224, 8448, 4096
    cores, core_info = host_cores()
    src, enc = BinarySource(), LDPC5GEncoder(k, n)
    mapper, demapper, awgn = Mapper("qam", 2), Demapper("app", "qam", 2), AWGN()
    enc_r = O.LDPC5GEncoderRef(k, n)
    ebnos = [float(v) for v in args.ebno_dbs.split(",")]
    out = {"config": "configs[1]: LDPC5G k=4224 n=8448, QPSK, AWGN, 20 iterations, batch 4096",
           "device": torch.cuda.get_device_name(0), "host_threads": cores, "host_detail": core_info, "rules": {}}

    def draw(ebno_db, count):
        no = ebnodb2no(ebno_db, 2, k / n)
        b = src([count, k])
        return b, demapper(awgn(mapper(enc(b)), no), no)

    for rule in args.rules.split(","):
        pure = rule in ("minsum", "offset-minsum", "min")
        config.seed = 1234
        dec = LDPC5GDecoder(enc, num_iter=20, cn_update=rule, hard_out=True)
        dec_ref = LDPC5GDecoder(enc, num_iter=20, cn_update=rule, hard_out=True, sum_order="reference")
        orc = O.LDPC5GDecoderRef(enc_r, cn_update=rule, num_iter=20)

        def mc_fun(batch_size, ebno_db):
            b, llr = draw(ebno_db, batch_size)
            return b, dec(llr)

        t0 = time.time()
        ber, bler = sim_ber(mc_fun, ebnos, batch, args.max_batches, num_target_block_errors=200, early_stop=False,
                            verbose=False)
        torch.cuda.synchronize()
        rule_out = {"gpu_seconds": time.time() - t0, "points": []}
        for i, e in enumerate(ebnos):
            row = {"ebno_db": e, "gpu": {"ber": float(ber[i]), "bler": float(bler[i])}}
            if e <= args.paired_up_to_db:
                config.seed = 5000 + int(round(100 * e))
                tot = {"cw": 0, "gpu_bit": 0, "gpu_blk": 0, "orc_bit": 0, "orc_blk": 0, "ref_bit": 0, "ref_blk": 0,
                       "diff_gpu_orc": 0, "diff_ref_orc": 0}
                d_cw, t_orc = [], 0.0
                while tot["cw"] < args.oracle_max_cw and tot["orc_blk"] < 100:
                    b, llr = draw(e, 512)
                    b_np, llr_np = b.cpu().numpy(), llr.cpu().numpy()
                    u_gpu = dec(llr).cpu().numpy()
                    u_ref = dec_ref(llr).cpu().numpy()
                    t1 = time.time()
                    u_orc = orc(llr_np, math_mode=0, order="reference", num_threads=cores, pure=pure)
                    t_orc += time.time() - t1
                    eg, eo, er = (u_gpu != b_np), (u_orc != b_np), (u_ref != b_np)
                    tot["cw"] += 512
                    tot["gpu_bit"] += int(eg.sum()); tot["gpu_blk"] += int(eg.any(axis=1).sum())
                    tot["orc_bit"] += int(eo.sum()); tot["orc_blk"] += int(eo.any(axis=1).sum())
                    tot["ref_bit"] += int(er.sum()); tot["ref_blk"] += int(er.any(axis=1).sum())
                    tot["diff_gpu_orc"] += int((u_gpu != u_orc).sum())
                    tot["diff_ref_orc"] += int((u_ref != u_orc).sum())
                    d_cw.append(eg.sum(axis=1).astype(np.float64) - eo.sum(axis=1))
                d = np.concatenate(d_cw) / k                      # per-codeword BER difference, paired
                nb = tot["cw"] * k
                row["paired"] = {
                    "codewords": tot["cw"], "oracle": "libm, reference list orders" + (", build without sb_math.h" if pure else ""),
                    "oracle_seconds": t_orc,
                    "bit_errors": {"gpu_default": tot["gpu_bit"], "gpu_sum_order_reference": tot["ref_bit"], "oracle": tot["orc_bit"]},
                    "block_errors": {"gpu_default": tot["gpu_blk"], "gpu_sum_order_reference": tot["ref_blk"], "oracle": tot["orc_blk"]},
                    "ber": {"gpu_default": tot["gpu_bit"] / nb, "oracle": tot["orc_bit"] / nb},
                    "delta_ber_gpu_default_minus_oracle": float(d.mean()),
                    "delta_ber_ci95": float(1.96 * d.std(ddof=1) / np.sqrt(len(d))),
                    "hard_bits_differing": {"gpu_default_vs_oracle": tot["diff_gpu_orc"],
                                            "gpu_sum_order_reference_vs_oracle": tot["diff_ref_orc"]},
                    "bits": nb}
            rule_out["points"].append(row)
            print(rule, json.dumps(row), flush=True)
        # ---- soft outputs after 20 iterations vs the libm oracle ------------------------------------------------
        dev = []
        for e in (1.0, 2.5):
            config.seed = 9000 + int(10 * e)
            b, llr = draw(e, 64)
            llr_np = llr.cpu().numpy()
            soft = O.LDPC5GDecoderRef(enc_r, cn_update=rule, num_iter=20, hard_out=False, return_infobits=False)
            x_orc = soft(llr_np, math_mode=0, order="reference", num_threads=cores, pure=pure)
            x_km = None if pure else soft(llr_np, math_mode=1, order="reference", num_threads=cores)
            entry = {"ebno_db": e, "codewords": 64}
            for name, d_ in (("gpu_default", LDPC5GDecoder(enc, num_iter=20, cn_update=rule, hard_out=False, return_infobits=False)),
                             ("gpu_sum_order_reference", LDPC5GDecoder(enc, num_iter=20, cn_update=rule, hard_out=False,
                                                                     return_infobits=False, sum_order="reference"))):
                x = d_(llr).cpu().numpy()
                rel = np.abs(x - x_orc) / np.maximum(np.abs(x_orc), 1e-6)
                entry[name] = {"share_rel_diff_gt_1e-4": float((rel > 1e-4).mean()), "max_abs_diff": float(np.abs(x - x_orc).max()),
                               "hard_decisions_differing": int(((x > 0) != (x_orc > 0)).sum()),
                               "bit_exact_vs_libm_oracle": bool(np.array_equal(x, x_orc))}
                if x_km is not None and name == "gpu_sum_order_reference":
                    entry[name]["bit_exact_vs_kernel_math_oracle_reference_order"] = bool(np.array_equal(x, x_km))
            dev.append(entry)
            print(rule, "soft", json.dumps(entry), flush=True)
        rule_out["soft_deviation_20_iterations"] = dev
        out["rules"][rule] = rule_out
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
