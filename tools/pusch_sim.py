#!/usr/bin/env python
"""BASELINE.json configs[4]: end-to-end 5G NR PUSCH link (PUSCHTransmitter -> TDL channel -> PUSCHReceiver), global
batch 8192 sharded over the GPUs of one node, one NCCL all-reduce of the four error counters per batch (sim_ber).

  python tools/pusch_sim.py --out profiles/r01_pusch_1gpu.json
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \\
         tools/pusch_sim.py --out profiles/r01_pusch_2gpu.json

Link: 1 UE, 2 layers on 2 antenna ports, 8 receive antennas, 16 PRBs x 14 symbols, MCS 14 (16-QAM, r = 0.54), DMRS
type 1 with one additional position, TDL-B 100 ns block fading; LS channel estimation with linear interpolation and CDM
de-spreading, LMMSE detection, 20 BP iterations. Reports BER/BLER per Eb/N0 and the whole-job rate in transport
blocks/s and information bits/s (all ranks, slowest rank's device time).
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
    ap.add_argument("--out", default=None)
    ap.add_argument("--global-batch", type=int, default=8192)
    ap.add_argument("--max-batches", type=int, default=4)
    ap.add_argument("--ebno-dbs", default="-4,-2,0,2,4")
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep NCCL's banner / debug lines off stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from sionna_b200.phy import config
    from sionna_b200.phy.nr import PUSCHConfig, PUSCHTransmitter, PUSCHReceiver
    from sionna_b200.phy.channel import ApplyOFDMChannel, TDL, subcarrier_frequencies, cir_to_ofdm_channel
    from sionna_b200.phy.utils import ebnodb2no, sim_ber
    config.device = torch.device("cuda", local)
    config.seed = 2024
    config.rank_offset = rank                                   # per-rank Philox stream
    pc = PUSCHConfig(num_layers=2, num_antenna_ports=2)
    pc.carrier.n_size_grid = 16
    pc.dmrs.additional_position = 1
    pc.tb.mcs_index = 14
    tx = PUSCHTransmitter(pc)
    rx = PUSCHReceiver(tx)
    rg = tx.resource_grid
    tdl = TDL("B", 100e-9, 3.5e9, num_rx_ant=8, num_tx_ant=2)
    freqs = subcarrier_frequencies(rg.fft_size, rg.subcarrier_spacing)
    chan = ApplyOFDMChannel()
    m, r = pc.tb.num_bits_per_symbol, pc.tb_size / pc.num_coded_bits
    per_rank = args.global_batch // world

    def mc_fun(batch_size, ebno_db):
        no = ebnodb2no(ebno_db, m, r, rg)
        x, b = tx(batch_size)
        a, tau = tdl(batch_size, rg.num_ofdm_symbols, 1.0)
        h = cir_to_ofdm_channel(freqs, a, tau, normalize=True)
        y = chan(x, h, no)
        return b, rx(y, no)

    ebnos = [float(v) for v in args.ebno_dbs.split(",")]
    mc_fun(per_rank, ebnos[0])                                  # warm-up (graph tables, allocator)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    # sim_ber divides max_mc_iter by the number of replicas (as the reference does): every rank runs max_batches steps
    ber, bler = sim_ber(mc_fun, ebnos, per_rank, args.max_batches * world, early_stop=False, verbose=(rank == 0),
                        distribute="all" if world > 1 else None)
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=config.device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    blocks = args.max_batches * len(ebnos) * per_rank * world
    if rank == 0:
        out = {"config": "configs[4]: PUSCH 2 layers / 2 ports, 8 rx antennas, 16 PRB, MCS 14, TDL-B 100 ns, LS(lin) + "
                         "LMMSE + BP-20", "n_gpus": world, "global_batch": per_rank * world,
               "tb_size": int(pc.tb_size), "num_coded_bits": int(pc.num_coded_bits), "ebno_db": ebnos,
               "ber": [float(v) for v in ber], "bler": [float(v) for v in bler],
               "transport_blocks_per_s": blocks / (ms.item() * 1e-3),
               "info_bits_per_s": blocks * pc.tb_size / (ms.item() * 1e-3), "ms_total": ms.item(),
               "timed_region": "tx + channel + rx + error counting (+ all-reduce), device time, max over ranks",
               "device": torch.cuda.get_device_name(local)}
        print(json.dumps(out), flush=True)
        if args.out:
            with open(os.path.join(ROOT, args.out) if not os.path.isabs(args.out) else args.out, "w") as f:
                json.dump(out, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
