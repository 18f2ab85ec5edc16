#!/usr/bin/env python
"""bench.py -- headline benchmark of BASELINE.json: coded bits/s of LDPC5G BP decoding, n=8448 (k=4224), 20
iterations, batch 4096 per GPU (weak scaling over 1/2/4/8 B200; one process per GPU, NCCL all-reduce of the four
int64 error counters per step, as sim_ber's replicas do: /root/reference/src/sionna/phy/utils/misc.py:614-655).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--cn-update boxplus-phi|minsum|...] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" = one LDPC5GDecoder call on a [4096, 8448] fp32 logit tensor resident in HBM + error counting (+ the
counter all-reduce for N > 1). Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_INFO, N_CODE, BATCH, NUM_ITER = 4224, 8448, 4096, 20
EBNO_DB = 2.0
# SURVEY.md 8(d) / BASELINE.md: algorithmic bytes per codeword for flooding BP, one fp32 message per edge read +
# written once per iteration, one channel LLR read per VN per iteration, plus compulsory I/O.
E_EDGES, N_VNS = 40320, 8832
ALG_BYTES_PER_CW = NUM_ITER * (8 * E_EDGES + 4 * N_VNS) + 4 * N_CODE + 4 * K_INFO      # 7 208 448


def measured_traffic(cn_update):
    """DRAM bytes per launch of the decode kernel from the committed `ncu --set full` capture (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        if cn_update in d:
            return d[cn_update]["traffic_bytes_per_launch"]
    return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons while the timed region runs: NVML (`pynvml`, ~20 ms period) when it is
    importable, else `nvidia-smi` polling (the recipe's clocks line, ~5 samples/s)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()   # samples: (sm_mhz, max_mhz, set(reasons))

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        masks = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._halt.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.samples.append((int(sm), int(mx), {n for n, m in masks.items() if bits & m}))
            self._halt.wait(0.02)
        nv.nvmlShutdown()

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [v.strip() for v in out.split(",")]
                if len(f) >= 6 and f[0].isdigit() and f[1].isdigit():
                    self.samples.append((int(f[0]), int(f[1]),
                                         {n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")}))
            except Exception:
                pass
            self._halt.wait(0.2)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max((s[1] for s in self.samples), default=None),
                "reasons": sorted(set().union(*[s[2] for s in self.samples])) if self.samples else [],
                "samples": len(sm)}


def make_inputs(seed, batch, device=None):
    """Seeded synthetic channel logits for the all-zero codeword (a valid codeword of the linear code) sent with
    the reference's BPSK-equivalent mapping over AWGN at Eb/N0 = 2 dB: logit = log p(1)/p(0) = 4 y / no with
    y = -1 + w. Decoder throughput is data independent (no early stopping, decoding.py:105-107)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    no = 1.0 / (10 ** (EBNO_DB / 10) * (K_INFO / N_CODE))
    y = -1.0 + rng.standard_normal((batch, N_CODE), dtype=np.float32) * np.float32(np.sqrt(no / 2))
    return (np.float32(4.0 / no) * y).astype(np.float32)


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores. TensorFlow (the reference's backend) is not
    installable offline, so this times the CPU restatement (oracle/, libm math, all host threads) on a bounded
    sample of the same workload: SAMPLE codewords of the [4096, 8448] batch per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import ldpc as O
    cores = os.cpu_count() or 1
    sample = max(cores * 2, 32)
    enc = O.LDPC5GEncoderRef(K_INFO, N_CODE)
    dec = O.LDPC5GDecoderRef(enc, cn_update=args.cn_update, num_iter=NUM_ITER)
    llr = make_inputs(1234, sample)
    for _ in range(args.warmup):
        dec(llr[:cores], num_threads=cores)
    t0 = time.perf_counter()
    errs = 0
    for _ in range(args.steps):
        u_hat = dec(llr, num_threads=cores)
        errs += int(u_hat.sum())
    dt = time.perf_counter() - t0
    val = sample * N_CODE * args.steps / dt
    line = {"metric": "coded bits/s, LDPC5G n=8448 k=4224 BP-20 decode", "value": val, "unit": "coded bits/s",
            "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"LDPC5GDecoder(LDPC5GEncoder(4224,8448)) {args.cn_update} 20 it, "
                                   f"{sample}-codeword sample of the batch-4096 workload per step"},
            "cpu_baseline": {"value": val, "unit": "coded bits/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} codewords/step x {args.steps} steps, oracle/ldpc_bp_ref.c libm mode, "
                                       f"OpenMP {cores} threads (TensorFlow reference not installable offline)"},
            "e2e": {"value": val, "unit": "coded bits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_RESULT_FD = None


def emit(line):
    """Writes the ONE result line to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    # stdout carries exactly one JSON line: everything else that might write to fd 1 (NCCL's version banner, library
    # printf, build logs) is routed to stderr for the lifetime of the process
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cn-update", default="boxplus-phi",
                    choices=["boxplus-phi", "boxplus", "minsum", "offset-minsum"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="codewords for the cpu_baseline leg (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ebno-db", type=float, default=EBNO_DB,
                    help="Eb/N0 of the synthetic inputs (default 2 dB, SURVEY.md section 8d). The boxplus-phi kernel skips "
                         "provably-zero phi terms of saturated messages, so its speed depends on how early codewords converge")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from sionna_b200 import _lib
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep NCCL's banner / debug lines off stdout
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "--gpus must match the torchrun world size"

    from sionna_b200.phy import config as sb_config
    from sionna_b200.phy.mapping import BinarySource, Mapper, Demapper
    from sionna_b200.phy.channel import AWGN
    from sionna_b200.phy.utils import ebnodb2no, ErrorCounter

    enc = LDPC5GEncoder(K_INFO, N_CODE)
    dec = LDPC5GDecoder(enc, cn_update=args.cn_update, num_iter=NUM_ITER, hard_out=True, return_infobits=True)
    assert dec.on_chip and dec.num_edges == E_EDGES and dec.num_vns == N_VNS

    # Synthetic inputs, generated once on the device by the package's own transmit chain (per-rank Philox stream):
    # BinarySource -> LDPC5GEncoder -> QPSK Mapper -> AWGN(Eb/N0 = 2 dB) -> Demapper("app"). Two distinct input sets
    # (2 x 138 MB > 126 MB L2) are alternated between steps: nothing the decoder reads can be an L2 hit left over
    # from the previous step.
    sb_config.seed = 100 + 1000 * rank
    no = ebnodb2no(args.ebno_db, 2, K_INFO / N_CODE)
    src, mapper, demapper, awgn = BinarySource(), Mapper("qam", 2), Demapper("app", "qam", 2), AWGN()
    d_u, d_in = [], []
    for _ in range(2):
        u = src([BATCH, K_INFO])
        d_u.append(u)
        d_in.append(demapper(awgn(mapper(enc(u)), no), no))
    h_in = [t.cpu().pin_memory() for t in d_in]
    counter = ErrorCounter(dev)                                   # device int64[4]: bit errors, block errors, bits, blocks
    reduced = torch.zeros(4, dtype=torch.int64, device=dev)

    def step(i):
        u_hat = dec(d_in[i & 1])
        counter.update(d_u[i & 1], u_hat)                         # sb_count_errors
        if world > 1:
            dist.all_reduce(reduced.copy_(counter.counters), op=dist.ReduceOp.SUM)
        return u_hat

    for i in range(args.warmup):
        step(i)
    counter.reset()
    torch.cuda.synchronize()

    # ---- timed region: K steps; CUDA events on the launching stream around every decode call for the roofline --------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.lib().sb_launch_count()
    t_start.record()
    for i in range(args.steps):
        ev[i][0].record()
        u_hat = dec(d_in[i & 1])
        ev[i][1].record()
        counter.update(d_u[i & 1], u_hat)
        if world > 1:
            dist.all_reduce(reduced.copy_(counter.counters), op=dist.ReduceOp.SUM)
    t_stop.record()
    launches = _lib.lib().sb_launch_count() - launches0
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = t_start.elapsed_time(t_stop)
    kern_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * BATCH * N_CODE * args.steps / (ms_total * 1e-3)
    if world > 1:
        dist.all_reduce(reduced.copy_(counter.counters), op=dist.ReduceOp.SUM)
        totals = reduced.cpu().tolist()
    else:
        totals = counter.counters.cpu().tolist()

    # ---- end to end through the public API with HOST buffers: every step copies its logits from pinned host memory to
    # the device, decodes, and copies the decoded bits back to pinned host memory. Two streams are used round-robin so the
    # copies of one step overlap the kernel of the other (a double-buffered serving loop).
    h_out = [torch.empty((BATCH, K_INFO), dtype=torch.float32).pin_memory() for _ in range(2)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    e2e_steps = max(4, min(args.steps, 10))

    def e2e_step(i):
        with torch.cuda.stream(streams[i & 1]):
            x = h_in[i & 1].to(dev, non_blocking=True)
            h_out[i & 1].copy_(dec(x), non_blocking=True)

    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s_ in streams:
        s_.wait_event(e0)
    for i in range(e2e_steps):
        e2e_step(i)
    for s_ in streams:
        torch.cuda.current_stream().wait_stream(s_)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * BATCH * N_CODE * e2e_steps / (float(t.item()) * 1e-3)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        achieved = ALG_BYTES_PER_CW * BATCH / (kern_ms * 1e-3) / 1e9
        c = totals
        line = {
            "metric": "coded bits/s, LDPC5G n=8448 k=4224 BP-20 decode", "value": value, "unit": "coded bits/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[1]: LDPC5GDecoder(LDPC5GEncoder(4224,8448)), cn_update={args.cn_update}, "
                                   f"20 BP iterations, batch 4096 per GPU, QPSK/AWGN Eb/N0 {args.ebno_db:g} dB",
                       "cn_update": args.cn_update, "batch_per_gpu": BATCH, "parallelism": f"replicas x{world}", "ebno_db": args.ebno_db,
                       "l2": "2 alternating input sets of 138 MB each (> 126 MB L2)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(args.cn_update), "peak_source": peak_src,
                         "kernel": "ldpc_bp_qc_kernel",
                         "kernel_ms": kern_ms, "alg_bytes_per_launch": ALG_BYTES_PER_CW * BATCH},
            "e2e": {"value": e2e_val, "unit": "coded bits/s", "h2d_bytes_per_step": BATCH * N_CODE * 4,
                    "d2h_bytes_per_step": BATCH * K_INFO * 4, "steps": e2e_steps,
                    "pipeline": "2 CUDA streams, double-buffered pinned host buffers"},
            "gpu_launches": launches, "clocks": clocks,
            "ber": {"bit_errors": c[0], "block_errors": c[1], "bits": c[2], "blocks": c[3]},
        }
        # the north-star's min-sum rule on the same inputs (the headline rule above is the reference's default boxplus-phi)
        if args.cn_update != "minsum":
            dec_ms = LDPC5GDecoder(enc, cn_update="minsum", num_iter=NUM_ITER, hard_out=True, return_infobits=True)
            for i in range(3):
                dec_ms(d_in[i & 1])
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            a0.record()
            for i in range(reps):
                dec_ms(d_in[i & 1])
            a1.record()
            torch.cuda.synchronize()
            ms = a0.elapsed_time(a1) / reps
            line["other_rules"] = {"minsum": {"value": BATCH * N_CODE / (ms * 1e-3), "unit": "coded bits/s (1 GPU, device)",
                                              "kernel_ms": ms,
                                              "roofline_frac": ALG_BYTES_PER_CW * BATCH / (ms * 1e-3) / 1e9 / peak,
                                              "traffic": measured_traffic("minsum")}}
        if not args.no_cpu_baseline and world == 1:           # reported baseline: rank 0 at N = 1 only
            from oracle import ldpc as O
            cores = os.cpu_count() or 1
            sample = min(BATCH, args.cpu_sample or max(32 * cores, 1024))   # ~10 s of CPU work on 128 threads
            ref = O.LDPC5GDecoderRef(O.LDPC5GEncoderRef(K_INFO, N_CODE), cn_update=args.cn_update, num_iter=NUM_ITER)
            x = h_in[0][:sample].numpy()
            ref(x[:cores], num_threads=cores)
            t0 = time.perf_counter()
            u_ref = ref(x, num_threads=cores)
            dt = time.perf_counter() - t0
            u_gpu = dec(d_in[0][:sample].contiguous()).cpu().numpy()
            line["cpu_baseline"] = {"value": sample * N_CODE / dt, "unit": "coded bits/s", "cores": cores,
                                    "kind": "port",
                                    "sample": f"first {sample} codewords of the step-0 batch, oracle/ldpc_bp_ref.c libm "
                                              f"mode, {cores} OpenMP threads",
                                    "bit_mismatch_vs_gpu": int((u_ref != u_gpu).sum())}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
