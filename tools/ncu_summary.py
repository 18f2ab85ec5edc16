"""Write a compact text summary of an .ncu-rep (raw page) for profiles/.  usage: ncu_summary.py rep.ncu-rep out.txt"""
import csv, subprocess, sys, io
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]
with open(out, "w") as f:
    for r in rows[2:]:
        f.write("=" * 100 + "\n")
        for k in KEYS:
            for i, h in enumerate(hdr):
                if h == k:
                    f.write(f"{h:75s} {r[i]:>24s} {units[i]}\n")
        stalls = [(float(r[i] or 0), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
        f.write("stall reasons (average warps stalled per issue-active cycle):\n")
        for v, h in sorted(stalls, reverse=True)[:8]:
            f.write(f"    {h:75s} {v:8.2f}\n")
print(open(out).read())
