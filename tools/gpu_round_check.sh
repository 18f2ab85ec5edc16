#!/bin/bash
# One GPU session that re-validates and re-measures everything the round reports (outputs under gpurun_out/).
set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 > gpurun_out/r01_bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 300 gpurun_out/r01_bench_1gpu.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference_arm.json 2>/dev/null; tail -c 200 gpurun_out/r01_bench_reference_arm.json
python bench.py --steps 10 --warmup 3 --cn-update minsum --no-cpu-baseline > gpurun_out/r01_bench_1gpu_minsum.json 2>/dev/null
python tools/bench_phy_kernels.py --out gpurun_out/r01_phy_kernels.json 2>&1 | tail -12
python tools/pusch_sim.py --out gpurun_out/r01_pusch_1gpu.json 2>&1 | tail -1 | cut -c400-560
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_phi_v8.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ldpc_bp_qc_kernel -s 3 -c 1 -f -o gpurun_out/phi_v8 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ldpc_bp_qc_kernel -s 3 -c 1 -f -o gpurun_out/minsum_v8 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --cn-update minsum > /dev/null 2>&1
bash tools/ncu_phy_kernels.sh > /dev/null 2>&1
ls gpurun_out/*.ncu-rep | wc -l
