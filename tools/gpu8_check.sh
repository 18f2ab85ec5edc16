#!/bin/bash
# 8-GPU session of round 2 (gpurun --gpus 8): headline bench, PUSCH workload and the Monte-Carlo PUSCH link, one rank per GPU.
R=r02
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
$TR --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/${R}_bench_8gpu.json 2> gpurun_out/bench_8gpu.err
tail -c 600 gpurun_out/${R}_bench_8gpu.json; echo
$TR --master-port 29512 bench.py --gpus 8 --workload pusch --steps 5 --no-cpu-baseline > gpurun_out/${R}_bench_pusch_8gpu.json 2> gpurun_out/bench_pusch_8gpu.err
head -c 400 gpurun_out/${R}_bench_pusch_8gpu.json; echo
$TR --master-port 29513 tools/pusch_sim.py --out gpurun_out/${R}_pusch_8gpu.json 2>&1 | tail -4
wc -l gpurun_out/${R}_bench_8gpu.json gpurun_out/${R}_bench_pusch_8gpu.json
