#!/bin/bash
# One GPU session that re-validates and re-measures everything round 2 reports (outputs under gpurun_out/).
#   gpurun --timeout 2400 -- 'bash tools/gpu_round_check.sh'
set -x
R=r02
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 400 gpurun_out/${R}_bench_1gpu.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2>/dev/null; tail -c 200 gpurun_out/${R}_bench_reference_arm.json
for w in qpsk_awgn ofdm_siso mimo_ofdm pusch; do
  python bench.py --workload $w --steps 10 > gpurun_out/${R}_bench_$w.json 2> gpurun_out/bench_$w.err
done
python tools/bench_phy_kernels.py --out gpurun_out/${R}_phy_kernels.json 2>&1 | tail -14
python tools/pusch_sim.py --out gpurun_out/${R}_pusch_1gpu.json 2>&1 | tail -1 | cut -c1-300
# launch lists (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_phi.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-links --no-variants --no-traffic > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 500 --csv --log-file gpurun_out/${R}_launches_pusch.csv python tools/pusch_sim.py --max-batches 2 --ebno-dbs 0,2 > /dev/null 2>&1
# full captures of the kernels the round worked on
cap() {  # name kernel-regex command...
  n=$1; k=$2; shift 2
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s ${SKIP:-2} -c 1 -f -o gpurun_out/${R}_$n "$@" > /dev/null 2>&1
}
cap ldpc_bp_phi ldpc_bp_qc_kernel python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-links --no-variants --no-traffic
cap ldpc_bp_minsum ldpc_bp_qc_kernel python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-links --no-variants --no-traffic --cn-update minsum
cap frontend_mimo ofdm_frontend_kernel python bench.py --workload mimo_ofdm --steps 2 --no-cpu-baseline
cap frontend_pusch ofdm_frontend_kernel python bench.py --workload pusch --steps 2 --no-cpu-baseline --batch 2048
SKIP=0 cap cir_apply cir_apply_kernel python tools/pusch_sim.py --max-batches 1 --ebno-dbs 0 --global-batch 2048
SKIP=0 cap lmmse_diag ofdm_lmmse_diag_kernel python tools/bench_phy_kernels.py --only ofdm_lmmse_4x16
SKIP=0 cap fft76 ofdm_fft_small_kernel python tools/bench_phy_kernels.py --only ofdm_demodulate_76
SKIP=1 cap fft4096 ofdm_fft_r16_kernel python tools/bench_phy_kernels.py --only ofdm_demodulate_4096
ls -la gpurun_out/*.ncu-rep | wc -l
# text summaries on the box (the .ncu-rep files come back too while they fit gpurun_out's 64 MiB)
for f in gpurun_out/${R}_*.ncu-rep; do python tools/ncu_summary.py $f ${f%.ncu-rep}_ncu.txt > /dev/null 2>&1; done
while [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; do rm -f "$(ls -S gpurun_out/*.ncu-rep | head -1)"; done
du -sh gpurun_out
