#!/bin/sh
# ncu recipe of /opt/skills/guides/B200_PROFILING.md for the bench workload (run under gpurun, 1 GPU).
#   1. plain run must exit 0   2. launch list (gpu__time_duration)   3. --set full of the two big kernels
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_bin_segsum|k_trigemm_chi2' -s 6 -c 3 -f -o gpurun_out/prof_r1i $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/plain.log
ls -la gpurun_out
