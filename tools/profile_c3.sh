#!/bin/sh
# ncu capture of the DMMA binning kernel on BASELINE config 3 (dense SPT-like windows), 65536 walkers
set -x
mkdir -p gpurun_out
CMD="python bench.py --workload spt_multifreq --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_bin_residual' -s 4 -c 1 -f -o gpurun_out/prof_r1_c3 $CMD > gpurun_out/ncu_c3.log 2>&1
tail -c 1500 gpurun_out/plain_c3.log
