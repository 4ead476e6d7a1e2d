#!/bin/bash
mkdir -p gpurun_out
export CSL_BENCH_PREROLL=3
CMD="python bench.py --steps 4 --warmup 3 --walkers 8192 --secondary none --no-cpu-baseline"
$CMD > gpurun_out/plain_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_small.csv $CMD > gpurun_out/ncu_launches_small.log 2>&1
echo "launch list rc=$?"
python tools/ncu_summary.py launches gpurun_out/r2_launches_small.csv > gpurun_out/r2_launches_bench_planck_lite_w8192.csv
cat gpurun_out/r2_launches_bench_planck_lite_w8192.csv
