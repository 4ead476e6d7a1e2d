#!/bin/bash
# round-2 ncu evidence (1 GPU): launch list of the bench step + --set full of the FP64 kernels.
# The .ncu-rep files are summarised ON THE BOX (tools/ncu_summary.py) and deleted: gpurun_out/ carries text only.
mkdir -p gpurun_out
export CSL_BENCH_PREROLL=3
export CSL_BENCH_NO_GATE=1      # ncu serialises launches: the host would never reach the release of the gate
CMD="python bench.py --steps 2 --warmup 3 --secondary none --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python tools/ncu_summary.py launches gpurun_out/r2_launches.csv > gpurun_out/r2_launches_bench_planck_lite.csv
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_bin_segsum|k_trigemm_chi2' -s 12 -c 3 -f -o /tmp/prof_r2_planck $CMD > gpurun_out/ncu_full_planck.log 2>&1
echo "full planck rc=$?"
python tools/ncu_summary.py full /tmp/prof_r2_planck.ncu-rep k_trigemm_chi2 > gpurun_out/r2_ncu_full_k_trigemm_chi2.txt
python tools/ncu_summary.py full /tmp/prof_r2_planck.ncu-rep k_bin_segsum > gpurun_out/r2_ncu_full_k_bin_segsum.txt
CMD3="python bench.py --workload spt_multifreq --walkers 16384 --steps 2 --warmup 3 --secondary none --no-cpu-baseline"
$CMD3 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_bin_residual|k_trigemm_chi2' -s 8 -c 2 -f -o /tmp/prof_r2_spt $CMD3 > gpurun_out/ncu_full_spt.log 2>&1
echo "full spt rc=$?"
python tools/ncu_summary.py full /tmp/prof_r2_spt.ncu-rep k_bin_residual > gpurun_out/r2_ncu_full_k_bin_residual_config3.txt
python tools/ncu_summary.py full /tmp/prof_r2_spt.ncu-rep k_trigemm_chi2 > gpurun_out/r2_ncu_full_k_trigemm_chi2_config3.txt
CMD4="python bench.py --chi2 trsm --steps 2 --warmup 3 --secondary none --no-cpu-baseline"
$CMD4 > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_trsm_chi2' -s 4 -c 1 -f -o /tmp/prof_r2_trsm $CMD4 > gpurun_out/ncu_full_trsm.log 2>&1
echo "full trsm rc=$?"
python tools/ncu_summary.py full /tmp/prof_r2_trsm.ncu-rep k_trsm_chi2 > gpurun_out/r2_ncu_full_k_trsm_chi2_slab.txt
wc -l gpurun_out/r2_*.txt gpurun_out/r2_launches_bench_planck_lite.csv
unset CSL_BENCH_NO_GATE CSL_BENCH_PREROLL
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/pipes tools/micro/pipes.cu && /tmp/pipes > gpurun_out/r2_pipes.txt 2>&1; cat gpurun_out/r2_pipes.txt
python tools/measure_peaks.py > gpurun_out/measure_peaks.log 2>&1; tail -1 gpurun_out/measure_peaks.log
python -c "
import __graft_entry__ as g
g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
