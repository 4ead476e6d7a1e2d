#!/bin/bash
# ncu --set full of the two FP64 kernels at 4096-walker launches (the 8192-walker ensemble one of 8 GPUs runs)
mkdir -p gpurun_out
export CSL_BENCH_PREROLL=3
export CSL_BENCH_NO_GATE=1
CMD="python bench.py --steps 2 --warmup 3 --walkers 8192 --secondary none --no-cpu-baseline"
$CMD > gpurun_out/plain_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_bin_segsum|k_trigemm_chi2' -s 12 -c 2 -f -o /tmp/prof_small $CMD > gpurun_out/ncu_full_small.log 2>&1
echo "full small rc=$?"
python tools/ncu_summary.py full /tmp/prof_small.ncu-rep k_trigemm_chi2 > gpurun_out/r2_ncu_full_k_trigemm_chi2_w4096.txt
python tools/ncu_summary.py full /tmp/prof_small.ncu-rep k_bin_segsum > gpurun_out/r2_ncu_full_k_bin_segsum_w4096.txt
cat gpurun_out/r2_ncu_full_k_trigemm_chi2_w4096.txt gpurun_out/r2_ncu_full_k_bin_segsum_w4096.txt
