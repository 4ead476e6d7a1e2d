#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipes.py -q > gpurun_out/pytest_pipes.log 2>&1; echo "pipes rc=$?"; tail -8 gpurun_out/pytest_pipes.log
timeout 600 python tools/r2_chi2_sweep.py planck trsm > gpurun_out/chi2_sweep_trsm.jsonl 2> gpurun_out/chi2_sweep_trsm.err; echo "sweep trsm rc=$?"; tail -3 gpurun_out/chi2_sweep_trsm.err
for k in spt spt_lowl; do timeout 600 python tools/r2_bin_sweep.py $k > gpurun_out/bin_sweep_$k.jsonl 2> gpurun_out/bin_sweep_$k.err; echo "bin sweep $k rc=$?"; tail -3 gpurun_out/bin_sweep_$k.err; done
