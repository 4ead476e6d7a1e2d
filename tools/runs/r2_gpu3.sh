#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipes.py -q -x > gpurun_out/pytest_pipes.log 2>&1; echo "pipes rc=$?"; tail -15 gpurun_out/pytest_pipes.log
for k in planck spt; do timeout 600 python tools/r2_chi2_sweep.py $k > gpurun_out/chi2_sweep_$k.jsonl 2> gpurun_out/chi2_sweep_$k.err; echo "sweep $k rc=$?"; tail -3 gpurun_out/chi2_sweep_$k.err; done
for k in spt spt_lowl planck_dmma; do timeout 600 python tools/r2_bin_sweep.py $k > gpurun_out/bin_sweep_$k.jsonl 2> gpurun_out/bin_sweep_$k.err; echo "bin sweep $k rc=$?"; tail -3 gpurun_out/bin_sweep_$k.err; done
for p in 4 8 50; do
CSL_CLOCK_PERIOD_MS=$p timeout 300 python bench.py --steps 20 --warmup 5 --secondary none --no-cpu-baseline > gpurun_out/bench_clk$p.json 2> gpurun_out/bench_clk$p.err
python -c "
import json;d=json.load(open('gpurun_out/bench_clk$p.json'));print('period $p ms:',d['ms_per_step'],d['value'],d['clocks'])"
done
