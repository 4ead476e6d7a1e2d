#!/bin/bash
# PDL along the native stretch loop: parity (GPU tests) + A/B at 65 536 and 8192 walkers
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for w in 65536 8192; do
  for nopdl in 0 1; do
    CSL_NO_PDL=$nopdl timeout 300 python bench.py --steps 50 --warmup 5 --walkers $w --secondary none --no-cpu-baseline > gpurun_out/bench_pdl${nopdl}_w$w.json 2> gpurun_out/bench_pdl${nopdl}_w$w.err
    python - <<PY
import json
d=json.load(open('gpurun_out/bench_pdl${nopdl}_w$w.json'))
print('walkers $w NO_PDL=$nopdl', d['ms_per_step'], d['value'], d['acceptance'])
PY
  done
done
