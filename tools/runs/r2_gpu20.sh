#!/bin/bash
# k_bin_segsum_pair capped at 96 registers (5 CTAs per SM): step time at 8192 and 65 536 walkers (before: 0.2401 / 1.2775 ms)
mkdir -p gpurun_out
for w in 8192 8192 65536; do
  timeout 300 python bench.py --steps 50 --warmup 5 --walkers $w --secondary none --no-cpu-baseline > gpurun_out/bench_cap_w$w.json 2> gpurun_out/bench_cap_w$w.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_cap_w$w.json'))
print('walkers $w', d['ms_per_step'], d['value'], {k:round(v['ms_per_launch'],4) for k,v in d['roofline']['kernels'].items()})
PY
done
timeout 600 python -m pytest tests/test_gpu_pipes.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2
