#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print('n1', d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e'].get('repetitions'), d['roofline']['frac'], d['roofline'].get('traffic'), d['gpu_launches'])
for s in d['secondary']: print('  ', s['name'], s.get('error'), s.get('value'), s.get('ms_per_step'), (s.get('roofline') or {}).get('frac'))
PY
timeout 300 python -m pytest tests/test_gpu_native_mh.py tests/test_gpu_unbinned.py -m gpu -q 2>&1 | tail -2
