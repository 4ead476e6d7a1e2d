#!/bin/bash
# final single-GPU evidence of round 2: GPU tests, the bench line, the reference arm, ncu summaries
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print('n1', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'], d['cpu_baseline'])
for s in d['secondary']: print('  ', s['name'], s.get('error'), s.get('value'), s.get('ms_per_step'), (s.get('roofline') or {}).get('frac'))
r=json.load(open('gpurun_out/r2_bench_ref.json')); print('ref', r['value'], r['cpu_baseline'])
PY
bash tools/r2_profile.sh
