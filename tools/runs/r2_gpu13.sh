#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/r2_chi2_sweep.py planck trsm > gpurun_out/chi2_sweep_slab.jsonl 2> gpurun_out/chi2_sweep_slab.err
timeout 600 python tools/r2_chi2_sweep.py spt trsm >> gpurun_out/chi2_sweep_slab.jsonl 2>> gpurun_out/chi2_sweep_slab.err
tail -3 gpurun_out/chi2_sweep_slab.err
python - <<'PY'
import json
for ln in open('gpurun_out/chi2_sweep_slab.jsonl'):
    d=json.loads(ln); print(d.get('kind'), d.get('W'), d.get('variant'), d.get('ms'), d.get('frac'), d.get('bitwise'), d.get('maxrel'), d.get('error'))
PY
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for w in 65536 8192; do
  timeout 300 python bench.py --steps 50 --warmup 5 --walkers $w --secondary none --no-cpu-baseline > gpurun_out/bench_b_w$w.json 2> gpurun_out/bench_b_w$w.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_b_w$w.json'))
print('walkers $w', d['ms_per_step'], d['value'], d['acceptance'], d['roofline']['other_stages_ms'])
PY
done
