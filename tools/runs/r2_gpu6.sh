#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipes.py tests/test_gpu_native_mh.py -q > gpurun_out/pytest_new.log 2>&1; echo "new tests rc=$?"; tail -8 gpurun_out/pytest_new.log
timeout 600 python tools/r2_chi2_sweep.py planck trsm > gpurun_out/chi2_sweep_trsm.jsonl 2> gpurun_out/chi2_sweep_trsm.err; echo "sweep trsm rc=$?"; tail -3 gpurun_out/chi2_sweep_trsm.err
python - <<'PY'
import json
for ln in open('gpurun_out/chi2_sweep_trsm.jsonl'):
    d=json.loads(ln); print(d.get('W'), d.get('variant'), d.get('ms'), d.get('frac'), d.get('bitwise'), d.get('error'))
PY
timeout 600 python bench.py --steps 20 --warmup 5 --secondary spt_multifreq,spt_lowl_mh --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_a.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], {k:(v['ms_per_launch'],round(v['frac_of_fp64_peak'],3)) for k,v in d['roofline']['kernels'].items()})
for s in d['secondary']:
    print(s['name'], s.get('error'), s.get('value'), s.get('ms_per_step'), s.get('roofline',{}).get('frac'), s.get('roofline',{}).get('kernels') or s.get('roofline',{}).get('stages_ms'))
PY
