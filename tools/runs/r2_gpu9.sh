#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 --deselect tests/test_gpu_multi.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout 600 python tools/r2_chi2_sweep.py planck trsm > gpurun_out/chi2_sweep_trsm.jsonl 2> gpurun_out/chi2_sweep_trsm.err
python - <<'PY'
import json
for ln in open('gpurun_out/chi2_sweep_trsm.jsonl'):
    d=json.loads(ln); print(d.get('W'), d.get('variant'), d.get('ms'), d.get('frac'), d.get('bitwise'), d.get('error'))
PY
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
timeout 300 python bench.py --steps 20 --warmup 5 --walkers 8192 --secondary none --no-cpu-baseline > gpurun_out/bench_n1_w8192.json 2> gpurun_out/bench_n1_w8192.err
CSL_NO_FOLD_ROWS=1 timeout 300 python bench.py --steps 20 --warmup 5 --walkers 8192 --secondary none --no-cpu-baseline > gpurun_out/bench_n1_w8192_nofold.json 2> gpurun_out/bench_n1_w8192_nofold.err
python - <<'PY'
import json
for f in ('bench_n1','bench_n1_w8192','bench_n1_w8192_nofold'):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'], {k:(round(v['ms_per_launch'],4),round(v['frac_of_fp64_peak'],3)) for k,v in d['roofline']['kernels'].items()})
    for s in d.get('secondary') or []:
        print('   ', s['name'], s.get('error'), s.get('value'), s.get('ms_per_step'), (s.get('roofline') or {}).get('frac'))
PY
