#!/bin/bash
mkdir -p gpurun_out
export SLAB_REPEATS=6
timeout 900 python tools/r2_chi2_sweep.py planck trsm > gpurun_out/chi2_sweep_slab.jsonl 2> gpurun_out/chi2_sweep_slab.err
timeout 600 python tools/r2_chi2_sweep.py spt trsm >> gpurun_out/chi2_sweep_slab.jsonl 2>> gpurun_out/chi2_sweep_slab.err
tail -3 gpurun_out/chi2_sweep_slab.err
python - <<'PY'
import json, collections
agg=collections.OrderedDict()
for ln in open('gpurun_out/chi2_sweep_slab.jsonl'):
    d=json.loads(ln)
    k=(d.get('kind'), d.get('W'), json.dumps(d.get('variant')))
    a=agg.setdefault(k,{'ms':[], 'bit':[], 'rel':[], 'err':[]})
    a['ms'].append(d.get('ms')); a['bit'].append(d.get('bitwise')); a['rel'].append(d.get('maxrel')); a['err'].append(d.get('error'))
for k,a in agg.items():
    ms=[m for m in a['ms'] if m is not None]
    print(k, 'n=%d'%len(a['ms']), 'ms min/max', min(ms) if ms else None, max(ms) if ms else None, 'bitwise all' , all(a['bit']), 'maxrel', max([r for r in a['rel'] if r is not None] or [None]), [e for e in a['err'] if e][:1])
PY
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
export CSL_BENCH_PREROLL=3
CMD4="python bench.py --chi2 trsm --steps 2 --warmup 3 --secondary none --no-cpu-baseline"
$CMD4 > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_trsm_chi2' -s 4 -c 1 -f -o /tmp/prof_r2_trsm $CMD4 > gpurun_out/ncu_full_trsm.log 2>&1
echo "full trsm rc=$?"
python tools/ncu_summary.py full /tmp/prof_r2_trsm.ncu-rep k_trsm_chi2 > gpurun_out/r2_ncu_full_k_trsm_chi2_slab.txt
cat gpurun_out/r2_ncu_full_k_trsm_chi2_slab.txt
unset CSL_BENCH_PREROLL
timeout 300 python bench.py --steps 50 --warmup 5 --secondary none --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
timeout 300 python bench.py --chi2 trsm --steps 50 --warmup 5 --secondary none --no-cpu-baseline > gpurun_out/bench_c_trsm.json 2> gpurun_out/bench_c_trsm.err
python - <<'PY'
import json
for f in ('bench_c','bench_c_trsm'):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, d['ms_per_step'], d['value'], {k:(round(v['ms_per_launch'],4), round(v['frac_of_fp64_peak'],3)) for k,v in d['roofline']['kernels'].items()})
PY
