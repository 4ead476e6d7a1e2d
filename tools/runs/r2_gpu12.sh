#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for w in 65536 8192; do
  for off in 0 1; do
    CSL_NO_STAGE_TABLES=$off timeout 300 python bench.py --steps 50 --warmup 5 --walkers $w --secondary none --no-cpu-baseline > gpurun_out/bench_stage${off}_w$w.json 2> gpurun_out/bench_stage${off}_w$w.err
    python - <<PY
import json
d=json.load(open('gpurun_out/bench_stage${off}_w$w.json'))
print('walkers $w NO_STAGE=$off', d['ms_per_step'], d['value'], d['acceptance'], d['roofline']['other_stages_ms'])
PY
  done
done
timeout 300 python tools/r2_group_sweep.py time 32768 > gpurun_out/group_time.jsonl 2>gpurun_out/group_time.err; cat gpurun_out/group_time.jsonl
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_trigemm --csv --log-file gpurun_out/group_ncu.csv python tools/r2_group_sweep.py ncu 32768 > gpurun_out/group_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/group_ncu.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hd=rows[h]; idc=hd.index('ID'); mn=hd.index('Metric Name'); mv=hd.index('Metric Value')
d={}
for r in rows[h+1:]:
    if len(r)>mv: d.setdefault(int(r[idc]),{})[r[mn]]=r[mv]
groups=[8,16,32,48,64,96,146,256]
ids=sorted(d)
for gi,g in enumerate(groups):
    i=ids[gi*4+3] if gi*4+3 < len(ids) else None
    if i is not None: print(g, d[i])
PY
