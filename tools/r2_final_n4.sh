#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n4.json'))
print('n4', d['ms_per_step'], d['value'], d['e2e']['value'], d.get('exchange_parity'))
print('   weak', d['weak']['ms_per_step'], d['weak']['value'])
for s in d['secondary']: print('  ', s['name'], s.get('error'), s.get('value'), s.get('ms_per_step'), (s.get('roofline') or {}).get('frac'))
PY
