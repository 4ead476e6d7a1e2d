#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515"
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "full rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n8.json'))
w=d['weak']
print('strong', d['ms_per_step'], d['value'], d['e2e']['value'], d['launches'], d['exchange_parity'], d['clocks'])
print('weak', w['ms_per_step'], w['value'], w['e2e']['value'])
for s in d['secondary']: print('  ', s['name'], s.get('error'), s.get('value'), s.get('ms_per_step'), (s.get('roofline') or {}).get('frac'))
PY
timeout 300 $TR bench.py --gpus 8 --steps 200 --warmup 5 --secondary none --no-cpu-baseline > gpurun_out/r2_bench_n8_s200.json 2> gpurun_out/r2_bench_n8_s200.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n8_s200.json')); print('200 steps: strong', d['ms_per_step'], d['value'], 'weak', d['weak']['ms_per_step'], d['weak']['value'])"
