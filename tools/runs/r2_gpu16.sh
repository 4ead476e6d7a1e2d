#!/bin/bash
# 8 GPUs: the full bench line (literal + weak + secondary), then the literal config alone at 20 / 50 / 200 steps
# with and without the gate
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "full rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n8.json'))
w=d['weak']
print('strong', d['ms_per_step'], d['value'], d['launches'], d['exchange_parity'], d['clocks'])
print('weak', w['ms_per_step'], w['value'])
for s in d['secondary']: print('  ', s['name'], s.get('error'), s.get('value'), s.get('ms_per_step'), (s.get('roofline') or {}).get('frac'))
PY
B="--warmup 5 --secondary none --no-cpu-baseline --no-weak"
for cfg in "20 A=1 gate" "20 CSL_BENCH_NO_GATE=1 nogate" "50 A=1 gate" "200 A=1 nogate"; do
  set -- $cfg
  env $2 timeout 300 $TR bench.py --gpus 8 $B --steps $1 > gpurun_out/g8_s$1_$3.json 2> gpurun_out/g8_s$1_$3.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/g8_s$1_$3.json')); print('N=8 steps $1 $3', d['ms_per_step'], d['value'], d.get('launches'))
except Exception as e: print('N=8 steps $1 $3 FAILED', e)"
done
