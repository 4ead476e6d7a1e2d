#!/bin/bash
# 2 GPUs: multi-GPU tests + bench line + trsm sweep on one of them
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 1500 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/pytest_multi.log 2>&1; echo "multi rc=$?"; tail -15 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print('literal', d['value'], d['ms_per_step'], d['ms_per_rank'], d['exchange_parity'], d['e2e']['value'])
print('weak', d['weak']['value'], d['weak']['ms_per_step'], d['weak']['ms_per_rank'])
for s in d['secondary']:
    print(s['name'], s.get('error'), s.get('value'), s.get('ms_per_step'), (s.get('config') or {}).get('adaptations'))
PY
CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/r2_chi2_sweep.py planck trsm > gpurun_out/chi2_sweep_trsm.jsonl 2> gpurun_out/chi2_sweep_trsm.err
python - <<'PY'
import json
for ln in open('gpurun_out/chi2_sweep_trsm.jsonl'):
    d=json.loads(ln); print(d.get('W'), d.get('variant'), d.get('ms'), d.get('frac'), d.get('bitwise'), d.get('error'))
PY
