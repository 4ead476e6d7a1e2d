#!/bin/bash
# sanity of the multi-GPU paths on the final build: 2-GPU tests + the 2-GPU bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_dist.py -m gpu -q > gpurun_out/r2_pytest_gpu_multi.log 2>&1; echo "pytest multi rc=$?"; tail -2 gpurun_out/r2_pytest_gpu_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
print('n2', d['ms_per_step'], d['value'], d['e2e']['value'], d.get('exchange_parity'))
print('   weak', d['weak']['ms_per_step'], d['weak']['value'])
for s in d['secondary']: print('  ', s['name'], s.get('error'), s.get('value'), s.get('ms_per_step'), (s.get('roofline') or {}).get('frac'))
PY
