#!/bin/bash
# final 2-GPU evidence of round 2: the multi-GPU tests, the bench line at N = 1 and N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_dist.py -m gpu -q > gpurun_out/r2_pytest_gpu_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_multi.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_n1','r2_bench_n2'):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e'].get('repetitions'), d['roofline']['frac'], d.get('exchange_parity'), d['roofline'].get('traffic'))
    if d.get('weak'): print('   weak', d['weak']['ms_per_step'], d['weak']['value'])
    for s in d['secondary']: print('  ', s['name'], s.get('error'), s.get('value'), s.get('ms_per_step'), (s.get('roofline') or {}).get('frac'))
PY
