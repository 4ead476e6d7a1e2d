#!/bin/bash
# where does the exchange time go?  2 GPUs: per-GPU compute alone (N=1) vs the sharded step with parts of the
# exchange switched off (CSL_XCHG_DEBUG: timing only, results wrong)
mkdir -p gpurun_out
B="--steps 50 --warmup 5 --secondary none --no-cpu-baseline --no-weak"
for w in 8192 32768; do
  timeout 300 python bench.py $B --walkers $w > gpurun_out/x_n1_w$w.json 2> gpurun_out/x_n1_w$w.err
  python -c "
import json; d=json.load(open('gpurun_out/x_n1_w$w.json')); print('N=1 walkers/GPU $w', d['ms_per_step'])"
done
for w in 16384 65536; do
  for dbg in 0 1 2 4 7; do
    CSL_XCHG_DEBUG=$dbg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $B --walkers $w > gpurun_out/x_n2_w${w}_d$dbg.json 2> gpurun_out/x_n2_w${w}_d$dbg.err
    python -c "
import json
try:
    d=json.load(open('gpurun_out/x_n2_w${w}_d$dbg.json')); print('N=2 total $w debug $dbg', d['ms_per_step'], d['ms_per_rank'], d.get('exchange_parity'))
except Exception as e: print('N=2 total $w debug $dbg FAILED', e)"
  done
done
