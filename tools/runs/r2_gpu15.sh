#!/bin/bash
# does pre-queueing the K steps behind a host-released gate remove the one-off cost of short timed regions?
mkdir -p gpurun_out
B="--warmup 5 --secondary none --no-cpu-baseline --no-weak"
run1() { # name, env, steps, walkers
  env $2 timeout 300 python bench.py $B --steps $3 --walkers $4 > gpurun_out/g_$1.json 2> gpurun_out/g_$1.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/g_$1.json')); print('$1', d['ms_per_step'], d['value'], d.get('launches'))
except Exception as e: print('$1 FAILED', e)"
}
run2() {
  env $2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 $B --steps $3 --walkers $4 > gpurun_out/g_$1.json 2> gpurun_out/g_$1.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/g_$1.json')); print('$1', d['ms_per_step'], d['value'], d.get('launches'), d.get('exchange_parity'))
except Exception as e: print('$1 FAILED', e)"
}
run1 n1_w8192_s20_gate "A=1" 20 8192
run1 n1_w8192_s20_nogate "CSL_BENCH_NO_GATE=1" 20 8192
run1 n1_w8192_s200 "A=1" 200 8192
run1 n1_w65536_s20_gate "A=1" 20 65536
run1 n1_w65536_s20_nogate "CSL_BENCH_NO_GATE=1" 20 65536
run2 n2_w16384_s20_gate "A=1" 20 16384
run2 n2_w16384_s20_nogate "CSL_BENCH_NO_GATE=1" 20 16384
run2 n2_w16384_s200 "A=1" 200 16384
run2 n2_w65536_s20_gate "A=1" 20 65536
run2 n2_w65536_s20_nogate "CSL_BENCH_NO_GATE=1" 20 65536
tail -3 gpurun_out/g_n2_w16384_s20_gate.err
