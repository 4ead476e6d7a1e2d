#!/bin/bash
# end-to-end call: polling wait (default) against the blocking cudaStreamSynchronize, started plainly and under taskset
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --secondary none --no-cpu-baseline"
run() { # name, env, prefix
  env $2 $3 timeout 300 python bench.py $B > gpurun_out/e2e_$1.json 2> gpurun_out/e2e_$1.err
  python -c "
import json
d=json.load(open('gpurun_out/e2e_$1.json')); print('$1', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['repetitions'])"
}
run poll_free "A=1" ""
run poll_taskset "A=1" "taskset -c 0-15"
run block_free "CSL_E2E_BLOCKING=1" ""
run block_taskset "CSL_E2E_BLOCKING=1" "taskset -c 0-15"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2
