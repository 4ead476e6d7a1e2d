#!/bin/bash
# is the spread of the end-to-end number (43 .. 50 M/s) a NUMA placement lottery?
mkdir -p gpurun_out
lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)" ; nvidia-smi topo -m 2>/dev/null | head -8
for f in /sys/bus/pci/devices/*/numa_node; do d=$(dirname $f); if [ -e $d/class ] && grep -q "^0x0302\|^0x0300" $d/class 2>/dev/null; then echo "$d numa_node=$(cat $f)"; fi; done
B="--steps 20 --warmup 5 --secondary none --no-cpu-baseline"
run() { # name, prefix
  $2 timeout 300 python bench.py $B > gpurun_out/numa_$1.json 2> gpurun_out/numa_$1.err
  python -c "
import json
d=json.load(open('gpurun_out/numa_$1.json')); print('$1', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['repetitions'], 'instrumented', d['ms_per_step_instrumented_pass'])"
}
run free ""
NODES=$(lscpu | grep -i "NUMA node(s)" | awk '{print $NF}')
for n in $(seq 0 $((NODES-1))); do
  CPUS=$(lscpu | grep -i "NUMA node$n CPU" | awk '{print $NF}')
  run node$n "taskset -c $CPUS"
done
run free2 ""
