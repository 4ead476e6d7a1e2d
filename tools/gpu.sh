#!/bin/bash
# usage: tools/gpu.sh <log> <timeout> [--gpus N] -- <command...>   : gpurun with retries while the pod answers busy (exit 3)
log=$1; shift; to=$1; shift
extra=""
if [ "$1" = "--gpus" ]; then extra="--gpus $2"; shift; shift; fi
[ "$1" = "--" ] && shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to $extra -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (attempt $i)" >> "$log"; exit $rc; fi
  sleep 90
done
echo "gave up: pod busy" >> "$log"; exit 3
