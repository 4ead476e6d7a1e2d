#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/r2_trsm_debug.py > gpurun_out/trsm_debug.jsonl 2> gpurun_out/trsm_debug.err; echo "rc=$?"; tail -3 gpurun_out/trsm_debug.err
cat gpurun_out/trsm_debug.jsonl
