#!/bin/bash
# round-2 GPU check 1: full GPU test suite + one bench line (1 GPU)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -x --deselect tests/test_gpu_multi.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"
tail -c 600 gpurun_out/bench_n1.err
head -c 1500 gpurun_out/bench_n1.json
