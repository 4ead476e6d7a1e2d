#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=25 --deselect tests/test_gpu_multi.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for s in 200; do
timeout 600 python bench.py --steps $s --warmup 5 --secondary none --no-cpu-baseline > gpurun_out/bench_n1_s$s.json 2> gpurun_out/bench_n1_s$s.err
echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_n1_s$s.json'));print(d['ms_per_step'],d['value'],d['e2e']['value'],d['ms_per_step_instrumented_pass'],d['roofline']['other_stages_ms'])"
done
