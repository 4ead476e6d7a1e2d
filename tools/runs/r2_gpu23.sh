#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_xtra_params.py -m gpu -q > gpurun_out/pytest_mspec.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_mspec.log
