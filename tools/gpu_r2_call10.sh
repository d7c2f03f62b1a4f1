#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_leapfrog.py tests/test_cpp_api.py -x -q -m gpu > gpurun_out/c10_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c10_pytest.log
tail -25 gpurun_out/c10_pytest.log
