#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_leapfrog.py tests/test_gpu_update.py -x -q > gpurun_out/c8_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c8_pytest.log
tail -25 gpurun_out/c8_pytest.log
timeout 600 python -c "
import json, rakau_b200 as rk
print(json.dumps(rk.leapfrog_benchmark(16000000, 10), indent=1))
" 2>&1 | tee gpurun_out/c8_leapfrog.log
