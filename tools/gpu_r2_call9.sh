#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_leapfrog.py tests/test_cpp_api.py -x -q -m gpu > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c9_pytest.log
tail -25 gpurun_out/c9_pytest.log
benchmark/bin/benchmark_leapfrog --nparts 16000000 --steps 5 --device 2>&1 | tail -4 | tee gpurun_out/c9_lf.log
benchmark/bin/benchmark_leapfrog --nparts 16000000 --steps 3 2>&1 | tail -3 | tee -a gpurun_out/c9_lf.log
benchmark/bin/benchmark_acc --nparts 4000000 2>&1 | grep Elapsed | tee -a gpurun_out/c9_lf.log
