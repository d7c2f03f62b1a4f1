#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c4_pytest.log
tail -12 gpurun_out/c4_pytest.log
timeout 600 python tools/variant_probe.py default nosteal r01 2>&1 | tee gpurun_out/c4_variants.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "bench rc=$?"; cat gpurun_out/c4_bench.json; tail -3 gpurun_out/c4_bench.err
