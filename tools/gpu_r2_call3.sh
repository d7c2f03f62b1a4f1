#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c3_pytest.log
tail -12 gpurun_out/c3_pytest.log
RK_DEBUG_LAUNCH=1 timeout 600 python tools/variant_probe.py default tp0 r01 2>&1 | tee gpurun_out/c3_variants.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; echo "bench rc=$?"; cat gpurun_out/c3_bench.json; tail -3 gpurun_out/c3_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traverse_kernel -s 2 -c 1 -o gpurun_out/r02_prof_traverse -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c3_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
