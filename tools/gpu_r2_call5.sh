#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/variant_probe.py default nosteal ns_c4 ns_c4np4 ns_c4u8 ns_u2 ns_u8 2>&1 | tee gpurun_out/c5_variants.log
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c5_pytest.log
tail -5 gpurun_out/c5_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; echo "bench rc=$?"; cat gpurun_out/c5_bench.json; tail -3 gpurun_out/c5_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c5_bench_ref.json 2>> gpurun_out/c5_bench.err; cat gpurun_out/c5_bench_ref.json
