#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/dbg_counts.py 200000 2>&1 | tail -30 | tee gpurun_out/c2_dbg.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c2_pytest.log
tail -15 gpurun_out/c2_pytest.log
timeout 600 python tools/variant_probe.py default tp0 r01 ipt8 2>&1 | tee gpurun_out/c2_variants.log
timeout 300 python tools/perf_probe.py 4000000 2>&1 | head -3 | tee gpurun_out/c2_perf.log
RK_LIB=rakau_b200/lib/variants/librakau_b200_ipt8.so timeout 300 python tools/perf_probe.py 4000000 2>&1 | head -3 | tee -a gpurun_out/c2_perf.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; echo "bench rc=$?"; cat gpurun_out/c2_bench.json; tail -3 gpurun_out/c2_bench.err
