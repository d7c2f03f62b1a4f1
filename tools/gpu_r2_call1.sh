#!/bin/bash
# round 2, call 1: new two-phase kernel parity + A/B timing, bottom-up props check, build-kernel byte counters
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c1_pytest.log
tail -15 gpurun_out/c1_pytest.log
timeout 600 python tools/variant_probe.py default tp0 2>&1 | tee gpurun_out/c1_variants.log
timeout 600 python tests/studies/props_bottomup_check.py 2>&1 | tee gpurun_out/c1_props.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/c1_build_kernels.csv python tools/prof_build.py > gpurun_out/c1_prof_build.log 2>&1
tail -2 gpurun_out/c1_prof_build.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench rc=$?"; cat gpurun_out/c1_bench.json; tail -3 gpurun_out/c1_bench.err
