#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c7_pytest.log
tail -12 gpurun_out/c7_pytest.log
for v in default tt512 tt1024; do
  if [ $v != default ]; then export RK_LIB=rakau_b200/lib/variants/librakau_b200_$v.so; else unset RK_LIB; fi
  echo "== $v"; timeout 300 python tools/perf_probe.py 4000000 2>&1 | head -2 | tail -1 | cut -c1-220
done 2>&1 | tee gpurun_out/c7_perf.log
unset RK_LIB
timeout 300 python tools/perf_probe.py 32000000 2>&1 | head -2 | tail -1 | cut -c1-220 | tee -a gpurun_out/c7_perf.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/c7_bench.json; tail -3 gpurun_out/c7_bench.err
