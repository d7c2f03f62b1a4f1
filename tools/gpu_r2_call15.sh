#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c15_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c15_pytest.log
tail -6 gpurun_out/c15_pytest.log
for v in default lb1 lb16 lb8i8; do
  if [ $v != default ]; then export RK_LIB=rakau_b200/lib/variants/librakau_b200_$v.so; else unset RK_LIB; fi
  echo "== $v"; timeout 300 python tools/perf_probe.py 4000000 2>&1 | head -2 | tail -1 | cut -c1-200
  timeout 300 python tools/perf_probe.py 16000000 2>&1 | head -2 | tail -1 | cut -c1-200
done 2>&1 | tee gpurun_out/c15_perf.log
unset RK_LIB
timeout 600 python tools/variant_probe.py default c4b128 2>&1 | tee gpurun_out/c15_variants.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"onesweep_kernel|topo_emit_kernel|window_fused_kernel|topo_finalize" -s 12 -c 6 -o gpurun_out/r02_prof_build -f python tools/prof_build.py > gpurun_out/c15_ncu_build.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c15_bench.json 2> gpurun_out/c15_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/c15_bench.json; tail -3 gpurun_out/c15_bench.err
