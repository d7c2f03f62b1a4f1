#!/bin/bash
# One gpurun call that produces the evidence kept under profiles/: GPU tests, bench lines of both arms, the ncu launch
# list of a bench step and full captures of the traversal kernel and of the build kernels.
#   gpurun --timeout 2400 -- 'bash tools/gpu_profile_round.sh r02'
tag=${1:-rXX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${tag}_bench_1gpu.json; tail -3 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err; cut -c1-300 gpurun_out/${tag}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:traverse_kernel -s 2 -c 1 -o gpurun_out/${tag}_prof_traverse -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/${tag}_build_kernels.csv python tools/prof_build.py > gpurun_out/${tag}_prof_build.log 2>&1
ls -la gpurun_out | tail -12
