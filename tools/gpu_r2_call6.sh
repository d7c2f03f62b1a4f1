#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/c6_pytest_multi.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c6_pytest_multi.log
tail -25 gpurun_out/c6_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --nparts 32000000 > gpurun_out/c6_bench2.json 2> gpurun_out/c6_bench2.err; echo "bench2 rc=$?"; cat gpurun_out/c6_bench2.json | cut -c1-3000; tail -5 gpurun_out/c6_bench2.err
