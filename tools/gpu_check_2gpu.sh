#!/bin/bash
# 2-GPU check: build + multi-GPU tests, then a 32 M bench with the hardware parity check: gpurun --gpus 2 --timeout 1500 -- "bash tools/gpu_check_2gpu.sh"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/n2_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/n2_pytest.log
tail -15 gpurun_out/n2_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --nparts 32000000 > gpurun_out/n2_bench2.json 2> gpurun_out/n2_bench2.err; echo "bench2 rc=$?"; tail -3 gpurun_out/n2_bench2.err
python - <<'PY'
import json
for ln in open('gpurun_out/n2_bench2.json'):
    if ln.startswith('{'):
        d=json.loads(ln)
        print(d['tree'].get('output_exchange'), d['tree'].get('codes_gather'), d['parity_checked'], d['ms_per_step'], d['ms_build'], d['ms_traverse_and_exchange'], d['build_phases_rank0_ms'], d['ms_traverse_kernel_per_rank'], d['e2e']['ms_per_step'])
PY
