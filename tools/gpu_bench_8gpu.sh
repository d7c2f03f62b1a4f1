#!/bin/bash
# 8-GPU run of BASELINE config 5 (128 M particles): gpurun --gpus 8 --timeout 1200 -- "bash tools/gpu_bench_8gpu.sh"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
RK_DEBUG_BARRIER=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/n8_bench8.json 2> gpurun_out/n8_bench8.err; echo "bench8 rc=$?"; tail -5 gpurun_out/n8_bench8.err
python - <<'PY'
import json
for ln in open('gpurun_out/n8_bench8.json'):
    if ln.startswith('{'):
        d=json.loads(ln)
        print('parity', d['parity_checked'], d['parity'])
        print(d['tree'].get('output_exchange'), d['tree'].get('codes_gather')); print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'build', d['ms_build'], 'trav+exch', d['ms_traverse_and_exchange'], 'e2e', d['e2e']['ms_per_step'])
        print('phases rank0', d['build_phases_rank0_ms'])
        print('kernel per rank', d['ms_traverse_kernel_per_rank'])
        print('replicated phases', d['build_phases_ms'], 'imbalance', d['tree']['shard_cost_imbalance'])
PY
