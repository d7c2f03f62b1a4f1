mkdir -p gpurun_out
RK_DEBUG_BARRIER=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/n4_bench4.json 2> gpurun_out/n4_bench4.err; echo "bench4 rc=$?"; tail -2 gpurun_out/n4_bench4.err
python - <<'PY'
import json
for ln in open('gpurun_out/n4_bench4.json'):
    if ln.startswith('{'):
        d=json.loads(ln)
        print(d['tree'].get('output_exchange'), d['parity_checked'], d['ms_per_step'], d['ms_build'], d['ms_traverse_and_exchange'], d['e2e']['ms_per_step'], d['ms_traverse_kernel_per_rank'])
PY
