"""Probe (run under torchrun, 2+ GPUs): is CUDA peer memory through torch symmetric memory usable on this box, and
how fast are copy-engine pushes into every peer's buffer while an SM-filling kernel runs?"""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 64 * 1024 * 1024  # 256 MB of float32
try:
    buf = symm_mem.empty(n, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    print(rank, "rendezvous ok: world", hdl.world_size, "multicast", hdl.has_multicast_support, flush=True)
except Exception as e:  # noqa
    print(rank, "symmetric memory unavailable:", repr(e)[:300], flush=True)
    dist.destroy_process_group(); sys.exit(0)
per = n // world
src = torch.full((per,), float(rank + 1), device=dev)
peers = [hdl.get_buffer(r, (n,), torch.float32) for r in range(world)]
side = torch.cuda.Stream()
for it in range(3):
    buf.zero_(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    busy = torch.randn(8192, 8192, device=dev)
    with torch.cuda.stream(side):
        e0.record(side)
        for r in range(world):
            peers[r][rank * per:(rank + 1) * per].copy_(src, non_blocking=True)
        e1.record(side)
    for _ in range(4):
        busy = busy @ busy * 1e-4  # SM-filling work on the main stream meanwhile
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ok = all(bool((buf[r * per:(r + 1) * per] == r + 1).all()) for r in range(world))
    ms = e0.elapsed_time(e1)
    print(rank, f"push {world} x {per * 4 / 1e6:.0f} MB: {ms:.3f} ms = {world * per * 4 / ms / 1e6:.0f} GB/s egress, data ok: {ok}", flush=True)
dist.destroy_process_group()
