"""Collective bandwidth probe (run under torchrun): what NCCL delivers on this box for the exchanges of the
multi-GPU path."""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
def bench(name, fn, nbytes, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    if rank == 0: print(f"{name:40s} {ms:8.3f} ms  {nbytes/ms/1e6:8.1f} GB/s (bytes received per rank / time)", flush=True)
n = 128_000_000
per = n // world
full = torch.empty(n, dtype=torch.float32, device=dev)
mine = torch.randn(per, dtype=torch.float32, device=dev)
bench("all_gather_into_tensor 512MB total", lambda: dist.all_gather_into_tensor(full, mine), 4 * n * (world - 1) / world)
bench("all_reduce 512MB", lambda: dist.all_reduce(full), 4 * n)
def p2p():
    ops = []
    for r in range(world):
        if r != rank:
            ops.append(dist.P2POp(dist.isend, mine, r)); ops.append(dist.P2POp(dist.irecv, full[r*per:(r+1)*per], r))
    for w in dist.batch_isend_irecv(ops): w.wait()
bench("batch_isend_irecv all-gather 512MB", p2p, 4 * n * (world - 1) / world)
outl = [full[r*per:(r+1)*per] for r in range(world)]
bench("all_gather(list) 512MB", lambda: dist.all_gather(outl, mine), 4 * n * (world - 1) / world)
big = torch.empty(n * 6, dtype=torch.float32, device=dev); mb = torch.randn(per * 6, dtype=torch.float32, device=dev)
bench("all_gather_into_tensor 3GB total", lambda: dist.all_gather_into_tensor(big, mb), 24 * n * (world - 1) / world, reps=3)
a2o = torch.empty(per, dtype=torch.float32, device=dev)
bench("all_to_all_single 64MB per rank", lambda: dist.all_to_all_single(a2o, mine), 4 * per * (world - 1) / world)
dist.destroy_process_group()
