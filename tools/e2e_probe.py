"""Where the end-to-end (host buffer) step spends its time: per-call wall clock and CUDA-event figures."""
import time
import numpy as np
import torch
import rakau_b200 as rk

n = 4_000_000
h = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(4)]
rk.plummer(n, 0, n, fp=32, out=[h[3].numpy(), h[0].numpy(), h[1].numpy(), h[2].numpy()])
hn = [t.numpy() for t in h]
ho = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(3)]
hon = [t.numpy() for t in ho]
t = rk.Octree()
t.set_stream(0)
d = [a.cuda() for a in h]
do = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(3)]
for mode in ("device", "host"):
    for it in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if mode == "host":
            bi = t.build(*hn, where=rk.RK_HOST)
        else:
            bi = t.build(*d, where=rk.RK_DEVICE, n=n)
        t1 = time.perf_counter()
        if mode == "host":
            t.acc_pot(0, 0.75, out=hon, where=rk.RK_HOST)
        else:
            t.acc_pot(0, 0.75, out=do, where=rk.RK_DEVICE)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        ei = t.eval_info
        if it >= 3:
            print(f"{mode}: build wall {1e3*(t1-t0):.3f} (events {bi.ms_total:.3f}: enc {bi.ms_encode:.3f} sort {bi.ms_sort:.3f} "
                  f"perm {bi.ms_permute:.3f} topo {bi.ms_topology:.3f} props {bi.ms_props:.3f})  acc_pot wall {1e3*(t2-t1):.3f} "
                  f"(kernel {ei.ms_kernel:.3f} total {ei.ms_total:.3f} launches {ei.kernel_launches})  sum {1e3*(t2-t0):.3f}")
# raw copy speeds
for nb, src, dst in ((64e6, torch.empty(16_000_000, dtype=torch.float32).pin_memory(), torch.empty(16_000_000, dtype=torch.float32, device="cuda")),):
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
        src.copy_(dst, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"64 MB H2D {1e3*(t1-t0):.3f} ms, D2H {1e3*(t2-t1):.3f} ms")
