"""Where the end-to-end time of one evaluation with pinned host buffers goes (4M Plummer): wall clock of the two C
calls against the GPU-side times the library reports."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rakau_b200 as rk
n = 4_000_000
h = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(4)]
hn = [t.numpy() for t in h]
rk.plummer(n, 0, n, fp=32, out=[hn[3], hn[0], hn[1], hn[2]])
ho = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(3)]
hon = [t.numpy() for t in ho]
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
t = rk.Octree()
for zc in (0, 1, 0, 1):
  t.set_option("zero_copy_out", zc)
  rows = []
  for it in range(12):
      flush.zero_(); torch.cuda.synchronize()
      t0 = time.perf_counter()
      bi = t.build(hn[0], hn[1], hn[2], hn[3], where=rk.RK_HOST)
      t1 = time.perf_counter()
      t.acc_pot(0, 0.75, out=hon, where=rk.RK_HOST)
      t2 = time.perf_counter()
      ei = t.eval_info
      rows.append(dict(wall_build=(t1 - t0) * 1e3, wall_eval=(t2 - t1) * 1e3, gpu_build=bi.ms_total, enc=bi.ms_encode, sort=bi.ms_sort,
                       perm=bi.ms_permute, topo=bi.ms_topology, props=bi.ms_props, eval_total=ei.ms_total, eval_kernel=ei.ms_kernel))
  rows = rows[4:]
  print("zero_copy_out", zc, json.dumps({k: round(float(np.median([r[k] for r in rows])), 3) for k in rows[0]}))
  ref = [a.copy() for a in hon] if zc == 0 else ref
  print("  equal to the copied results:", all((a == b).all() for a, b in zip(ref, hon)))
