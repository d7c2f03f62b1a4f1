"""Developer probe: build + traversal timings at large N (no oracle). Run under gpurun."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rakau_b200 as _rk
import rakau_b200 as rk

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
cfgs = [(16, 128), (16, 64), (16, 256), (32, 128), (8, 64), (32, 256), (64, 256)]
t0 = time.time(); m, x, y, z = _rk.plummer(N); print("gen", time.time() - t0, flush=True)
g = rk.Octree()
for mln, nc in cfgs:
    for it in range(3):
        bi = g.build(x, y, z, m, max_leaf_n=mln, ncrit=nc)
    out = None
    for it in range(3):
        out = g.acc_pot(0, 0.75)
    ei = g.eval_info.asdict(); b = bi.asdict()
    inter = ei["interactions"]
    print(f"mln={mln} nc={nc}: nodes {b['n_nodes']} crit {b['n_crit']} maxg {b['max_group']} build {b['ms_total']:.3f} ms (enc {b['ms_encode']:.3f} sort {b['ms_sort']:.3f} perm {b['ms_permute']:.3f} topo {b['ms_topology']:.3f} props {b['ms_props']:.3f}) | trav kernel {ei['ms_kernel']:.3f} ms total {ei['ms_total']:.3f} | inter {inter/1e9:.3f}G mac {ei['mac_tests']/1e6:.1f}M acc {ei['accepted']/1e6:.1f}M p2p {ei['p2p_pairs']/1e9:.3f}G self {ei['self_pairs']/1e9:.4f}G | {inter/ei['ms_kernel']/1e6:.1f} Ginter/s", flush=True)
