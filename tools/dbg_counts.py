"""Developer probe: per-group interaction counts of the CUDA traversal vs the oracle (which groups differ, where in
their run). Run under gpurun."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle, rakau_b200 as rk
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
m, x, y, z = oracle.plummer(n)
o = oracle.OracleTree(x, y, z, m)
oo, cnt, pg = o.acc_pot(0, 0.75, nthreads=8, per_group=True)
g = rk.Octree(); g.build(x, y, z, m)
go = g.acc_pot(0, 0.75)
ei = g.eval_info.asdict()
print({k: (ei[k], cnt[k]) for k in ("mac_tests", "accepted", "p2p_pairs", "self_pairs", "interactions")})
gc = g.group_costs().astype(np.int64); pg = pg.astype(np.int64)
bad = np.flatnonzero(gc != pg)
print("groups", len(gc), "differing", len(bad))
cb = g.crit()[:, 1].astype(np.int64)
for j in bad[:20]:
    print(j, "begin", cb[j], "window", cb[j] // 256, "T", (g.crit()[j, 2] - g.crit()[j, 1]), "gpu", gc[j], "oracle", pg[j])
rel = np.linalg.norm(np.stack(go, 1).astype(np.float64) - np.stack(oo, 1), axis=1) / np.linalg.norm(np.stack(oo, 1).astype(np.float64), axis=1)
print("rel err median", np.median(rel), "max", rel.max(), "q999", np.quantile(rel, 0.999))
