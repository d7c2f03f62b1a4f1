"""Developer smoke: GPU build + traversal vs the oracle at a few sizes, with timings. Run under gpurun."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle
from oracle import OracleTree
import rakau_b200 as rk

def check(N, mln=16, nc=128, fp=32, mac="bh", theta=0.75, Q=0, eps=0.0, G=1.0):
    m, x, y, z = oracle.plummer(N, fp=fp)
    t0 = time.time(); o = OracleTree(x, y, z, m, max_leaf_n=mln, ncrit=nc, fp=fp, mac=mac); tob = time.time() - t0
    g = rk.Octree(fp=fp, mac=mac)
    bi = g.build(x, y, z, m, max_leaf_n=mln, ncrit=nc)
    print(f"N={N} fp={fp} mac={mac} mln={mln} nc={nc}: box {g.box_size} vs {o.box_size}; nodes {g.nnodes} vs {len(o.nodes())}; crit {g.ncrit_nodes} vs {len(o.crit()[0])}; build {bi.asdict()}")
    ok = g.box_size == o.box_size
    ok &= bool((g.codes() == o.codes()).all()); print("  codes", ok)
    for w in range(3):
        e = bool((g.perm(w) == o.perm(w)).all()); ok &= e
        if not e: print("  perm", w, "MISMATCH")
    gp, op = g.parts(), o.parts()
    for a, b in zip(gp, op): ok &= bool((a == b).all())
    print("  parts/perm ok", ok)
    gn, on = g.nodes(), o.nodes()
    if len(gn) == len(on):
        for f in ("begin", "end", "n_children", "code", "level"):
            e = bool((gn[f] == on[f]).all()); ok &= e
            if not e:
                bad = np.nonzero(gn[f] != on[f])[0]; print("  node field", f, "MISMATCH at", bad[:5], gn[f][bad[:5]], on[f][bad[:5]])
        rel = np.abs(gn["props"].astype(np.float64) - on["props"]) / np.maximum(np.abs(on["props"]), 1e-30)
        print("  props max rel err", rel.max(axis=0), "dim eq", bool((gn["dim"] == on["dim"]).all()), "delta maxabs", np.abs(gn["delta"]-on["delta"]).max())
    else:
        ok = False
    gc, (oc, _) = g.crit(), o.crit()
    e = gc.shape == oc.shape and bool((gc == oc).all()); ok &= e; print("  crit ok", e)
    # traversal
    t0 = time.time(); oo, cnt = o.acc_pot(Q, theta, G=G, eps=eps, nthreads=8); toa = time.time() - t0
    go = g.acc_pot(Q, theta, G=G, eps=eps)
    ei = g.eval_info.asdict()
    print("  eval", ei, "oracle", cnt, f"oracle time {toa:.2f}s")
    for k in ("mac_tests", "accepted", "p2p_pairs", "self_pairs", "interactions"):
        if ei[k] != cnt[k]: print("   COUNT MISMATCH", k, ei[k], cnt[k]); ok = False
    if Q != 1:
        ga = np.stack(go[:3], 1).astype(np.float64); oa = np.stack(oo[:3], 1).astype(np.float64)
        rel = np.linalg.norm(ga - oa, axis=1) / np.linalg.norm(oa, axis=1)
        print(f"  acc rel err vs oracle: median {np.median(rel):.3e} max {rel.max():.3e} nan {np.isnan(ga).sum()}")
    if Q != 0:
        gpz = go[-1].astype(np.float64); opz = oo[-1].astype(np.float64)
        rel = np.abs(gpz - opz) / np.abs(opz)
        print(f"  pot rel err vs oracle: median {np.median(rel):.3e} max {rel.max():.3e}")
    ex = g.exact(5); oe = o.exact(5)
    print("  exact", ex, oe)
    print("  OK" if ok else "  FAILED")
    return ok

if __name__ == "__main__":
    print("devices", rk.device_count())
    allok = True
    allok &= check(1000)
    allok &= check(1000, mln=1, nc=1)
    allok &= check(20000, mln=8, nc=16, Q=2, eps=0.01, G=2.5)
    allok &= check(20000, fp=64, theta=0.5)
    allok &= check(20000, mac="bh_geom", Q=1)
    allok &= check(300000)
    print("ALL OK" if allok else "SOME FAILED")
