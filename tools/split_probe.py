"""Split evaluation (walk launch + evaluation launch, rk_tree_set_option("split_lists", 1)) against the fused kernel:
bit equality of the results and kernel times on the 4M Plummer workload and a few other shapes."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rakau_b200 as rk

def run(n, Q, theta, ncrit=128, eps=0.0, ordered=False, host=False, reps=5):
    h = [np.empty(n, dtype=np.float32) for _ in range(4)]
    rk.plummer(n, 0, n, fp=32, out=[h[3], h[0], h[1], h[2]])
    d = [torch.from_numpy(a).cuda() for a in h]
    nres = {0: 3, 1: 1, 2: 4}[Q]
    t = rk.Octree(); t.set_stream(0)
    t.build(*d, where=rk.RK_DEVICE, n=n, ncrit=ncrit)
    res = {}
    outs = {}
    for mode in (0, 1):
        t.set_option("split_lists", mode)
        ks = []
        for it in range(reps):
            if host:
                o = t.acc_pot(Q, theta, eps=eps, ordered=ordered)
            else:
                o = [torch.zeros(n, dtype=torch.float32, device="cuda") for _ in range(nres)]
                t.acc_pot(Q, theta, eps=eps, ordered=ordered, out=o, where=rk.RK_DEVICE)
                torch.cuda.synchronize()
            ks.append(t.eval_info.ms_kernel)
        outs[mode] = [np.asarray(a.cpu()) if hasattr(a, "cpu") else a for a in o]
        res["ms_%d" % mode] = [round(k, 3) for k in ks]
        res["inter_%d" % mode] = int(t.eval_info.interactions)
        res["kernel_%d" % mode] = t.last_kernel()
    res["bit_equal"] = all((a.view(np.uint32) == b.view(np.uint32)).all() for a, b in zip(outs[0], outs[1]))
    if not res["bit_equal"]:
        res["maxdiff"] = max(float(np.abs(a - b).max()) for a, b in zip(outs[0], outs[1]))
        res["ndiff"] = int(sum((a.view(np.uint32) != b.view(np.uint32)).sum() for a, b in zip(outs[0], outs[1])))
    return res

cases = [dict(n=200000, Q=0, theta=0.75), dict(n=4000000, Q=0, theta=0.75), dict(n=4000000, Q=2, theta=0.75, eps=0.01),
         dict(n=1000000, Q=1, theta=0.4), dict(n=1000000, Q=0, theta=0.75, ordered=True),
         dict(n=4000000, Q=0, theta=0.75, host=True), dict(n=4000000, Q=0, theta=0.75, ncrit=256)]
sel = sys.argv[1:] 
for i, c in enumerate(cases):
    if sel and str(i) not in sel:
        continue
    try:
        print(i, c, json.dumps(run(**c)), flush=True)
    except Exception as e:
        print(i, c, "FAILED", repr(e), flush=True)
