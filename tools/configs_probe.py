"""Kernel/build timings for BASELINE configs 1-3 on one GPU (4M Plummer): fp32 accs, fp32 accs+pots eps/G, fp64 theta 0.5."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rakau_b200 as rk
n = 4_000_000
res = {}
for name, fp, Q, theta, eps, G, mac in (("config1_fp32_accs", 32, 0, 0.75, 0.0, 1.0, "bh"), ("config2_fp32_accs_pots_eps_G", 32, 2, 0.75, 0.01, 2.5, "bh"),
                                        ("config3_fp64_accs_theta0.5", 64, 0, 0.5, 0.0, 1.0, "bh"), ("fp32_accs_bh_geom", 32, 0, 0.75, 0.0, 1.0, "bh_geom"),
                                        ("fp32_pots", 32, 1, 0.75, 0.0, 1.0, "bh")):
    if len(sys.argv) > 1 and not any(a in name for a in sys.argv[1:]):
        continue
    dt = np.float32 if fp == 32 else np.float64
    h = [np.empty(n, dtype=dt) for _ in range(4)]
    rk.plummer(n, 0, n, fp=fp, out=[h[3], h[0], h[1], h[2]])
    d = [torch.from_numpy(a).cuda() for a in h]
    nres = {0: 3, 1: 1, 2: 4}[Q]
    out = [torch.empty(n, dtype=torch.float32 if fp == 32 else torch.float64, device="cuda") for _ in range(nres)]
    t = rk.Octree(fp=fp, mac=mac); t.set_stream(0)
    bs, ks = [], []
    for it in range(6):
        bi = t.build(*d, where=rk.RK_DEVICE, n=n)
        t.acc_pot(Q, theta, G=G, eps=eps, out=out, where=rk.RK_DEVICE)
        bs.append(bi.ms_total); ks.append(t.eval_info.ms_kernel)
    b, k = sorted(bs[2:])[2], sorted(ks[2:])[2]
    inter = t.eval_info.interactions
    res[name] = {"kernel": t.last_kernel(), "ms_build": round(b, 3), "ms_traverse": round(k, 3), "ms_eval": round(b + k, 3), "interactions": inter,
                 "Ginteractions_per_s_kernel": round(inter / k / 1e6, 1)}
    print(name, res[name], flush=True)
    del t, d, out
    torch.cuda.empty_cache()
json.dump(res, open(os.path.join("gpurun_out", "configs_1gpu.json"), "w"), indent=1)
