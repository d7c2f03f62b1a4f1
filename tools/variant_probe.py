"""Time the traversal kernel of several tuning builds (tools/build_variants.sh) on the 4M Plummer workload.
usage: python tools/variant_probe.py name1 name2 ...   (name 'default' = the product library)"""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, json
import numpy as np, torch
sys.path.insert(0, %r)
import rakau_b200 as rk
n = 4_000_000
h = [np.empty(n, dtype=np.float32) for _ in range(4)]
rk.plummer(n, 0, n, fp=32, out=[h[3], h[0], h[1], h[2]])
d = [torch.from_numpy(a).cuda() for a in h]
out = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(4)]
t = rk.Octree(); t.set_stream(0)
res = {}
for (nc, Q, theta) in ((128, 0, 0.75), (256, 0, 0.75), (128, 2, 0.75)):
    t.build(*d, where=rk.RK_DEVICE, n=n, ncrit=nc)
    ks = []
    for it in range(7):
        t.acc_pot(Q, theta, out=out[:{0: 3, 1: 1, 2: 4}[Q]], where=rk.RK_DEVICE)
        ks.append(t.eval_info.ms_kernel)
    ks = sorted(ks[2:])
    res["nc%%d_Q%%d" %% (nc, Q)] = (round(ks[len(ks) // 2], 3), round(t.eval_info.interactions / ks[len(ks) // 2] / 1e6, 1))
res["chk"] = float(out[0].double().abs().sum().item())
print(json.dumps(res))
''' % ROOT
for name in sys.argv[1:]:
    env = dict(os.environ)
    if name != "default":
        env["RK_LIB"] = os.path.join(ROOT, "rakau_b200", "lib", "variants", "librakau_b200_%s.so" % name)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
    print(name, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:], flush=True)
