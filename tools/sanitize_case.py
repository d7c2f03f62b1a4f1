"""A small workload that launches every kernel of librakau_b200.so, meant to be run under compute-sanitizer
(tools/gpu_sanitize.sh). Sizes are small because the sanitizer slows kernels down by 10-100 x; every run of the tree is
inside the work-stealing tail at these sizes, and RK_N can be raised to cover the big-tile paths."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rakau_b200 as rk  # noqa: E402

n = int(os.environ.get("RK_N", "30000"))
which = os.environ.get("RK_CASE", "all")


def case_eval(fp, mac):
    F = np.float32 if fp == 32 else np.float64
    m, x, y, z = rk.plummer(n, fp=fp)
    t = rk.Octree(fp=fp, mac=mac)
    t.build(x, y, z, m)
    for Q in (0, 1, 2):
        t.acc_pot(Q, 0.75, G=1.5, eps=0.0 if Q == 0 else 0.01)
    t.acc_pot(2, 0.5, ordered=True)
    t.exact(5)
    t.update_positions(x * F(1.01), y, z)
    t.acc_pot(0, 0.75)
    t.update_masses(m * F(2))
    t.acc_pot(1, 0.75)
    t.nodes(), t.crit(), t.codes(), t.perm(), t.digest(), t.crit_lower_bound([0, n // 2, n - 1])
    c = t.clone()
    c.acc_pot(0, 0.75)
    t.set_option("props_bottom_up", 1)
    t.update_positions(x, y, z)
    t.acc_pot(0, 0.75)
    t.close(), c.close()


def case_leapfrog():
    x, y, z, vx, vy, vz = rk.plummer_leapfrog(n)
    m = np.full(x.size, 1.0 / x.size, dtype=np.float32)
    for track in (False, True):
        t = rk.Octree()
        t.build(x, y, z, m)
        t.leapfrog_init(vx, vy, vz, 0.75, eps=1e-3, track_integrals=track)
        for _ in range(3):
            t.leapfrog_step(1e-3)
        t.leapfrog_get(0), t.leapfrog_get(1)
        t.close()


def case_shard():
    import torch

    m, x, y, z = rk.plummer(n)
    d = [torch.from_numpy(a).cuda() for a in (x, y, z, m)]
    t = rk.Octree()
    t.build(x, y, z, m)
    box = t.box_size
    s = rk.Octree()
    s.encode_shard(*d, n, box)
    codes = torch.empty(n, dtype=torch.int64, device="cuda")
    s.codes_device(codes)
    spl = torch.sort(codes).values[n // 4::n // 4][:3].contiguous()
    s.partition_shard(spl)
    s.sort_shard(*d, n, box)
    cs = torch.empty(n, dtype=torch.int64, device="cuda")
    s.codes_device(cs)
    ps = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(4)]
    s.parts_device(*ps)
    pm = torch.empty(n, dtype=torch.int32, device="cuda")
    s.perm_device(pm)
    u = rk.Octree()
    u.build_presorted(*ps, cs, pm, n, box)
    u.acc_pot(0, 0.75)
    print("presorted digests equal:", bool((u.digest() == t.digest()).all()))
    s.close(), t.close(), u.close()


def case_external():
    m, x, y, z = rk.plummer(n)
    t = rk.Octree()
    t.build(x, y, z, m)
    rk.traverse_external_tree(t.nodes(), t.parts(), t.codes(), 0, 1.0 / 0.75 ** 2)
    t.close()


cases = {"f32bh": lambda: case_eval(32, "bh"), "f32geom": lambda: case_eval(32, "bh_geom"),
         "f64bh": lambda: case_eval(64, "bh"), "leapfrog": case_leapfrog, "shard": case_shard,
         "external": case_external}
for name, fn in cases.items():
    if which in ("all", name):
        fn()
        print("case", name, "ok", flush=True)
print("launches", rk.kernel_launch_count())
