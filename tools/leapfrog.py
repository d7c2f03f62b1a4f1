#!/usr/bin/env python
"""BASELINE config 4: kick-drift-kick leapfrog of the reference's benchmark_leapfrog.cpp:286-384 with the tree
resident on the GPU (per step: kick, drift + update_particles (full rebuild), traversal, re-index velocities
through last_perm, kick). Positions never visit the host. Prints one JSON line.

  python tools/leapfrog.py [--nparts 16000000] [--steps 10] [--dt 1e-4] [--track]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nparts", type=int, default=16_000_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--dt", type=float, default=1e-4)
    ap.add_argument("--theta", type=float, default=0.75)
    ap.add_argument("--max-leaf-n", type=int, default=16)
    ap.add_argument("--ncrit", type=int, default=128)
    ap.add_argument("--track", action="store_true", help="track total energy with accs_pots (Q = 2)")
    a = ap.parse_args()
    import torch
    import rakau_b200 as rk
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    t0 = time.time()
    x, y, z, vx, vy, vz = rk.plummer_leapfrog(a.nparts)
    n = x.size
    eps = 0.45 * float(n) ** -0.73  # Athanassoula heuristic, benchmark_leapfrog.cpp:222
    m = np.full(n, 1.0 / n, dtype=np.float32)
    t_gen = time.time() - t0
    stream = torch.cuda.current_stream()
    tree = rk.Octree(fp=32, mac="bh", device=0)
    tree.set_stream(stream.cuda_stream)
    tree.build(x, y, z, m, max_leaf_n=a.max_leaf_n, ncrit=a.ncrit)
    # device state in the tree's Morton order
    perm = torch.from_numpy(tree.perm(rk.RK_PERM).astype(np.int64)).to(dev)
    v = [torch.from_numpy(c).to(dev)[perm] for c in (vx, vy, vz)]
    pos = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(3)]
    mass = torch.empty(n, dtype=torch.float32, device=dev)
    nres = 4 if a.track else 3
    acc = [torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(nres)]
    lp = torch.empty(n, dtype=torch.int32, device=dev)
    Q = 2 if a.track else 0

    def evaluate():
        tree.acc_pot(Q, a.theta, eps=eps, out=acc, where=rk.RK_DEVICE)

    def energy():
        kin = 0.5 * float((mass * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2])).double().sum())
        pot = 0.5 * float(acc[3].double().sum())
        return kin, pot

    evaluate()
    tree.parts_device(pos[0], pos[1], pos[2], mass)
    e0 = energy() if a.track else None
    half = 0.5 * a.dt
    evs = []
    torch.cuda.synchronize()
    for s in range(a.steps):
        e_a, e_b, e_c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e_a.record()
        for j in range(3):
            v[j].add_(acc[j], alpha=half)          # kick, benchmark_leapfrog.cpp:349-356
            pos[j].add_(v[j], alpha=a.dt)           # drift, 359-369
        bi = tree.update_positions(pos[0], pos[1], pos[2], where=rk.RK_DEVICE)  # sync(): full rebuild
        e_b.record()
        evaluate()                                   # 372
        tree.perm_device(lp, rk.RK_LAST_PERM)
        idx = lp.long()
        for j in range(3):
            v[j] = v[j][idx].add_(acc[j], alpha=half)   # re-index through last_perm + second kick, 375-383
        tree.parts_device(pos[0], pos[1], pos[2], mass)
        e_c.record()
        evs.append((e_a, e_b, e_c, tree.eval_info.asdict(), bi.asdict()))
    torch.cuda.synchronize()
    ms_step = [ea.elapsed_time(ec) for ea, _, ec, _, _ in evs]
    ms_reb = [ea.elapsed_time(eb) for ea, eb, _, _, _ in evs]
    out = {
        "workload": f"plummer_leapfrog_{a.nparts}_clipped_{n}_fp32_theta{a.theta}", "nparts": n, "steps": a.steps,
        "dt": a.dt, "eps": eps, "track_integrals": a.track,
        "ms_per_step": float(np.mean(ms_step)), "ms_per_step_min": float(np.min(ms_step)),
        "ms_per_step_median": float(np.median(ms_step)), "ms_steps": [round(float(v), 3) for v in ms_step],
        "ms_kick_drift_rebuild": float(np.mean(ms_reb)),
        "ms_traverse_kernel": float(np.mean([e[3]["ms_kernel"] for e in evs])),
        "ms_rebuild_device": float(np.mean([e[4]["ms_total"] for e in evs])),
        "ginteractions_per_s": float(np.mean([e[3]["interactions"] for e in evs]) / np.mean(ms_step) / 1e6),
        "interactions_per_step": float(np.mean([e[3]["interactions"] for e in evs])),
        "n_nodes": evs[-1][4]["n_nodes"], "ic_generation_s": t_gen,
    }
    if a.track:
        e1 = energy()
        out["energy_initial"] = {"kin": e0[0], "pot": e0[1], "tot": e0[0] + e0[1]}
        out["energy_final"] = {"kin": e1[0], "pot": e1[1], "tot": e1[0] + e1[1]}
        out["rel_energy_drift"] = abs((e1[0] + e1[1]) - (e0[0] + e0[1])) / abs(e0[0] + e0[1])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
