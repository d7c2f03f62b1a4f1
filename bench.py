#!/usr/bin/env python
"""bench.py — headline benchmark of the Barnes-Hut hot path (BASELINE.json): one "step" = one acceleration
evaluation = tree build (Morton encode, sort, octree, node properties) + ncrit-grouped traversal.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W

N = 1: BASELINE config 1 (3D Plummer, 4M particles, fp32, theta = 0.75, accelerations, max_leaf_n 16,
ncrit 128 — the parameters behind the reference's published table). N > 1: BASELINE config 5 (128M
particles, strong scaling): inputs sharded over the ranks, all-gathered with NCCL, identical tree on every
GPU, critical nodes sharded by cost-weighted Morton ranges, outputs all-gathered.

Prints ONE JSON line (rank 0). `value` = interactions / device time with inputs resident in HBM;
`e2e` = the same through the C ABI with pinned HOST buffers (H2D + D2H inside the timed region).
`--impl reference` times the reference's CPU algorithm on the host cores (oracle/_ref when it was built,
else the scalar oracle port) — the only place besides `cpu_baseline` where oracle/ is executed.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Ginteractions/s (accel eval = tree build + traversal, Plummer, theta=0.75, fp32; ms per accel eval in ms_per_step)"
SLOTS_PER_INTERACTION = 12  # FP32 issue slots per pair (3 FADD + 3 FFMA + 3 FMUL + 3 FFMA), SURVEY §8(d) conv. B
FLOP_PER_INTERACTION_LIT = 20  # literature convention A


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nparts", type=int, default=0, help="0 = 4M for one GPU, 128M otherwise")
    ap.add_argument("--theta", type=float, default=0.75)
    ap.add_argument("--max-leaf-n", type=int, default=16)
    ap.add_argument("--ncrit", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.path = f"/tmp/rk_clocks_{os.getpid()}.csv"
        self.p = None

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            c = [v.strip() for v in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            # "under load": the upper half of the samples (idle gaps between steps pull the clock down)
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline (the ONLY users of oracle/ in this file)
# ----------------------------------------------------------------------------------------------------------
def cpu_reference_run(args, nparts, steps, warmup, target_seconds, chunk=0):
    """Times the reference's CPU implementation of the path on the host cores.

    Preferred: oracle/_ref (the UNMODIFIED reference header compiled against dependency stand-ins, SIMD + rsqrt
    path, OpenMP-backed TBB stand-in) — kind "reference": every step is a full build + full accs_u of the
    workload. Fallback: the scalar oracle port on a strided sample of the critical nodes — kind "port".
    Interactions per evaluation come from the oracle's counters on a strided sample of the same tree (the
    reference has no interaction counter)."""
    import oracle  # noqa: test infrastructure, allowed here only
    import rakau_b200 as rk
    cores = os.cpu_count() or 1
    m, x, y, z = rk.plummer(nparts, 0, nparts, chunk=chunk) if chunk else rk.plummer(nparts)
    t0 = time.time()
    otree = oracle.OracleTree(x, y, z, m, max_leaf_n=args.max_leaf_n, ncrit=args.ncrit)
    t_obuild = time.time() - t0
    ncrit_nodes = len(otree.crit()[0])
    # interactions of one evaluation, from a strided sample of the oracle's counters (exact when stride = 1)
    cstride = max(1, ncrit_nodes // 8000)
    t0 = time.time()
    c = otree.acc_pot_sample(0, args.theta, cstride, cstride // 2, nthreads=cores)
    t_probe = max(time.time() - t0, 1e-4)
    i_total = float(c["interactions"]) * cstride
    variant = oracle.best_ref_variant()
    nsteps = max(1, steps + warmup)
    if variant is not None:
        # pick the faster SIMD width on a small problem (AVX-512 is not always the faster one)
        cands = [v for v in ("avx512", "avx2") if oracle.ref_available(v) and (v != "avx512" or variant == "avx512")]
        if len(cands) > 1:
            pm, px, py, pz = rk.plummer(200000)
            best = None
            for v in cands:
                rt = oracle.RefTree(px, py, pz, pm, max_leaf_n=args.max_leaf_n, ncrit=args.ncrit, variant=v)
                rt.acc_pot(0, args.theta)
                t0 = time.time()
                rt.acc_pot(0, args.theta)
                dt = time.time() - t0
                if best is None or dt < best[0]:
                    best = (dt, v)
            variant = best[1]
        tb, ta = [], []
        label = None
        for s in range(nsteps):
            t0 = time.time()
            rt = oracle.RefTree(x, y, z, m, max_leaf_n=args.max_leaf_n, ncrit=args.ncrit, variant=variant)
            t1 = time.time()
            rt.acc_pot(0, args.theta)
            t2 = time.time()
            label = rt.variant()
            del rt
            if s >= warmup:
                tb.append(t1 - t0)
                ta.append(t2 - t1)
            if sum(tb) + sum(ta) > target_seconds and len(ta) >= 1:
                break
        t_full = float(np.mean(tb) + np.mean(ta))
        sample = (f"{label}: full workload per step (construct octree {np.mean(tb):.3f} s + accs_u {np.mean(ta):.3f} s), "
                  f"{len(ta)} timed steps, {cores} host threads; interactions/eval {i_total:.4g} from the oracle's "
                  f"counters on every {cstride}-th critical node")
        return dict(value=i_total / t_full / 1e9, ms_per_step=t_full * 1e3, cores=cores, kind="reference",
                    sample=sample, traversal_ginter_s=i_total / float(np.mean(ta)) / 1e9, build_s=float(np.mean(tb)),
                    interactions=i_total)
    # ---- fallback: scalar oracle port, strided sample ----
    est_full = t_probe * cstride
    stride = max(1, int(np.ceil(est_full * nsteps / max(target_seconds, 1.0))))
    times, inter = [], []
    for s in range(nsteps):
        t0 = time.time()
        c = otree.acc_pot_sample(0, args.theta, stride, s % stride, nthreads=cores)
        dt = time.time() - t0
        if s >= warmup:
            times.append(dt)
            inter.append(c["interactions"])
    rate = sum(inter) / sum(times)
    i_total = float(np.mean(inter)) * stride
    t_full = t_obuild + i_total / rate
    sample = (f"oracle scalar port: full build once ({t_obuild:.2f} s, 1 thread) + traversal of every {stride}-th "
              f"critical node per step ({len(times)} timed steps, {cores} threads, {sum(times):.1f} s CPU wall); "
              f"extrapolated by interaction count")
    return dict(value=i_total / t_full / 1e9, ms_per_step=t_full * 1e3, cores=cores, kind="port", sample=sample,
                traversal_ginter_s=rate / 1e9, build_s=t_obuild, interactions=i_total)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the arm's own workload: 4 M particles at N = 1, the 128 M strong-scaling workload (chunked generator) at N > 1,
    # where the number of timed steps is bounded by wall clock (one evaluation takes ~15-20 s on 16 host cores)
    big = args.gpus > 1 and not args.nparts
    nparts = args.nparts or (128_000_000 if big else 4_000_000)
    r = cpu_reference_run(args, nparts, args.steps, min(args.warmup, 1) if big else args.warmup, 60.0 if big else 150.0,
                          chunk=(1 << 20) if args.gpus > 1 else 0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "Ginteractions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"plummer_{nparts}_fp32_theta{args.theta}_accs", "max_leaf_n": args.max_leaf_n,
                   "ncrit": args.ncrit, "nparts": nparts},
        "cpu_baseline": {"value": r["value"], "unit": "Ginteractions/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "Ginteractions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def strong_scaling_base(nparts, ms_per_step):
    """The same workload on ONE GPU, from the committed profile (the N = 1 line of this script measures the 4 M
    headline workload, not this one): lets a reader form the strong-scaling speed-up on equal work."""
    try:
        path = os.path.join(ROOT, "profiles", "r01_bench_1gpu_128M_v7.json")
        d = json.loads(open(path).read().strip().splitlines()[-1])
        if int(d["config"]["nparts"]) != int(nparts):
            return None
        return {"n_gpus": 1, "ms_per_step": d["ms_per_step"], "value": d["value"], "source": os.path.relpath(path, ROOT),
                "speedup_vs_1gpu": d["ms_per_step"] / ms_per_step}
    except Exception:  # informational only
        return None


# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import rakau_b200 as rk
    from rakau_b200 import sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nparts = args.nparts or (4_000_000 if world == 1 else 128_000_000)
    chunk = 1 << 20
    per = (nparts + world - 1) // world
    per = (per + chunk - 1) // chunk * chunk if world > 1 else nparts
    first = min(rank * per, nparts)
    count = max(0, min(nparts, first + per) - first)

    # ---- synthetic inputs: pinned host shards (benchmark/common.hpp Plummer sphere) ----
    hx, hy, hz, hm = (torch.empty(count, dtype=torch.float32).pin_memory() for _ in range(4))
    rk.plummer(nparts, first, count, fp=32, chunk=(chunk if world > 1 else 0),
               out=[hm.numpy(), hx.numpy(), hy.numpy(), hz.numpy()])
    stream = torch.cuda.current_stream()
    tree = rk.Octree(fp=32, mac="bh", device=local)
    tree.set_stream(stream.cuda_stream)

    # resident copies of the shard (inputs in HBM when the timed region starts)
    dsh = [t.to(dev, non_blocking=True) for t in (hx, hy, hz, hm)]
    counts = [max(0, min(nparts, min(r * per, nparts) + per) - min(r * per, nparts)) for r in range(world)]
    out_dev = [torch.zeros(nparts, dtype=torch.float32, device=dev) for _ in range(3)] if world == 1 else None
    hout = [torch.empty(nparts, dtype=torch.float32).pin_memory() for _ in range(3)]
    hnp = [t.numpy() for t in (hx, hy, hz, hm)]
    hout_np = [t.numpy() for t in hout]
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2
    state = {"info": None, "bi": None, "imbalance": None, "d2h": 0}

    sharded = None
    if world > 1:
        from rakau_b200.distributed import ShardedTree
        sharded = ShardedTree(dist, dev, fp=32, mac="bh", max_leaf_n=args.max_leaf_n, ncrit=args.ncrit)
        tree = sharded.tree

    def step(e2e):
        if world == 1 and e2e:
            # end to end through the C ABI with (pinned) HOST buffers: rk_tree_build copies the shard in,
            # rk_tree_acc_pot copies the accelerations out (pipelined with the traversal launches)
            state["bi"] = tree.build(hnp[0], hnp[1], hnp[2], hnp[3], max_leaf_n=args.max_leaf_n, ncrit=args.ncrit,
                                     where=rk.RK_HOST)
            tree.acc_pot(0, args.theta, out=hout_np, where=rk.RK_HOST)
            state["info"] = tree.eval_info.asdict()
            state["d2h"] = 12 * nparts
            return
        src = [t.to(dev, non_blocking=True) for t in (hx, hy, hz, hm)] if e2e else dsh
        if world == 1:
            state["bi"] = tree.build(src[0], src[1], src[2], src[3], max_leaf_n=args.max_leaf_n, ncrit=args.ncrit,
                                     where=rk.RK_DEVICE, n=nparts)
            tree.acc_pot(0, args.theta, out=out_dev, where=rk.RK_DEVICE)
            state["info"] = tree.eval_info.asdict()
            lo, hi = 0, nparts
        else:
            # distributed sample sort + replicated topology, Morton-range sharded traversal, output exchange
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            state["bi"] = sharded.build(src[0], src[1], src[2], src[3], first_index=first)
            ev[1].record()
            # (outputs: library-owned peer-memory buffers, complete on every rank when the call returns)
            state["info"], outs = sharded.acc_pot(0, args.theta)
            ev[2].record()
            state["phase_events"] = ev
            lo, hi = int(sharded.cut_particles[rank]), int(sharded.cut_particles[rank + 1])
        if e2e:
            for j in range(3):
                hout[j][lo:hi].copy_(outs[j][lo:hi], non_blocking=True)  # each rank returns the slice it owns
            stream.synchronize()
        state["d2h"] = 12 * (hi - lo)

    def refresh_costs():
        if world > 1:
            state["imbalance"] = sharded.rebalance(kernel_ms=state["info"]["ms_kernel"])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(nsteps, e2e):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
        infos = []
        barrier()
        for a, b in evs:
            flush.zero_()  # L2 flush between timed iterations (outside the timed region)
            a.record()
            step(e2e)
            b.record()
            infos.append((state["info"], state["bi"].asdict()))
        barrier()
        ms = [a.elapsed_time(b) for a, b in evs]
        tot = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()), ms, infos

    # ---- warm-up (also establishes the cost-weighted split for N > 1) ----
    for w in range(max(args.warmup, 3)):
        step(False)
        refresh_costs()  # N > 1: cost-weighted cuts from the previous evaluation (converges in 2-3 steps)
    step(True)
    fp32_peak = rk.measure_fp32_peak(local)

    clocks = ClockSampler(local)
    clocks.start()
    l0 = rk.kernel_launch_count()
    tot_ms, ms, infos = timed(args.steps, False)
    launches = rk.kernel_launch_count() - l0
    clk = clocks.stop()
    e2e_tot_ms, e2e_ms, _ = timed(args.steps, True)

    # whole-job interactions per step = sum over ranks
    inter = torch.tensor([infos[-1][0]["interactions"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(inter)
    inter = float(inter.item())
    ms_per_step = tot_ms / args.steps
    value = inter / (ms_per_step * 1e-3) / 1e9
    e2e_value = inter / (e2e_tot_ms / args.steps * 1e-3) / 1e9
    k_ms = float(np.mean([i[0]["ms_kernel"] for i in infos]))
    b_ms = float(np.mean([i[1]["ms_total"] for i in infos]))
    my_inter = infos[-1][0]["interactions"]
    achieved = SLOTS_PER_INTERACTION * 2 * my_inter / (k_ms * 1e-3) / 1e12  # FMA-equivalent TFLOP/s
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traverse_kernel_latest.json")))
        traffic = prof.get("dram_bytes_per_launch")
    except Exception:
        pass
    bi = infos[-1][1]
    # algorithmic build bytes (fp32, u64 codes), SURVEY §8(d): pack 16+16, encode 16+8, sort passes x (8 + 12 + 12),
    # gather 4+16+16, perms 4+4+4, topology ~12, plus 64 B per node
    build_bytes = nparts * (32 + 24 + bi["sort_passes"] * 32 + 36 + 12 + 12) + bi["n_nodes"] * 64
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    line = {
        "metric": METRIC, "value": value, "unit": "Ginteractions/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"plummer_{nparts}_fp32_theta{args.theta}_accs", "nparts": nparts,
                   "max_leaf_n": args.max_leaf_n, "ncrit": args.ncrit, "mac": "bh", "G": 1.0, "eps": 0.0,
                   "parallelism": (f"sample_sort_build+morton_range_traversal_x{world}" if world > 1 else "single_gpu"),
                   "l2": "256 MiB buffer written between timed iterations", "interactions_per_step": inter,
                   "n_nodes": bi["n_nodes"], "n_crit": bi["n_crit"], "shard_cost_imbalance": state["imbalance"]},
        "ms_build": b_ms, "ms_traverse_kernel": k_ms,
        "build_phases_ms": {k: bi[k] for k in ("ms_encode", "ms_sort", "ms_permute", "ms_topology", "ms_props")},
        "roofline": {"bound": "fp32", "kernel": "traverse_kernel<float,0,0,64>", "achieved": achieved,
                     "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak, "traffic": traffic,
                     "convention": "12 FP32 issue slots per interaction counted as FMA (2 flop); peak = FFMA "
                                   "microbenchmark measured in this run",
                     "gflops_literature_20flop": FLOP_PER_INTERACTION_LIT * my_inter / (k_ms * 1e-3) / 1e9},
        "roofline_build": {"bound": "hbm", "achieved": build_bytes / (b_ms * 1e-3) / 1e9, "peak": hbm_peak,
                           "unit": "GB/s", "frac": build_bytes / (b_ms * 1e-3) / 1e9 / hbm_peak,
                           "peak_source": "measured" if peaks else "fallback"},
        "e2e": {"value": e2e_value, "unit": "Ginteractions/s", "ms_per_step": e2e_tot_ms / args.steps,
                "h2d_bytes_per_step": 16 * count, "d2h_bytes_per_step": state["d2h"],
                "note": "per rank: its input shard in, the output slice it owns out" if world > 1 else "rk_tree_build + rk_tree_acc_pot with pinned HOST buffers: all inputs in, all outputs out"},
        "gpu_launches": int(launches), "clocks": clk,
        "vs_published_ms": {"note": "reference README traversal-only times, other hardware", "v100_ms": 95,
                            "xeon6148x2_ms": 82, "ours_traverse_ms": k_ms},
    }
    if world > 1:
        torch.cuda.synchronize()
        ev = state["phase_events"]
        line["ms_build"] = ev[0].elapsed_time(ev[1])
        line["ms_traverse_and_exchange"] = ev[1].elapsed_time(ev[2])
        line["build_phases_rank0_ms"] = sharded.phase_ms()
        if os.environ.get("RK_DEBUG_BARRIER"):
            allp = [None] * world
            dist.all_gather_object(allp, sharded.phase_ms())
            line["build_phases_all_ranks_ms"] = allp
        km = torch.tensor([state["info"]["ms_kernel"]], dtype=torch.float64, device=dev)
        kall = [torch.zeros_like(km) for _ in range(world)]
        dist.all_gather(kall, km)
        line["ms_traverse_kernel_per_rank"] = [float(k.item()) for k in kall]
        line["build_note"] = ("ms_build = distributed sample sort (local sort, all-to-all, bucket sort, all-gather) + "
                              "replicated topology/properties; build_phases_ms covers the replicated part only")
        line["strong_scaling_base"] = strong_scaling_base(nparts, ms_per_step)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_run(args, nparts, 4, 1, args.cpu_seconds)
            line["cpu_baseline"] = {"value": r["value"], "unit": "Ginteractions/s", "cores": r["cores"],
                                    "kind": r["kind"], "sample": r["sample"]}
        except Exception as e:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": "Ginteractions/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {e}"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
