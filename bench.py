#!/usr/bin/env python
"""bench.py — headline benchmark of the Barnes-Hut hot path (BASELINE.json): one "step" = one acceleration
evaluation = tree build (Morton encode, sort, octree, node properties) + ncrit-grouped traversal.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W

N = 1: BASELINE config 1 (3D Plummer, 4M particles, fp32, theta = 0.75, accelerations, max_leaf_n 16,
ncrit 128 — the parameters behind the reference's published table). The same line carries, under "configs",
driver-visible measurements of BASELINE configs 2 (accs+pots, eps, G), 3 (fp64, theta 0.5, FP64-pipe fraction),
4 (16M leapfrog, device-resident and end to end) and the one-GPU base of config 5 (128M, same chunked generator
as the N > 1 runs). N > 1: BASELINE config 5 (128M particles, strong scaling): inputs sharded over the ranks,
distributed sample sort, identical tree on every GPU, critical nodes sharded by cost-weighted Morton ranges,
outputs exchanged over peer memory; before timing, the sharded tree and the gathered accelerations are checked
against a single-GPU evaluation of the same particles, bit for bit ("parity_checked").

Prints ONE JSON line (rank 0). `value` = interactions / device time with inputs resident in HBM;
`e2e` = the same through the C ABI with pinned HOST buffers (H2D + D2H inside the timed region).
`--impl reference` times the reference's CPU algorithm on the host cores (oracle/_ref when it was built,
else the scalar oracle port) — the only place besides `cpu_baseline` where oracle/ is executed; that arm does not
load the product library.
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
L2_NOTE = "GPU arm: 256 MiB device buffer written between timed iterations"


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
    ap.add_argument("--no-extras", action="store_true", help="skip the configs 2/3/4/5-base blocks of the N = 1 line")
    ap.add_argument("--no-parity-check", action="store_true", help="N > 1: skip the single-GPU comparison")
    ap.add_argument("--perturb", action="store_true",
                    help="N > 1: move the particles a little between steps (cost-weighted cuts must survive a rebuild)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def workload_config(args, nparts):
    """The workload, with the same keys and values in both arms (the driver compares them)."""
    return {"workload": f"plummer_{nparts}_fp32_theta{args.theta}_accs", "nparts": nparts, "fp": 32, "theta": args.theta,
            "max_leaf_n": args.max_leaf_n, "ncrit": args.ncrit, "mac": "bh", "G": 1.0, "eps": 0.0,
            "generator": "benchmark/common.hpp:39-126, " + ("chunks of 2^20 particles, each seeded with its first index"
                                                           if args.gpus > 1 else "sequential branch"),
            "l2": L2_NOTE}


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
            time.sleep(0.25)  # nvidia-smi needs ~0.2 s before its first sample
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
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline (the ONLY users of oracle/ in this file; the product library is not imported here)
# ----------------------------------------------------------------------------------------------------------
def exact_interactions(args, nparts, chunk):
    """Interactions of one evaluation, when a committed count exists. Config 1: tests/golden/baseline_sizes.json, counted
    by the oracle on the reference's own tree. Config 5 (128 M, chunked generator): the count of the CUDA kernel on the
    device-built tree (profiles/r02_bench_1gpu.json, configs.strong_scaling_base_128M) - the kernel's counts equal the
    oracle's on the same tree (T1) and the device tree differs from the reference's only in the last ulps of the node
    COMs (a relative 1e-7 of the count at 4 M); counting 229 G interactions with the scalar oracle would take the
    reference arm minutes of host time."""
    if args.theta != 0.75 or args.max_leaf_n != 16 or args.ncrit != 128:
        return None, None
    if chunk == (1 << 20) and nparts == 128_000_000:
        return 229285644471.0, "counted by the CUDA kernel on the device-built tree (profiles/r02_bench_1gpu.json)"
    if chunk or nparts != 4_000_000:
        return None, None
    try:
        fix = json.load(open(os.path.join(ROOT, "tests", "golden", "baseline_sizes.json")))
        return float(fix["config1_fp32_accs"]["counters"]["interactions"]), "exact (tests/golden/baseline_sizes.json)"
    except Exception:
        return None, None


def cpu_reference_run(args, nparts, steps, warmup, target_seconds, chunk=0):
    """Times the reference's CPU implementation of the path on the host cores.

    Preferred: oracle/_ref (the UNMODIFIED reference header compiled against dependency stand-ins, SIMD + rsqrt
    path, OpenMP-backed TBB stand-in) — kind "reference": every step is a full build + full accs_u of the
    workload. Fallback: the scalar oracle port on a strided sample of the critical nodes — kind "port".
    Interactions per evaluation: the committed exact count for config 1, else the oracle's counters on a strided
    sample of the same tree (the reference has no interaction counter). Returns the steps actually timed."""
    import oracle  # noqa: test infrastructure, allowed here only
    cores = os.cpu_count() or 1
    m, x, y, z = oracle.plummer(nparts, chunk=chunk, nthreads=cores)
    i_total, i_note = exact_interactions(args, nparts, chunk)
    otree, t_obuild, t_probe, cstride = None, 0.0, 0.0, 1
    variant = oracle.best_ref_variant()
    if i_total is None or variant is None:
        t0 = time.time()
        otree = oracle.OracleTree(x, y, z, m, max_leaf_n=args.max_leaf_n, ncrit=args.ncrit)
        t_obuild = time.time() - t0
        ncrit_nodes = len(otree.crit()[0])
        cstride = max(1, ncrit_nodes // 8000)
        t0 = time.time()
        c = otree.acc_pot_sample(0, args.theta, cstride, cstride // 2, nthreads=cores)
        t_probe = max(time.time() - t0, 1e-4)
        if i_total is None:
            i_total = float(c["interactions"]) * cstride
            i_note = f"the oracle's counters on every {cstride}-th critical node"
    nsteps = max(1, steps + warmup)
    if variant is not None:
        # pick the faster SIMD width on a small problem (AVX-512 is not always the faster one)
        cands = [v for v in ("avx512", "avx2") if oracle.ref_available(v) and (v != "avx512" or variant == "avx512")]
        if len(cands) > 1:
            pm, px, py, pz = oracle.plummer(200000)
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
        t_start = time.time()
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
            if time.time() - t_start > target_seconds and len(ta) >= 1:
                break
        t_full = float(np.mean(tb) + np.mean(ta))
        sample = (f"{label}: full workload per step (construct octree {np.mean(tb):.3f} s + accs_u {np.mean(ta):.3f} s), "
                  f"{len(ta)} timed steps, {cores} host threads; interactions/eval {i_total:.6g}: {i_note}")
        return dict(value=i_total / t_full / 1e9, ms_per_step=t_full * 1e3, cores=cores, kind="reference",
                    sample=sample, traversal_ginter_s=i_total / float(np.mean(ta)) / 1e9, build_s=float(np.mean(tb)),
                    interactions=i_total, timed_steps=len(ta))
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
    t_full = t_obuild + i_total / rate
    sample = (f"oracle scalar port: full build once ({t_obuild:.2f} s, 1 thread) + traversal of every {stride}-th "
              f"critical node per step ({len(times)} timed steps, {cores} threads, {sum(times):.1f} s CPU wall); "
              f"extrapolated by interaction count")
    return dict(value=i_total / t_full / 1e9, ms_per_step=t_full * 1e3, cores=cores, kind="port", sample=sample,
                traversal_ginter_s=rate / 1e9, build_s=t_obuild, interactions=i_total, timed_steps=len(times))


def reference_cuda_run(args, nparts, reps=3):
    """Second baseline (north_star: "next to the reference's own CUDA backend if it builds offline"): the unmodified
    reference header with RAKAU_WITH_CUDA and its own src/rakau_cuda.cu compiled for sm_100a (oracle/_ref/libref_cuda.so),
    driven through the reference's public API: accs_u(..., split = {0, 1}) = everything on its GPU path. The tree is
    built by the reference on the CPU (not timed here); every call uploads tree + particles and downloads the results
    (stateless by design, rakau_cuda.cu:348-528), which its own timer includes as well (tree.hpp:3297)."""
    import oracle  # noqa: test infrastructure
    if not oracle.ref_available("cuda"):
        return {"unavailable": "oracle/_ref/libref_cuda.so was not built"}
    m, x, y, z = oracle.plummer(nparts)
    t0 = time.time()
    rt = oracle.RefTree(x, y, z, m, max_leaf_n=args.max_leaf_n, ncrit=args.ncrit, variant="cuda")
    t_build = time.time() - t0
    ts = []
    for _ in range(reps + 1):
        t0 = time.time()
        rt.acc_pot(0, args.theta, split=[0.0, 1.0])
        ts.append(time.time() - t0)
    return {"what": rt.variant(), "ms_per_accs_u_call": 1e3 * float(np.median(ts[1:])), "calls": reps,
            "cpu_tree_build_ms_not_included": 1e3 * t_build,
            "note": "per-particle MAC and traversal (rakau_cuda.cu:152-335), host-to-device copy of tree and particles "
                    "and device-to-host copy of the results inside every call; README V100 figure for this workload: 95 ms"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the arm's own workload: 4 M particles at N = 1, the 128 M strong-scaling workload (chunked generator) at N > 1,
    # where one evaluation takes ~15 s on 16 host cores: the run is bounded by wall clock and reports the steps it
    # actually timed
    big = args.gpus > 1 and not args.nparts
    nparts = args.nparts or (128_000_000 if big else 4_000_000)
    r = cpu_reference_run(args, nparts, args.steps, min(args.warmup, 1) if big else args.warmup, 120.0 if big else 150.0,
                          chunk=(1 << 20) if args.gpus > 1 else 0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "Ginteractions/s", "n_gpus": args.gpus,
        "steps": r["timed_steps"], "warmup": min(args.warmup, 1) if big else args.warmup,
        "steps_requested": args.steps, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, nparts),
        "cpu_baseline": {"value": r["value"], "unit": "Ginteractions/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "Ginteractions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    if os.environ.get("RK_BENCH_PRINT_MAPS"):  # tests: the reference arm must not map the product library
        libs = sorted({ln.split()[-1] for ln in open("/proc/self/maps") if ln.rstrip().endswith(".so")})
        print("\n".join(libs), file=sys.stderr)


# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
def strong_scaling_base(nparts, ms_per_step):
    """The same workload on ONE GPU, from the committed N = 1 line (its `configs.strong_scaling_base_128M` block is
    measured by every N = 1 run of this script with the same chunked generator): lets a reader form the strong-scaling
    speed-up on equal work without mixing it with the 4 M headline workload."""
    try:
        path = os.path.join(ROOT, "profiles", "r02_bench_1gpu.json")
        b = json.load(open(path))["configs"]["strong_scaling_base_128M"]
        if int(b["nparts"]) != int(nparts):
            return None
        return {"n_gpus": 1, "ms_per_step": b["ms_per_step"], "ginteractions_per_s": b["ginteractions_per_s"],
                "source": os.path.relpath(path, ROOT) + " configs.strong_scaling_base_128M",
                "speedup_vs_1gpu": b["ms_per_step"] / ms_per_step}
    except Exception:  # informational only
        return None


def median(v):
    return float(statistics.median(v))


def kernel_traffic(kernel_name):
    """DRAM bytes per launch of the traversal kernel from the committed ncu capture - only when that capture is of
    the kernel variant this run launched."""
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traverse_kernel_latest.json")))
        if prof.get("kernel") and prof["kernel"].split(" ")[0] == kernel_name.split(" ")[0]:
            return prof.get("dram_bytes_per_launch")
    except Exception:
        pass
    return None


def extras_single_gpu(args, rk, torch, dev, tree, dsh, nparts, fp32_peak, flush):
    """Driver-visible measurements of BASELINE configs 2, 3, 4 and the one-GPU base of config 5 (N = 1 line)."""
    out = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def timed_evals(t, src, Q, theta, G, eps, nsteps, n, outs):
        ms, ks, info, bi = [], [], None, None
        for s in range(nsteps + 1):
            flush.zero_()
            a, b = ev(), ev()
            a.record()
            bi = t.build(src[0], src[1], src[2], src[3], max_leaf_n=args.max_leaf_n, ncrit=args.ncrit,
                         where=rk.RK_DEVICE, n=n)
            t.acc_pot(Q, theta, G=G, eps=eps, out=outs, where=rk.RK_DEVICE)
            b.record()
            torch.cuda.synchronize()
            if s:  # first = warm-up
                ms.append(a.elapsed_time(b))
                ks.append(t.eval_info.ms_kernel)
            info = t.eval_info.asdict()
        return median(ms), median(ks), info, bi.asdict()

    # ---- config 2: same particles, accelerations + potentials, softening, G != 1 ----
    def config2():
        outs = [torch.empty(nparts, dtype=torch.float32, device=dev) for _ in range(4)]
        ms, k, info, _ = timed_evals(tree, dsh, 2, args.theta, 2.5, 0.01, 5, nparts, outs)
        ach = (SLOTS_PER_INTERACTION + 1) * 2 * info["interactions"] / (k * 1e-3) / 1e12
        return {"ms_per_step": ms, "ms_traverse_kernel": k,
                "ginteractions_per_s": info["interactions"] / (ms * 1e-3) / 1e9,
                "interactions": info["interactions"], "kernel": tree.last_kernel(),
                "roofline": {"bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                             "frac": ach / fp32_peak, "convention": "13 FP32 slots per interaction (accs + pots)"}}

    # ---- config 3: fp64, theta = 0.5 ----
    def config3():
        h64 = rk.plummer(nparts, fp=64)
        d64 = [torch.from_numpy(a).to(dev) for a in (h64[1], h64[2], h64[3], h64[0])]
        t64 = rk.Octree(fp=64, mac="bh", device=dev.index)
        t64.set_stream(torch.cuda.current_stream().cuda_stream)
        outs = [torch.empty(nparts, dtype=torch.float64, device=dev) for _ in range(3)]
        ms, k, info, _ = timed_evals(t64, d64, 0, 0.5, 1.0, 0.0, 3, nparts, outs)
        fp64_peak = rk.measure_fp64_peak(dev.index)
        ach = SLOTS_PER_INTERACTION * 2 * info["interactions"] / (k * 1e-3) / 1e12
        res = {"ms_per_step": ms, "ms_traverse_kernel": k,
               "ginteractions_per_s": info["interactions"] / (ms * 1e-3) / 1e9,
               "interactions": info["interactions"], "kernel": t64.last_kernel(),
               "roofline": {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                            "frac": ach / fp64_peak,
                            "convention": "12 FP64 slots per interaction counted as DFMA (2 flop) + the double rsqrt "
                                          "(not counted); peak = DFMA microbenchmark of this run"}}
        t64.close()
        return res

    for name, fn in (("config2_accs_pots_eps0.01_G2.5", config2), ("config3_fp64_theta0.5", config3)):
        try:
            out[name] = fn()
        except Exception as e:  # each block is reported on its own
            out[name] = {"error": repr(e)}
        torch.cuda.empty_cache()
    # ---- config 4: 16 M leapfrog (benchmark_leapfrog.cpp), device-resident and end to end ----
    if hasattr(rk, "Leapfrog"):
        try:
            out["config4_leapfrog_16M"] = rk.leapfrog_benchmark(16_000_000, 10, theta=args.theta,
                                                                max_leaf_n=args.max_leaf_n, ncrit=args.ncrit,
                                                                device=dev.index)
        except Exception as e:  # reported, never fatal for the headline line
            out["config4_leapfrog_16M"] = {"error": repr(e)}
        torch.cuda.empty_cache()
    # ---- config 5, one GPU: the strong-scaling base, same chunked generator as the N > 1 runs ----
    try:
        nb = 128_000_000
        t0 = time.time()
        hb = rk.plummer(nb, 0, nb, fp=32, chunk=1 << 20)
        tg = time.time() - t0
        db = [torch.from_numpy(a).to(dev) for a in (hb[1], hb[2], hb[3], hb[0])]
        del hb
        outs = [torch.empty(nb, dtype=torch.float32, device=dev) for _ in range(3)]
        ms, k, info, bi = timed_evals(tree, db, 0, args.theta, 1.0, 0.0, 2, nb, outs)
        out["strong_scaling_base_128M"] = {
            "n_gpus": 1, "nparts": nb, "ms_per_step": ms, "ms_build": bi["ms_total"], "ms_traverse_kernel": k,
            "ginteractions_per_s": info["interactions"] / (ms * 1e-3) / 1e9, "interactions": info["interactions"],
            "n_nodes": bi["n_nodes"], "n_crit": bi["n_crit"], "generator_s": tg,
            "note": "same workload and generator as bench.py --gpus N > 1; divide by that run's ms_per_step for the "
                    "strong-scaling speed-up on equal work"}
        del db, outs
    except Exception as e:
        out["strong_scaling_base_128M"] = {"error": repr(e)}
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import rakau_b200 as rk

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
    out_dev = [torch.zeros(nparts, dtype=torch.float32, device=dev) for _ in range(3)] if world == 1 else None
    hout = [torch.empty(nparts if world == 1 else count, dtype=torch.float32).pin_memory() for _ in range(3)]
    hnp = [t.numpy() for t in (hx, hy, hz, hm)]
    hout_np = [t.numpy() for t in hout]
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2
    state = {"info": None, "bi": None, "imbalance": None, "d2h": 0, "step": 0}

    sharded = None
    if world > 1:
        from rakau_b200.distributed import ShardedTree
        sharded = ShardedTree(dist, dev, fp=32, mac="bh", max_leaf_n=args.max_leaf_n, ncrit=args.ncrit)
        tree = sharded.tree

    def step(e2e):
        state["step"] += 1
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
            if args.perturb:
                # a small, deterministic displacement per step: the tree (and its critical nodes) changes
                s = 1e-4 * (1 + state["step"] % 3)
                src = [src[0] + s * src[1], src[1] - s * src[0], src[2], src[3]]
            # distributed sample sort + replicated topology, Morton-range sharded traversal, output exchange
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            state["bi"] = sharded.build(src[0], src[1], src[2], src[3], first_index=first)
            ev[1].record()
            # (outputs: library-owned peer-memory buffers, complete on every rank when the call returns)
            # (end to end: each finished launch's slice of this rank's range goes to the pinned host buffers while the
            # next launch runs)
            rng = None
            if e2e:
                sharded.ensure_cuts()
                rng = int(sharded.cut_particles[rank + 1]) - int(sharded.cut_particles[rank])
                if hout[0].numel() < rng:
                    hout[:] = [torch.empty(int(rng * 1.1), dtype=torch.float32).pin_memory() for _ in range(3)]
            state["info"], outs = sharded.acc_pot(0, args.theta, host_out=hout if rng is not None else None)
            ev[2].record()
            state["phase_events"] = ev
            state["outs"] = outs
            lo, hi = int(sharded.cut_particles[rank]), int(sharded.cut_particles[rank + 1])
        if e2e:
            if world > 1 and hout[0].numel() < hi - lo:  # the cost-weighted range of this rank outgrew the buffers
                hout[:] = [torch.empty(int((hi - lo) * 1.1), dtype=torch.float32).pin_memory() for _ in range(3)]
            if world == 1:
                for j in range(3):
                    hout[j][lo:hi].copy_(out_dev[j][lo:hi], non_blocking=True)
            elif rng is None:
                for j in range(3):
                    hout[j][:hi - lo].copy_(outs[j][lo:hi], non_blocking=True)  # each rank returns the slice it owns
            stream.synchronize()  # (the per-launch copies were ordered before the end of acc_pot)
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
    parity = None
    if world > 1 and not args.no_parity_check:
        perturb, args.perturb = args.perturb, False
        step(False)  # (an unperturbed evaluation of the resident shard, compared with a single-GPU one)
        args.perturb = perturb
        parity = sharded.check_against_single_gpu(dsh, first, args.theta, state["outs"])
    step(True)
    fp32_peak = rk.measure_fp32_peak(local)

    clocks = ClockSampler(local)
    clocks.start()
    l0 = rk.kernel_launch_count()
    tot_ms, ms, infos = timed(args.steps, False)
    launches = rk.kernel_launch_count() - l0
    clk = clocks.stop()
    e2e_tot_ms, e2e_ms, _ = timed(args.steps, True)
    kernel_name = tree.last_kernel()
    pageable_ms = None
    if world == 1:
        # the same end-to-end call from PAGEABLE host memory (what the C++ API's std::vector arguments are)
        pin, pout = [np.array(a) for a in hnp], [np.empty(nparts, dtype=np.float32) for _ in range(3)]
        ts = []
        for _ in range(6):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            tree.build(pin[0], pin[1], pin[2], pin[3], max_leaf_n=args.max_leaf_n, ncrit=args.ncrit, where=rk.RK_HOST)
            tree.acc_pot(0, args.theta, out=pout, where=rk.RK_HOST)
            ts.append(1e3 * (time.perf_counter() - t0))
        pageable_ms = median(ts[1:])

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
    bi = infos[-1][1]
    # algorithmic build bytes (fp32, u64 codes), SURVEY §8(d): pack 16+16, encode 16+8, one histogram pass 8,
    # sort passes x (12 + 12), gather 4+16+16, perms 4+4+4, topology ~12, plus 64 B per node
    build_bytes = nparts * (32 + 24 + 8 + bi["sort_passes"] * 24 + 36 + 12 + 12) + bi["n_nodes"] * 64
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
        "config": workload_config(args, nparts),
        "tree": {"interactions_per_step": inter, "n_nodes": bi["n_nodes"], "n_crit": bi["n_crit"],
                 "sharding": (f"sample_sort_build+morton_range_traversal_x{world}" if world > 1 else "single_gpu"),
                 "shard_cost_imbalance": state["imbalance"],
                 "output_exchange": sharded.exchange_mode if sharded is not None else None,
                 "codes_gather": sharded.codes_gather_mode if sharded is not None else None},
        "ms_build": b_ms, "ms_traverse_kernel": k_ms,
        "build_phases_ms": {k: bi[k] for k in ("ms_encode", "ms_sort", "ms_permute", "ms_topology", "ms_props")},
        "roofline": {"bound": "fp32", "kernel": kernel_name, "achieved": achieved,
                     "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                     "traffic": kernel_traffic(kernel_name),
                     "convention": "12 FP32 issue slots per interaction counted as FMA (2 flop); peak = FFMA "
                                   "microbenchmark measured in this run",
                     "gflops_literature_20flop": FLOP_PER_INTERACTION_LIT * my_inter / (k_ms * 1e-3) / 1e9},
        "roofline_build": {"bound": "hbm", "achieved": build_bytes / (b_ms * 1e-3) / 1e9, "peak": hbm_peak,
                           "unit": "GB/s", "frac": build_bytes / (b_ms * 1e-3) / 1e9 / hbm_peak,
                           "peak_source": "measured" if peaks else "fallback"},
        "e2e": {"value": e2e_value, "unit": "Ginteractions/s", "ms_per_step": e2e_tot_ms / args.steps,
                "h2d_bytes_per_step": 16 * count, "d2h_bytes_per_step": state["d2h"],
                "pageable_host_buffers_ms_per_step": pageable_ms,
                "note": "per rank: its input shard in, the output slice it owns out" if world > 1 else "rk_tree_build + rk_tree_acc_pot with pinned HOST buffers: all inputs copied in, all outputs written into the host buffers by the traversal kernel (mapped memory, no copy behind the launch)"},
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
        line["build_note"] = ("ms_build = distributed sample sort (local sort, bucket exchange, bucket sort, gather) + "
                              "replicated topology/properties; build_phases_ms covers the replicated part only")
        line["strong_scaling_base"] = strong_scaling_base(nparts, ms_per_step)
        line["parity_checked"] = bool(parity and parity.get("ok"))
        line["parity"] = parity
        line["perturbed_between_steps"] = bool(args.perturb)
    if world == 1 and not args.no_extras and not args.nparts:
        try:
            line["configs"] = extras_single_gpu(args, rk, torch, dev, tree, dsh, nparts, fp32_peak, flush)
        except Exception as e:  # the headline line must survive a failure of the additional blocks
            line["configs"] = {"error": repr(e)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_run(args, nparts, 4, 1, args.cpu_seconds)
            line["cpu_baseline"] = {"value": r["value"], "unit": "Ginteractions/s", "cores": r["cores"],
                                    "kind": r["kind"], "sample": r["sample"]}
        except Exception as e:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": "Ginteractions/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {e}"}
        try:
            line["reference_cuda_backend"] = reference_cuda_run(args, nparts)
            line["reference_cuda_backend"]["ours_same_call_ms"] = line["e2e"]["ms_per_step"]
        except Exception as e:
            line["reference_cuda_backend"] = {"error": repr(e)}
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
