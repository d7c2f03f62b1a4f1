"""Host-side logic of the multi-GPU path: how the critical-node list is cut into contiguous Morton ranges.

The tree is replicated (every GPU builds the identical tree from the all-gathered particles), so ranks only
have to agree on the cuts. First evaluation: equal particle counts, snapped to critical-node boundaries as the
reference snaps its first split index (tree.hpp:3147-3178). Later evaluations: equal shares of the previous
evaluation's per-group interaction counts (BASELINE north_star: "cost-weighted splits taken from the previous
evaluation's interaction counts"). Every rank evaluates only its own range, so the per-group costs are
summed over ranks (all_reduce) before cutting; all ranks then compute the same cuts.
"""
import numpy as np


def cuts_by_particles(crit_begin, nparts, world):
    """crit_begin: first particle of each critical node (ascending). Returns world+1 critical-node indices."""
    crit_begin = np.asarray(crit_begin, dtype=np.int64)
    targets = (np.arange(1, world, dtype=np.int64) * nparts) // world
    cuts = np.searchsorted(crit_begin, targets, side="left")
    return [0] + [int(c) for c in cuts] + [int(crit_begin.size)]


def cuts_by_cost(costs, world):
    """costs: interactions per critical node (already summed over ranks). Returns world+1 indices such that
    every range carries ~1/world of the total cost."""
    cum = np.cumsum(np.asarray(costs, dtype=np.float64))
    total = cum[-1] if cum.size else 0.0
    if total <= 0:
        step = max(1, cum.size // world)
        return [min(cum.size, r * step) for r in range(world)] + [int(cum.size)]
    targets = total * np.arange(1, world, dtype=np.float64) / world
    cuts = np.searchsorted(cum, targets, side="left") + 1
    cuts = np.minimum(np.maximum.accumulate(cuts), cum.size)
    return [0] + [int(c) for c in cuts] + [int(cum.size)]


def allreduce_costs(local_costs, dist=None, device=None):
    """Sum the per-group costs over ranks (each rank holds zeros outside its own range)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(local_costs).astype(np.int64))
    if device is not None:
        t = t.to(device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t.cpu().numpy()


def imbalance(costs, cuts):
    """max share / mean share of the given cuts (1.0 = perfect)."""
    costs = np.asarray(costs, dtype=np.float64)
    shares = np.array([costs[cuts[r]:cuts[r + 1]].sum() for r in range(len(cuts) - 1)])
    return float(shares.max() / max(shares.mean(), 1e-300))
