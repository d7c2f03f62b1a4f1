"""N > 1 host logic on CPU: two gloo ranks agree on cost-weighted Morton-range cuts (no GPU)."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rakau_b200 import sharding
    rng = np.random.default_rng(0)  # same tree on every rank
    C, nparts = 5000, 200000
    sizes = rng.integers(1, 80, size=C)
    crit_begin = np.concatenate([[0], np.cumsum(sizes)[:-1]]) * nparts // sizes.sum()
    true_cost = (rng.pareto(1.5, size=C) * 1000 + 10).astype(np.int64)  # heavy-tailed, like a Plummer core
    cuts0 = sharding.cuts_by_particles(crit_begin, nparts, world)
    local = np.zeros(C, dtype=np.int64)
    local[cuts0[rank]:cuts0[rank + 1]] = true_cost[cuts0[rank]:cuts0[rank + 1]]  # this rank's evaluation
    summed = sharding.allreduce_costs(local, dist)
    cuts1 = sharding.cuts_by_cost(summed, world)
    q.put((rank, cuts0, cuts1, bool((summed == true_cost).all()), sharding.imbalance(true_cost, cuts0),
           sharding.imbalance(true_cost, cuts1)))
    dist.destroy_process_group()


def test_two_ranks_agree_on_cost_weighted_cuts():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, c0a, c1a, oka, ib0, ib1), (_, c0b, c1b, okb, _, _) = res
    assert oka and okb                      # the all-reduce reassembles the full cost vector
    assert c0a == c0b and c1a == c1b        # identical cuts on both ranks
    assert c1a[0] == 0 and c1a[-1] == 5000 and c1a[1] > 0
    assert ib1 <= 1.02 and ib1 <= ib0 + 1e-9  # cost-weighted cuts balance better than particle counts


def test_cut_helpers_edge_cases():
    from rakau_b200 import sharding
    assert sharding.cuts_by_particles([0], 10, 4) == [0, 1, 1, 1, 1]
    assert sharding.cuts_by_cost([5], 3) == [0, 1, 1, 1]
    assert sharding.cuts_by_cost([0, 0, 0, 0], 2) == [0, 2, 4]
    c = sharding.cuts_by_cost(np.ones(100), 4)
    assert c == [0, 25, 50, 75, 100]
    c = sharding.cuts_by_particles(np.arange(0, 1000, 10), 1000, 4)
    assert c == [0, 25, 50, 75, 100]
