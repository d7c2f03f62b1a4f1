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


def _sharded_worker(rank, world, port, q):
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cpu_backend import CpuOctree, morton_codes
    from rakau_b200 import deduce_box
    from rakau_b200.distributed import ShardedTree
    rng = np.random.default_rng(7)  # the full particle set, identical on every rank
    N = 50_000
    full = [rng.normal(size=N).astype(np.float32) for _ in range(3)] + [rng.uniform(0.1, 1.9, N).astype(np.float32)]
    full[0][:200] = full[0][200:400]  # duplicated coordinates => equal Morton codes: the stable order must survive
    full[1][:200] = full[1][200:400]
    full[2][:200] = full[2][200:400]
    if world == 2:
        first, cnt = (0, 30_000) if rank == 0 else (30_000, N - 30_000)  # uneven shards
    else:  # three ranks, the middle one holds nothing
        first, cnt = [(0, 35_000), (35_000, 0), (35_000, N - 35_000)][rank]
    shard = [torch.from_numpy(a[first:first + cnt].copy()) for a in full]
    st = ShardedTree(dist, torch.device("cpu"), fp=32, ncrit=96, samples_per_rank=64, octree_factory=CpuOctree)
    st.build(*shard, first_index=first)
    # reference: one stable sort of everything
    box = deduce_box(float(max(np.abs(a).max() for a in full[:3])), 32)
    codes = morton_codes(full[0], full[1], full[2], box)
    p = np.argsort(codes, kind="stable")
    ok_build = bool((st.tree._codes == codes[p]).all()) and bool((st.tree._perm == p).all())
    ok_build &= all(bool((a == b[p]).all()) for a, b in zip(st.tree._parts, full))
    # evaluation 1 (cuts by particle count), rebalance, evaluation 2 (cuts by cost): every rank ends with everything
    info1, out1 = st.acc_pot(0, 0.75)
    cuts1 = list(st.cuts)
    out1 = [o.clone() for o in out1]
    imb = st.rebalance(kernel_ms=1.0 + rank)
    cuts2 = list(st.cuts)
    mine = [torch.full((N,), float("nan")) for _ in range(3)]
    info2, out2 = st.acc_pot(0, 0.75, out=mine)
    sorted_parts = [a[p] for a in full]
    ok_eval = all(bool((o.numpy()[:N] == CpuOctree.expected(sorted_parts, j)).all()) for j, o in enumerate(out1))
    ok_eval &= all(bool((o.numpy() == CpuOctree.expected(sorted_parts, j)).all()) for j, o in enumerate(out2))
    # a rebuild changes the critical nodes: the cost-weighted cuts must survive as particle boundaries, re-snapped to
    # the new critical nodes (not fall back to equal particle counts)
    pidx2 = [int(v) for v in st.cut_particles]
    st.build(*shard, first_index=first)
    info3, out3 = st.acc_pot(0, 0.75, out=[torch.full((N,), float("nan")) for _ in range(3)])
    begins = np.append(st.tree._begin, N)
    ok_resnap = all(int(begins[c]) >= pb and (c == 0 or int(begins[c - 1]) < pb) for c, pb in zip(st.cuts[1:-1], pidx2[1:-1]))
    ok_resnap &= [int(v) for v in st.cut_pidx] == pidx2
    ok_eval &= all(bool((o.numpy() == CpuOctree.expected(sorted_parts, j)).all()) for j, o in enumerate(out3))
    # results in the ORIGINAL (global) particle order, as the reference's `_o` functions return them
    _, out4 = st.acc_pot(0, 0.75, out=[torch.full((N,), float("nan")) for _ in range(3)], ordered=True)
    ok_eval &= all(bool((o.numpy() == CpuOctree.expected(full, j)).all()) for j, o in enumerate(out4))
    tot = torch.tensor([info1["interactions"], info2["interactions"]], dtype=torch.int64)
    dist.all_reduce(tot)
    q.put((rank, ok_build, ok_eval and ok_resnap, cuts1, cuts2, imb, tot.tolist()))
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_tree_orchestration(world):
    """rakau_b200.distributed.ShardedTree over gloo with the numpy backend of tests/cpu_backend.py: the distributed
    sample sort reproduces the single stable sort (codes, permutation, particle order, ties included), all
    ranks end with the complete result of both evaluations, and they agree on the range cuts. world = 3: one
    rank starts with an empty shard."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _, ba, ea, c1a, c2a, ia, ta = res[0]
    for _, bb, eb, c1b, c2b, ib, tb in res:
        assert bb and eb
        assert c1a == c1b and c2a == c2b and ia == ib and ta == tb
    assert len(c1a) == world + 1 and c1a[0] == 0 and c1a[-1] == c2a[-1]
    assert c2a != c1a  # the time-weighted rebalance moved the cuts (higher ranks reported slower kernels)
    assert ta[0] == ta[1]  # every group evaluated exactly once per evaluation
