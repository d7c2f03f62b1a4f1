"""GPU build (Morton encode, radix sort, octree topology, node properties, critical nodes) vs the oracle,
through the C ABI. Bit-exact for codes / permutations / topology; stated tolerance for mass and COM
(the GPU reduces in fp64 with a fixed, mass-independent tree; the reference sums sequentially in F)."""
import numpy as np
import pytest

from gpu_util import assert_same_tree, build_pair

pytestmark = pytest.mark.gpu

PROPS_TOL = {32: 4 * 1.2e-7, 64: 4 * 2.3e-16}  # floor of the node-property bound (see gpu_util.assert_same_tree)


@pytest.mark.parametrize("N", [1, 2, 33, 1000, 20000])
@pytest.mark.parametrize("mln,nc", [(1, 1), (16, 128), (8, 4), (3, 1000)])
def test_plummer_fp32(oracle_mod, rk, N, mln, nc):
    m, x, y, z = oracle_mod.plummer(N)
    o, g = build_pair(oracle_mod, rk, x, y, z, m, max_leaf_n=mln, ncrit=nc)
    assert_same_tree(o, g, PROPS_TOL[32])


@pytest.mark.parametrize("fp", [32, 64])
@pytest.mark.parametrize("mac", ["bh", "bh_geom"])
def test_uniform_types_and_macs(oracle_mod, rk, fp, mac):
    m, x, y, z = oracle_mod.Rng(7).uniform_particles(30000, 1.0, fp=fp)
    o, g = build_pair(oracle_mod, rk, x, y, z, m, fp=fp, mac=mac, box_size=1.25)
    assert_same_tree(o, g, PROPS_TOL[fp])
    if mac == "bh_geom":
        gn, on = g.nodes(), o.nodes()
        cnt = (on["end"] - on["begin"]).astype(np.float64)
        assert (np.abs(gn["delta"].astype(np.float64) - on["delta"]) <= (cnt + 8) * np.finfo(g.F).eps * 4 * o.box_size).all()


def test_large_fp32(oracle_mod, rk):
    m, x, y, z = oracle_mod.plummer(400000)
    o, g = build_pair(oracle_mod, rk, x, y, z, m)
    assert_same_tree(o, g, PROPS_TOL[32])


def test_duplicates_and_ties(oracle_mod, rk):
    """Identical positions reach level 21 (leaf larger than max_leaf_n); equal codes keep input order (stable)."""
    rng = np.random.default_rng(1)
    N = 5000
    x, y, z = (rng.normal(size=N).astype(np.float32) for _ in range(3))
    x[:700], y[:700], z[:700] = x[0], y[0], z[0]
    x[900:1000] = x[900] + 1e-7
    for mln, nc in ((1, 1), (16, 128), (8, 600)):
        o, g = build_pair(oracle_mod, rk, x, y, z, np.ones(N), box_size=20.0, max_leaf_n=mln, ncrit=nc)
        assert_same_tree(o, g, PROPS_TOL[32])
        codes = g.codes()
        assert (codes[1:] == codes[:-1]).sum() >= 699
        p = g.perm(0).astype(np.int64)
        same = codes[1:] == codes[:-1]
        assert (p[1:][same] > p[:-1][same]).all()  # stability


def test_all_identical_particles(oracle_mod, rk):
    N = 300
    x = np.full(N, 0.25, dtype=np.float32)
    o, g = build_pair(oracle_mod, rk, x, x, x, np.ones(N), box_size=2.0)
    assert_same_tree(o, g, PROPS_TOL[32])
    assert (g.perm(0) == np.arange(N)).all()


def test_empty_tree(rk):
    g = rk.Octree()
    e = np.zeros(0, dtype=np.float32)
    g.build(e, e, e, e)
    assert g.nparts == 0 and g.nnodes == 0 and g.ncrit_nodes == 0
    assert g.acc_pot(0, 0.75)[0].size == 0


def test_kat_node_centre_and_box(oracle_mod, rk):
    # reference test/node_centre.cpp:55-110 and test/basic.cpp:144-147 through the GPU path
    x = np.array([1, 1, 1, 1, -1, -1, -1, -1.0])
    y = np.array([1, 1, -1, -1, 1, 1, -1, -1.0])
    z = np.array([1, -1, 1, -1, 1, -1, 1, -1.0])
    for fp in (32, 64):
        g = rk.Octree(fp=fp)
        g.build(x, y, z, np.ones(8), box_size=10, max_leaf_n=1, ncrit=1)
        nd = g.nodes()
        assert len(nd) == 9 and nd[0]["n_children"] == 8
        exp = [(-1, -1, -1), (1, -1, -1), (-1, 1, -1), (1, 1, -1), (-1, -1, 1), (1, -1, 1), (-1, 1, 1), (1, 1, 1)]
        for i, e in enumerate(exp):
            assert nd[i + 1]["code"] == 8 + i and tuple(nd[i + 1]["props"][:3]) == e
        c = np.array([-10, 1, 2, 10.0])
        g.build(c, c, c, np.ones(4))
        assert g.box_size == 21.0
        g.build([0, 1, 2, 3], [-4, -5, -6, -7], [4, 5, 3, 1], np.ones(4), max_leaf_n=1, ncrit=1)
        assert g.F(g.box_size) == g.F(14) + g.F(0.7)  # test/auto_box_size.cpp:34-44


@pytest.mark.parametrize("fp", [32, 64])
def test_constructor_errors(rk, fp):
    # message substrings of test/basic.cpp:178-202; status codes map to the reference's exception types
    c = np.array([-10, 1, 2, 10.0])
    m = np.ones(4)
    g = rk.Octree(fp=fp)

    def err(**kw):
        with pytest.raises(rk.RakauError) as e:
            g.build(c, c, c, m, **kw)
        assert e.value.status == 1
        assert g.nparts == 0
        return str(e.value)
    assert "produced the floating-point value" in err(box_size=3, max_leaf_n=4, ncrit=5)
    assert "The box size must be a finite non-negative value, but it is" in err(box_size=-3)
    assert "The box size must be a finite non-negative value, but it is" in err(box_size=np.inf)
    assert "The maximum number of particles per leaf must be nonzero" in err(max_leaf_n=0, ncrit=5)
    assert "The critical number of particles for the vectorised computation of the" in err(max_leaf_n=4, ncrit=0)
    bad = c.copy()
    bad[2] = np.nan
    with pytest.raises(rk.RakauError) as e:
        g.build(bad, c, c, m)
    assert "non-finite coordinate" in str(e.value)
    with pytest.raises(rk.RakauError) as e:
        g.build(bad, c, c, m, box_size=100.0)
    assert "While trying to discretise the input coordinate" in str(e.value)
    with pytest.raises(rk.RakauError) as e:
        g.build(c, c, c, np.array([1, np.inf, 1, 1.0]), box_size=100.0)
    assert "non-finite" in str(e.value)


def test_full_size_properties(oracle_mod, rk):
    """BASELINE config 1 size (4M): size-independent invariants instead of an oracle comparison."""
    N = 4_000_000
    m, x, y, z = oracle_mod.plummer(N)
    g = rk.Octree()
    g.build(x, y, z, m)
    codes = g.codes()
    assert (codes[1:] >= codes[:-1]).all()
    p = g.perm(0).astype(np.int64)
    assert (np.sort(p) == np.arange(N)).all()
    assert (g.perm(2)[p] == np.arange(N)).all()
    px, py, pz, pm = g.parts()
    assert (px == x[p]).all() and (pm == m[p]).all()
    nd = g.nodes()
    assert nd[0]["begin"] == 0 and nd[0]["end"] == N and nd[0]["n_children"] == len(nd) - 1
    assert ((nd["end"] - nd["begin"]) > 0).all()
    leaves = nd[nd["n_children"] == 0]
    assert (leaves["begin"][1:] == leaves["end"][:-1]).all() and leaves["end"][-1] == N  # leaves tile the particles
    assert ((leaves["end"] - leaves["begin"] <= 16) | (leaves["level"] == 21)).all()
    cr = g.crit()
    assert cr[0, 1] == 0 and cr[-1, 2] == N and (cr[1:, 1] == cr[:-1, 2]).all()
    assert abs(nd[0]["props"][3] / m.astype(np.float64).sum() - 1) < 1e-6
    # every node's code is the common prefix of its particles' codes
    lvl = nd["level"].astype(np.uint64)
    sh = (np.uint64(3) * (np.uint64(21) - lvl))
    first = codes[nd["begin"]] >> sh
    last = codes[nd["end"] - np.uint64(1)] >> sh
    want = nd["code"] - (np.uint64(1) << (np.uint64(3) * lvl))
    assert (first == want).all() and (last == want).all()


def test_sample_sort_building_blocks_single_device(oracle_mod, rk):
    """The multi-GPU build on ONE device: two shards are sorted with the global box
    (rk_tree_sort_shard), merged through splitter buckets on the host, each bucket is sorted again with the
    received codes, and rk_tree_build_presorted builds the tree from the concatenated buckets. The result must
    be identical to the direct build (codes, perm, nodes, critical nodes)."""
    import torch
    dev = torch.device("cuda", 0)
    N = 150000
    m, x, y, z = oracle_mod.plummer(N)
    ref = rk.Octree()
    ref.build(x, y, z, m)
    box = ref.box_size
    assert rk.deduce_box(float(max(np.abs(x).max(), np.abs(y).max(), np.abs(z).max())), 32) == box
    cut = 70000
    shards = []
    for lo, hi in ((0, cut), (cut, N)):
        t = rk.Octree()
        d = [torch.from_numpy(np.ascontiguousarray(a[lo:hi])).to(dev) for a in (x, y, z, m)]
        t.sort_shard(d[0], d[1], d[2], d[3], hi - lo, box)
        codes = torch.empty(hi - lo, dtype=torch.int64, device=dev)
        cols = [torch.empty(hi - lo, dtype=torch.float32, device=dev) for _ in range(4)]
        lp = torch.empty(hi - lo, dtype=torch.int32, device=dev)
        t.codes_device(codes)
        t.parts_device(*cols)
        t.perm_device(lp, rk.RK_LAST_PERM)
        shards.append((codes, cols, lp + lo))
    split = int(np.median(ref.codes()))
    buckets = []
    for b in range(2):
        sel = [(c < split) if b == 0 else (c >= split) for c, _, _ in shards]
        bc = torch.cat([c[s] for (c, _, _), s in zip(shards, sel)])
        bcols = [torch.cat([cols[j][s] for (_, cols, _), s in zip(shards, sel)]) for j in range(4)]
        bi = torch.cat([g[s] for (_, _, g), s in zip(shards, sel)])
        t = rk.Octree()
        t.sort_shard(bcols[0], bcols[1], bcols[2], bcols[3], bc.numel(), box, codes=bc)
        lp = torch.empty(bc.numel(), dtype=torch.int32, device=dev)
        t.perm_device(lp, rk.RK_LAST_PERM)
        sc = torch.empty_like(bc)
        scols = [torch.empty_like(c) for c in bcols]
        t.codes_device(sc)
        t.parts_device(*scols)
        buckets.append((sc, scols, bi[lp.long()]))
    fc = torch.cat([b[0] for b in buckets])
    fcols = [torch.cat([b[1][j] for b in buckets]) for j in range(4)]
    fi = torch.cat([b[2] for b in buckets]).to(torch.int32)
    g = rk.Octree()
    g.build_presorted(fcols[0], fcols[1], fcols[2], fcols[3], fc, fi, N, box)
    assert (g.codes() == ref.codes()).all()
    assert (g.perm(0) == ref.perm(0)).all()
    assert (g.perm(2) == ref.perm(2)).all()
    assert (g.nodes() == ref.nodes()).all()
    assert (g.crit() == ref.crit()).all()
    a, b = g.acc_pot(0, 0.75, ordered=True), ref.acc_pot(0, 0.75, ordered=True)
    for j in range(3):
        assert (a[j] == b[j]).all()
    assert (g.crit_begin_at([0, g.ncrit_nodes]) == [0, N]).all()


@pytest.mark.parametrize("fp,mac,n", [(32, "bh", 200000), (64, "bh_geom", 100000), (32, "bh_geom", 50000)])
@pytest.mark.parametrize("bottom_up", [0, 1])
def test_node_properties_both_reductions(oracle_mod, rk, fp, mac, n, bottom_up):
    """Node properties top-down (small trees) and bottom-up (large trees, one launch per level): both within the
    reference's own summation error of the oracle, and both with a mass-independent reduction tree (scaling every mass
    by a power of two scales every node mass exactly, test/update_masses.cpp:56-68)."""
    m, x, y, z = oracle_mod.plummer(n, fp=fp)
    o = oracle_mod.OracleTree(x, y, z, m, fp=fp, mac=mac)
    g = rk.Octree(fp=fp, mac=mac)
    g.set_option("props_bottom_up", bottom_up)
    g.build(x, y, z, m)
    assert_same_tree(o, g, 4 * float(np.finfo(o.F).eps))
    n1 = g.nodes()
    g.update_masses((g.parts()[3] * 4).astype(o.F))
    n2 = g.nodes()
    assert (n2["props"][:, 3] == n1["props"][:, 3] * 4).all()
    assert (n2["props"][:, :3] == n1["props"][:, :3]).all()


def test_digest_and_crit_lower_bound(oracle_mod, rk):
    m, x, y, z = oracle_mod.plummer(150000)
    a, b = rk.Octree(), rk.Octree()
    a.build(x, y, z, m)
    b.build(x, y, z, m)
    assert (a.digest() == b.digest()).all() and a.digest()[:7].all()
    x2 = x.copy()
    x2[777] += 1e-3
    b.build(x2, y, z, m)
    assert (a.digest() != b.digest()).any()
    cb = a.crit()[:, 1].astype(np.int64)
    q = np.array([0, 1, cb[5], cb[5] + 1, cb[-1], cb[-1] + 1, a.nparts], dtype=np.int64)
    assert (a.crit_lower_bound(q).astype(np.int64) == np.searchsorted(cb, q, side="left")).all()
    assert "traverse_kernel" not in a.last_kernel()
    a.acc_pot(0, 0.75)
    assert a.last_kernel().startswith("traverse_kernel<float,Q=0,MAC=0")


def test_partition_building_blocks_single_device(oracle_mod, rk):
    """rk_tree_encode_shard + rk_tree_partition_shard (the sample sort without a local pre-sort): codes equal the
    oracle's, every particle lands in the bucket of its code, buckets are contiguous and keep the input order (that is
    what makes the receiving rank's stable sort reproduce the single-GPU order), sizes are reported."""
    import torch
    n = 300000
    m, x, y, z = oracle_mod.plummer(n)
    x[100:140] = x[50]  # equal codes across the shard
    y[100:140] = y[50]
    z[100:140] = z[50]
    o = oracle_mod.OracleTree(x, y, z, m)
    box = o.box_size
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(a).to(dev) for a in (x, y, z, m)]
    t = rk.Octree()
    t.set_stream(torch.cuda.current_stream().cuda_stream)
    t.encode_shard(d[0], d[1], d[2], d[3], n, box)
    codes = torch.empty(n, dtype=torch.int64, device=dev)
    t.codes_device(codes)
    inv = np.empty(n, dtype=np.int64)
    inv[o.perm(0).astype(np.int64)] = np.arange(n)
    want_codes = o.codes().astype(np.int64)[inv]  # the oracle's codes in input order
    assert (codes.cpu().numpy() == want_codes).all()
    split = torch.from_numpy(np.sort(want_codes)[[n // 5, n // 2, n // 2 + 7, 9 * n // 10]].copy()).to(dev)
    cnt = t.partition_shard(split).astype(np.int64)
    bucket = np.searchsorted(split.cpu().numpy(), want_codes, side="right")
    assert (cnt == np.bincount(bucket, minlength=5)).all() and cnt.sum() == n
    lp = torch.empty(n, dtype=torch.int32, device=dev)
    t.perm_device(lp, rk.RK_LAST_PERM)
    lp = lp.cpu().numpy().astype(np.int64)
    assert (lp == np.argsort(bucket, kind="stable")).all()  # buckets contiguous, input order kept inside
    t.codes_device(codes)
    assert (codes.cpu().numpy() == want_codes[lp]).all()
    cols = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(4)]
    t.parts_device(*cols)
    for got, a in zip(cols, (x, y, z, m)):
        assert (got.cpu().numpy() == a[lp]).all()
