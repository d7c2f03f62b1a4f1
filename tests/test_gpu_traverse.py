"""GPU traversal vs the oracle through the C ABI (T1/T3 of SURVEY §8c).

Tolerances (north_star): fp32 accelerations/potentials on the same tree and theta: median relative error
<= 1e-6, max <= 1e-4; fp64 <= 1e-12; interaction counts equal."""
import numpy as np
import pytest

from gpu_util import build_pair, rel_err_vec

pytestmark = pytest.mark.gpu

FP32_MEDIAN, FP32_MAX, FP64_MAX = 1e-6, 1e-4, 1e-12
COUNT_KEYS = ("mac_tests", "accepted", "p2p_pairs", "self_pairs", "interactions")


def _errors(go, oo, Q):
    errs = []
    if Q != 1:
        errs.append(rel_err_vec(go[:3], oo[:3]))
    if Q != 0:
        errs.append(np.abs(go[-1].astype(np.float64) - oo[-1]) / np.abs(oo[-1].astype(np.float64)))
    return errs


def check_eval(o, g, Q, theta, fp, rk=None, **kw):
    """T1: the GPU traversal on the ORACLE's node array (rk_traverse_external_tree, the literal drop-in for the
    reference's cuda_acc_pot_impl) must reproduce the interaction counts exactly and the values within the
    north_star tolerance. T3: on the GPU-built tree (node masses/COMs rounded differently, see
    gpu_util.assert_same_tree) a MAC decision within an ulp of its threshold may flip, so counts agree to 1e-5
    relative and the max-error bound applies to all but 0.1 % of the particles."""
    import rakau_b200
    oo, cnt = o.acc_pot(Q, theta, nthreads=8, **kw)
    F = o.F
    mac = "bh" if g.mac == 0 else "bh_geom"
    mac_value = F(1) / (F(theta) * F(theta)) if mac == "bh" else F(1) / F(theta)
    eps = F(kw.get("eps", 0.0))
    eo, einfo = rakau_b200.traverse_external_tree(o.nodes(), o.parts(), o.codes(), Q, mac_value, G=kw.get("G", 1.0),
                                                  eps2=eps * eps, mac=mac, fp=fp, ncrit=o_ncrit(o))
    for k in COUNT_KEYS:
        assert einfo[k] == cnt[k], (k, einfo[k], cnt[k])
    for e in _errors(eo, oo, Q):
        assert np.isfinite(e).all()
        if fp == 32:
            assert np.median(e) <= FP32_MEDIAN, np.median(e)
            assert e.max() <= FP32_MAX, e.max()
        else:
            assert e.max() <= FP64_MAX, e.max()
    go = g.acc_pot(Q, theta, **kw)
    ei = g.eval_info.asdict()
    for k in COUNT_KEYS:
        assert abs(ei[k] - cnt[k]) <= 1e-5 * cnt[k] + 2, (k, ei[k], cnt[k])
    for e in _errors(go, oo, Q):
        assert np.isfinite(e).all()
        if fp == 32:
            assert np.median(e) <= FP32_MEDIAN, np.median(e)
            assert np.quantile(e, 0.999) <= FP32_MAX, np.quantile(e, 0.999)
        else:
            assert np.quantile(e, 0.999) <= FP64_MAX, np.quantile(e, 0.999)
    return go, oo


def o_ncrit(o):
    return o._ncrit


@pytest.mark.parametrize("Q", [0, 1, 2])
@pytest.mark.parametrize("mac", ["bh", "bh_geom"])
def test_fp32_plummer(oracle_mod, rk, Q, mac):
    m, x, y, z = oracle_mod.plummer(50000)
    o, g = build_pair(oracle_mod, rk, x, y, z, m, mac=mac)
    check_eval(o, g, Q, 0.75, 32)


@pytest.mark.parametrize("Q", [0, 2])
@pytest.mark.parametrize("mac", ["bh", "bh_geom"])
def test_fp64_plummer(oracle_mod, rk, Q, mac):
    m, x, y, z = oracle_mod.plummer(30000, fp=64)
    o, g = build_pair(oracle_mod, rk, x, y, z, m, fp=64, mac=mac)
    check_eval(o, g, Q, 0.5, 64)


@pytest.mark.parametrize("mln,nc", [(1, 1), (2, 16), (8, 128), (16, 256), (16, 40), (4, 2), (64, 600)])
def test_group_shapes(oracle_mod, rk, mln, nc):
    """max_leaf_n / ncrit sweep of test/accuracy_acc.cpp, plus ncrit < max_leaf_n and groups beyond one pass."""
    m, x, y, z = oracle_mod.Rng(11).uniform_particles(8000, 1.0)
    o, g = build_pair(oracle_mod, rk, x, y, z, m, box_size=1.5, max_leaf_n=mln, ncrit=nc)
    check_eval(o, g, 2, 0.6, 32)


def test_config2_softening_and_G(oracle_mod, rk):
    # BASELINE config 2 at an oracle-sized N: accs+pots, eps = 0.01, G != 1
    m, x, y, z = oracle_mod.plummer(100000)
    o, g = build_pair(oracle_mod, rk, x, y, z, m)
    check_eval(o, g, 2, 0.75, 32, eps=0.01, G=2.5)


def test_huge_group_at_max_depth(oracle_mod, rk):
    """A level-21 leaf with 700 coincident particles is a critical node of 700 targets (> 256 per pass)."""
    rng = np.random.default_rng(3)
    N = 4000
    x, y, z = (rng.uniform(-1, 1, size=N).astype(np.float32) for _ in range(3))
    x[:700], y[:700], z[:700] = x[0], y[0], z[0]
    o, g = build_pair(oracle_mod, rk, x, y, z, np.ones(N), box_size=4.0)
    assert g.build_info.max_group >= 700
    check_eval(o, g, 2, 0.75, 32, eps=0.05)


@pytest.mark.parametrize("fp", [32, 64])
def test_accuracy_vs_direct_sum(oracle_mod, rk, fp):
    """test/accuracy_acc.cpp / accuracy_pot.cpp: theta = 0.001 against exact_*; double < 5e-10 (acc), 1e-10 (pot)."""
    for N, mln, nc in ((10, 1, 1), (100, 2, 16), (1000, 8, 128), (2000, 16, 256)):
        m, x, y, z = oracle_mod.Rng(3).uniform_particles(N, 1.0, fp=fp)
        g = rk.Octree(fp=fp)
        g.build(x, y, z, m, box_size=10.0, max_leaf_n=mln, ncrit=nc)
        out = g.acc_pot(2, 0.001)
        for o in out:
            assert np.isfinite(o).all()
        for i in range(0, N, max(1, N // 40)):
            e = g.exact(i)
            if fp == 64:
                for j in range(3):
                    assert abs((out[j][i] - e[j]) / e[j]) < 5e-10
                assert abs((out[3][i] - e[3]) / e[3]) < 1e-10
            else:
                a = np.array([out[0][i], out[1][i], out[2][i]], dtype=np.float64)
                assert np.linalg.norm(a - e[:3]) / np.linalg.norm(e[:3]) < 2e-3  # float bound of ordering_acc.cpp:91-98


def test_exact_matches_oracle(oracle_mod, rk):
    m, x, y, z = oracle_mod.plummer(20000, fp=64)
    o, g = build_pair(oracle_mod, rk, x, y, z, m, fp=64)
    for i in (0, 5, 19999):
        assert np.allclose(g.exact(i, G=1.5, eps=0.1), o.exact(i, G=1.5, eps=0.1), rtol=1e-12)
    inv = g.perm(2)
    assert np.allclose(g.exact(7, ordered=True), o.exact(int(inv[7])), rtol=1e-12)


@pytest.mark.parametrize("fp", [32, 64])
def test_g_linearity_bitexact_and_determinism(oracle_mod, rk, fp):
    # test/g_constant_acc.cpp:65-88 (N = 10000, theta = 0.75)
    m, x, y, z = oracle_mod.Rng(4).uniform_particles(10000, 1.0, fp=fp)
    g = rk.Octree(fp=fp)
    g.build(x, y, z, m, box_size=10.0)
    a1 = g.acc_pot(2, 0.75)
    a1b = g.acc_pot(2, 0.75)
    a0 = g.acc_pot(2, 0.75, G=0)
    a2 = g.acc_pot(2, 0.75, G=2)
    ah = g.acc_pot(2, 0.75, G=0.5, ordered=True)
    ao = g.acc_pot(2, 0.75, ordered=True)
    for j in range(4):
        assert (a1[j] == a1b[j]).all()  # run-to-run deterministic
        assert (a0[j] == 0).all()
        assert (a2[j] == a1[j] * 2).all()
        assert (ah[j] == ao[j] / 2).all()


def test_zero_masses(oracle_mod, rk):
    # test/zero_masses.cpp:53-74
    m, x, y, z = oracle_mod.Rng(5).uniform_particles(5000, 1.0)
    g = rk.Octree()
    g.build(x, y, z, m * 0, box_size=10.0)
    for o in g.acc_pot(2, 0.75):
        assert np.isfinite(o).all() and (o == 0).all()


def test_ordered_outputs_and_ranges(oracle_mod, rk):
    m, x, y, z = oracle_mod.plummer(60000)
    g = rk.Octree()
    g.build(x, y, z, m)
    au = g.acc_pot(2, 0.75, eps=0.01)
    full = g.eval_info.asdict()
    costs = g.group_costs()
    assert int(costs.sum()) == full["interactions"]
    ao = g.acc_pot(2, 0.75, eps=0.01, ordered=True)
    p = g.perm(0).astype(np.int64)
    for j in range(4):
        assert (ao[j][p] == au[j]).all()  # accs_o == accs_u scattered through perm (tree.hpp:3320-3330)
    # sharding by Morton range: two halves of the critical-node list reproduce the full result bit for bit
    C = g.ncrit_nodes
    cut = C // 3
    out = [np.full(g.nparts, np.nan, dtype=np.float32) for _ in range(4)]
    g.acc_pot(2, 0.75, eps=0.01, out=out, crit_range=(0, cut))
    n1 = g.eval_info.interactions
    g.acc_pot(2, 0.75, eps=0.01, out=out, crit_range=(cut, C))
    n2 = g.eval_info.interactions
    assert n1 + n2 == full["interactions"]
    for j in range(4):
        assert (out[j] == au[j]).all()


def test_split_validation(oracle_mod, rk):
    # tree.hpp:2857-2868; the CPU share folds into the GPU (documented deviation)
    m, x, y, z = oracle_mod.plummer(5000)
    g = rk.Octree()
    g.build(x, y, z, m)
    ref = g.acc_pot(0, 0.75)
    for bad, msg in (([0.5, np.nan], "cannot contain non-finite"), ([0.5, -1.0], "only non-negative"),
                     ([0.0, 0.0], "cannot all be zero")):
        with pytest.raises(rk.RakauError) as e:
            g.acc_pot(0, 0.75, split=bad)
        assert e.value.status == 1 and msg in str(e.value)
    got = g.acc_pot(0, 0.75, split=[0.5, 0.5])
    for j in range(3):
        assert (got[j] == ref[j]).all()
    with pytest.raises(rk.RakauError) as e:
        g.acc_pot(0, 0.75, split=[0.1] * 40)
    assert "accelerators, but only" in str(e.value)


def test_domain_errors(oracle_mod, rk):
    m, x, y, z = oracle_mod.plummer(100)
    g = rk.Octree()
    g.build(x, y, z, m)
    for kw, msg in ((dict(theta=0.0), "The MAC value must be finite and positive"),
                    (dict(theta=np.inf), "The MAC value must be finite and positive"),
                    (dict(theta=0.5, eps=-1.0), "The softening length must be finite and non-negative"),
                    (dict(theta=0.5, G=np.inf), "The value of the gravitational constant G must be finite")):
        with pytest.raises(rk.RakauError) as e:
            g.acc_pot(0, **kw)
        assert e.value.status == 2 and msg in str(e.value)


def test_duplicates_finite_with_softening(oracle_mod, rk):
    # test/softening_acc.cpp:115-146: coincident particles stay finite when eps != 0
    m, x, y, z = oracle_mod.Rng(6).uniform_particles(2000, 1.0)
    x[500:520], y[500:520], z[500:520] = x[0], y[0], z[0]
    for mln, nc in ((1, 1), (16, 128)):
        g = rk.Octree()
        g.build(x, y, z, m, box_size=10.0, max_leaf_n=mln, ncrit=nc)
        for o in g.acc_pot(2, 0.75, eps=0.1):
            assert np.isfinite(o).all()


def test_median_error_vs_direct_sum(oracle_mod, rk):
    """T3: the GPU's median error against direct summation is not worse than the oracle's own
    (test/median_error_acc.cpp prints but pins nothing; the oracle's values are the reference here)."""
    m, x, y, z = oracle_mod.plummer(5000)
    o, g = build_pair(oracle_mod, rk, x, y, z, m)
    ex = np.array([o.exact(i)[:3] for i in range(5000)])
    for theta in (0.2, 0.4, 0.6, 0.75, 0.8):
        oa, _ = o.acc_pot(0, theta)
        ga = g.acc_pot(0, theta)
        eo = np.median(np.linalg.norm(np.stack(oa, 1) - ex, axis=1) / np.linalg.norm(ex, axis=1))
        eg = np.median(np.linalg.norm(np.stack(ga, 1) - ex, axis=1) / np.linalg.norm(ex, axis=1))
        assert eg <= eo * 1.05 + 1e-6, (theta, eg, eo)


def test_external_tree_split_offset(oracle_mod, rk):
    """cuda_acc_pot_impl semantics (src/rakau_cuda.cu:348-528): only particles [split_indices[0], nparts) are
    evaluated, written at offset 0 (offset_output = false) or at their own index (true)."""
    m, x, y, z = oracle_mod.plummer(20000)
    o = oracle_mod.OracleTree(x, y, z, m)
    oo, _ = o.acc_pot(2, 0.75, eps=0.01)
    cr, _ = o.crit()
    first = int(cr[len(cr) // 2, 1])
    mv = np.float32(1) / (np.float32(0.75) * np.float32(0.75))
    e2 = np.float32(0.01) * np.float32(0.01)
    a, _ = rk.traverse_external_tree(o.nodes(), o.parts(), o.codes(), 2, mv, eps2=e2, first=first, offset_output=True)
    b, _ = rk.traverse_external_tree(o.nodes(), o.parts(), o.codes(), 2, mv, eps2=e2, first=first, offset_output=False)
    for j in range(4):
        assert (a[j][:first] == 0).all()
        assert (a[j][first:] == b[j]).all()
        rel = np.abs(a[j][first:].astype(np.float64) - oo[j][first:]) / np.abs(oo[j][first:].astype(np.float64))
        assert np.median(rel) <= 1e-6
    with pytest.raises(rk.RakauError):
        rk.traverse_external_tree(o.nodes(), o.parts(), o.codes(), 2, mv, first=first + 1)


def test_pipelined_host_path_matches_device_path(oracle_mod, rk):
    """n >= 2^20 with host buffers takes the pipelined route (masses uploaded underneath the sort, four traversal
    launches with overlapped device-to-host copies): results, permutation and counters must be bit-identical to
    the single-launch device-buffer route, and the masses must land on the right particles."""
    import torch
    n = (1 << 20) + 12345
    m, x, y, z = oracle_mod.plummer(n)
    m = (m * (1.0 + np.arange(n, dtype=np.float32) / n)).astype(np.float32)  # distinct masses
    gh = rk.Octree()
    gh.build(x, y, z, m)
    host = gh.acc_pot(2, 0.75, eps=0.001)
    ih = gh.eval_info.asdict()
    assert ih["kernel_launches"] == 4
    gd = rk.Octree()
    dev = [torch.from_numpy(a).cuda() for a in (x, y, z, m)]
    gd.build(*dev, where=rk.RK_DEVICE, n=n)
    out = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(4)]
    gd.acc_pot(2, 0.75, eps=0.001, out=out, where=rk.RK_DEVICE)
    idv = gd.eval_info.asdict()
    assert idv["kernel_launches"] == 1
    torch.cuda.synchronize()
    for k in COUNT_KEYS:
        assert ih[k] == idv[k], k
    assert (gh.perm(0) == gd.perm(0)).all()
    for a, b in zip(gh.parts(), gd.parts()):
        assert (a == b).all()
    assert (gh.parts()[3] == m[gh.perm(0).astype(np.int64)]).all()
    for j in range(4):
        assert (host[j] == out[j].cpu().numpy()).all(), j
    # a ranged host evaluation large enough to be pipelined as well
    C = gh.ncrit_nodes
    part = [np.full(n, np.nan, dtype=np.float32) for _ in range(4)]
    gh.acc_pot(2, 0.75, eps=0.001, out=part, crit_range=(3, C))
    lo = int(gh.crit_begin_at([3])[0])
    for j in range(4):
        assert (part[j][lo:] == host[j][lo:]).all() and np.isnan(part[j][:lo]).all()


@pytest.mark.parametrize("Q", [0, 1, 2])
def test_pinned_host_outputs_are_written_by_the_kernel(oracle_mod, rk, Q):
    """Unordered outputs in PINNED host memory: the kernel writes the final results straight into the caller's buffers
    (one launch, no device-to-host copy; the partial sums of phase 1 stay in device memory). Bit-identical to the
    copied (pageable-buffer) route, for a full evaluation, for a range of critical nodes (nothing outside the range is
    touched) and with the option switched off."""
    import torch
    n = (1 << 20) + 777
    m, x, y, z = oracle_mod.plummer(n)
    g = rk.Octree()
    g.build(x, y, z, m)
    nres = {0: 3, 1: 1, 2: 4}[Q]
    ref = g.acc_pot(Q, 0.75, G=1.5, eps=0.001)  # numpy arrays: pageable, copied
    assert g.eval_info.kernel_launches == 4
    pinned = [torch.full((n,), float("nan"), dtype=torch.float32).pin_memory() for _ in range(nres)]
    pn = [t.numpy() for t in pinned]
    g.acc_pot(Q, 0.75, G=1.5, eps=0.001, out=pn)
    assert g.eval_info.kernel_launches == 1
    for j in range(nres):
        assert (pn[j] == ref[j]).all(), j
    C = g.ncrit_nodes
    lo, hi = (int(v) for v in g.crit_begin_at([5, C - 7]))
    for a in pn:
        a[:] = np.nan
    g.acc_pot(Q, 0.75, G=1.5, eps=0.001, out=pn, crit_range=(5, C - 7))
    for j in range(nres):
        assert (pn[j][lo:hi] == ref[j][lo:hi]).all() and np.isnan(pn[j][:lo]).all() and np.isnan(pn[j][hi:]).all()
    g.set_option("zero_copy_out", 0)
    g.acc_pot(Q, 0.75, G=1.5, eps=0.001, out=pn)
    assert g.eval_info.kernel_launches == 4
    for j in range(nres):
        assert (pn[j] == ref[j]).all(), j


def test_output_mirrors_receive_every_result(oracle_mod, rk):
    """rk_tree_set_output_mirrors: the kernel stores each final result into the mirrors as well (on one GPU: two device
    buffers and a pinned host buffer shifted by the first particle of the range, as the multi-GPU exchange uses
    them). Mirrors equal the outputs inside the evaluated range and stay untouched outside it."""
    import torch
    n = 300000
    m, x, y, z = oracle_mod.plummer(n)
    g = rk.Octree()
    g.build(x, y, z, m)
    C = g.ncrit_nodes
    c0, c1 = 11, C - 5
    lo, hi = (int(v) for v in g.crit_begin_at([c0, c1]))
    out = [torch.full((n,), float("nan"), device="cuda") for _ in range(4)]
    mir = [[torch.full((n,), float("nan"), device="cuda") for _ in range(4)] for _ in range(2)]
    host = [torch.full((hi - lo,), float("nan")).pin_memory() for _ in range(4)]
    g.set_output_mirrors([[t.data_ptr() for t in mir[0]], [t.data_ptr() for t in mir[1]],
                          [t.data_ptr() - 4 * lo for t in host]])
    g.acc_pot(2, 0.75, eps=0.001, out=out, where=rk.RK_DEVICE, crit_range=(c0, c1))
    g.set_output_mirrors([])
    torch.cuda.synchronize()
    for j in range(4):
        o = out[j].cpu().numpy()
        assert np.isfinite(o[lo:hi]).all() and np.isnan(o[:lo]).all() and np.isnan(o[hi:]).all()
        for r in range(2):
            mm = mir[r][j].cpu().numpy()
            assert (mm[lo:hi] == o[lo:hi]).all() and np.isnan(mm[:lo]).all() and np.isnan(mm[hi:]).all()
        assert (host[j].numpy() == o[lo:hi]).all()
    # switched off again: the mirrors keep their old contents
    for t in mir[0]:
        t.fill_(7.0)
    g.acc_pot(2, 0.75, eps=0.001, out=out, where=rk.RK_DEVICE)
    torch.cuda.synchronize()
    assert all(bool((t == 7.0).all()) for t in mir[0])
    with pytest.raises(rk.RakauError):
        g.set_output_mirrors([[0, 0, 0, 0]] * 9)


def test_bcast_copy(rk):
    """rk_device_bcast_copy: the same bytes to several destinations, odd 8-byte offsets and a tail that is not a
    multiple of 8 bytes included."""
    import torch
    src = torch.randint(0, 255, (8 * 100003 + 5,), dtype=torch.uint8, device="cuda")
    dst = [torch.zeros(8 * 100010 + 64, dtype=torch.uint8, device="cuda") for _ in range(5)]
    offs = [0, 8, 24, 8 * 7, 16]
    rk.device_bcast_copy([d.data_ptr() + o for d, o in zip(dst, offs)], src.data_ptr(), src.numel(),
                         torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for d, o in zip(dst, offs):
        assert bool((d[o:o + src.numel()] == src).all()) and int(d[:o].sum()) == 0 and int(d[o + src.numel():].sum()) == 0
    with pytest.raises(rk.RakauError):
        rk.device_bcast_copy([dst[0].data_ptr() + 4], src.data_ptr(), 64, 0)
