"""Pins the oracle (oracle/rakau_oracle.cpp) against the known-answer vectors of the reference's own tests
(SURVEY §8c). CPU only."""
import numpy as np
import pytest

FPS = (32, 64)
MACS = ("bh", "bh_geom")


def test_morton_round_trip(oracle_mod):
    # reference test/morton.cpp:35-58: 10000 random triples, encode -> decode
    rng = np.random.default_rng(0)
    xyz = rng.integers(0, 1 << 21, size=(10000, 3))
    for x, y, z in xyz:
        c = oracle_mod.morton_encode(x, y, z)
        assert c < (1 << 63)
        assert oracle_mod.morton_decode(c) == [x, y, z]
    # bit layout of libmorton m3D_e_sLUT: x -> bit 0, y -> bit 1, z -> bit 2
    assert oracle_mod.morton_encode(1, 0, 0) == 1
    assert oracle_mod.morton_encode(0, 1, 0) == 2
    assert oracle_mod.morton_encode(0, 0, 1) == 4
    assert oracle_mod.morton_encode((1 << 21) - 1, (1 << 21) - 1, (1 << 21) - 1) == (1 << 63) - 1


@pytest.mark.parametrize("fp", FPS)
@pytest.mark.parametrize("mac", MACS)
def test_node_centre_kat(oracle_mod, fp, mac):
    # reference test/node_centre.cpp:55-110: 8 particles at (+-1,+-1,+-1), max_leaf_n=1, box 10 -> 9 nodes,
    # children ordered x fastest, then y, then z, centres at +-box/4
    x = np.array([1, 1, 1, 1, -1, -1, -1, -1.0])
    y = np.array([1, 1, -1, -1, 1, 1, -1, -1.0])
    z = np.array([1, -1, 1, -1, 1, -1, 1, -1.0])
    t = oracle_mod.OracleTree(x, y, z, np.ones(8), box_size=10, max_leaf_n=1, ncrit=1, fp=fp, mac=mac)
    nd = t.nodes()
    assert len(nd) == 9
    assert nd[0]["code"] == 1 and nd[0]["n_children"] == 8 and nd[0]["begin"] == 0 and nd[0]["end"] == 8
    assert np.allclose(t.node_centre(1), 0, atol=10 * np.finfo(t.F).eps)
    exp = [(-1, -1, -1), (1, -1, -1), (-1, 1, -1), (1, 1, -1), (-1, -1, 1), (1, -1, 1), (-1, 1, 1), (1, 1, 1)]
    for i, e in enumerate(exp):
        n = nd[i + 1]
        assert n["code"] == 8 + i and n["level"] == 1 and n["n_children"] == 0
        assert tuple(n["props"][:3]) == e and n["props"][3] == 1
        assert np.allclose(t.node_centre(n["code"]), 2.5 * np.array(e), atol=10 * np.finfo(t.F).eps * 2.5)


@pytest.mark.parametrize("fp", FPS)
def test_box_deduction_kat(oracle_mod, fp):
    F = oracle_mod.FDT[fp]
    # test/basic.cpp:144-147: coords +-10 -> 21 exactly
    c = np.array([-10, 1, 2, 10.0])
    t = oracle_mod.OracleTree(c, c, c, np.ones(4), fp=fp)
    assert F(t.box_size) == F(21)
    # test/auto_box_size.cpp:34-56
    t = oracle_mod.OracleTree([0, 1, 2, 3], [-4, -5, -6, -7], [4, 5, 3, 1], np.ones(4), max_leaf_n=1, ncrit=1, fp=fp)
    assert F(t.box_size) == F(14) + F(0.7)
    x, y, z, _ = t.parts()
    t.update_positions(x * 2, y * 2, z * 2)
    assert F(t.box_size) == F(28) + F(1.4)
    x, y, z, _ = t.parts()
    t.update_positions(x / 4, y / 4, z / 4)
    assert F(t.box_size) == F(7) + F(0.35)
    # ordered view after the updates (auto_box_size.cpp:57-72)
    x, y, z, _ = t.parts()
    ip = t.perm(2)
    assert list(x[ip]) == [0, 0.5, 1, 1.5]
    assert list(y[ip]) == [-2, -2.5, -3, -3.5]
    assert list(z[ip]) == [2, 2.5, 1.5, 0.5]


@pytest.mark.parametrize("fp", FPS)
def test_codes_sorted_and_reencode(oracle_mod, fp):
    # test/basic.cpp:322-354: 10000 uniform particles, box 1.25: codes sorted, c_it_o[i] == encode(disc(orig_i))
    F = oracle_mod.FDT[fp]
    m, x, y, z = oracle_mod.Rng(0).uniform_particles(10000, 1.0, fp=fp)
    t = oracle_mod.OracleTree(x, y, z, m, box_size=1.25, fp=fp)
    codes = t.codes()
    assert (codes[1:] >= codes[:-1]).all()
    inv = t.perm(2)
    inv_box = F(1) / F(1.25)

    def disc(v):
        # disc_single_coord with a real FMA: exact product+sum in float64/longdouble then one rounding
        if fp == 32:
            tmp = (v.astype(np.float64) * np.float64(inv_box) + 0.5).astype(np.float32)
        else:
            tmp = (v.astype(np.longdouble) * np.longdouble(inv_box) + np.longdouble(0.5)).astype(np.float64)
        tmp = tmp * F(1 << 21)
        return tmp.astype(np.uint64)
    dx, dy, dz = disc(x), disc(y), disc(z)
    for i in range(0, 10000, 7):
        assert codes[inv[i]] == oracle_mod.morton_encode(dx[i], dy[i], dz[i])


@pytest.mark.parametrize("fp", FPS)
def test_constructor_errors(oracle_mod, fp):
    # test/basic.cpp:178-202 message substrings
    c = np.array([-10, 1, 2, 10.0])
    m = np.ones(4)

    def err(**kw):
        with pytest.raises(oracle_mod.OracleError) as e:
            oracle_mod.OracleTree(c, c, c, m, fp=fp, **kw)
        return str(e.value)
    L = oracle_mod.lib()
    # explicit box of zero cannot be requested through the python wrapper's "0 = deduce" convention: go raw
    h = L.orc_create(fp, 0)
    arr = np.ascontiguousarray(c, dtype=oracle_mod.FDT[fp])
    mm = np.ascontiguousarray(m, dtype=oracle_mod.FDT[fp])
    p = lambda a: a.ctypes.data
    assert L.orc_build(h, p(arr), p(arr), p(arr), p(mm), 4, 0.0, 0, 4, 5) == 1
    assert b"While trying to discretise the input coordinate" in L.orc_last_error(h)
    L.orc_destroy(h)
    assert "produced the floating-point value" in err(box_size=3, max_leaf_n=4, ncrit=5)
    assert "The box size must be a finite non-negative value, but it is" in err(box_size=-3)
    assert "The box size must be a finite non-negative value, but it is" in err(box_size=np.inf)
    assert "The maximum number of particles per leaf must be nonzero" in err(max_leaf_n=0, ncrit=5)
    assert "The critical number of particles for the vectorised computation of the" in err(max_leaf_n=4, ncrit=0)


@pytest.mark.parametrize("fp", FPS)
def test_update_noop_and_swap(oracle_mod, fp):
    # test/update.cpp:65-155
    m, x, y, z = oracle_mod.Rng(1).uniform_particles(10000, 1.0, fp=fp)
    t = oracle_mod.OracleTree(x, y, z, m, box_size=10, fp=fp)
    perm0, parts0 = t.perm(0), t.parts()
    t.update_positions()  # no-op
    assert (t.perm(0) == perm0).all()
    assert (t.perm(1) == np.arange(10000)).all()
    for a, b in zip(t.parts(), parts0):
        assert (a == b).all()
    # rotate coordinates: positions follow through last_perm
    px, py, pz, pm = t.parts()
    t.update_positions(py, pz, px)
    lp = t.perm(1)
    nx, ny, nz, nm = t.parts()
    assert (nx == py[lp]).all() and (ny == pz[lp]).all() and (nz == px[lp]).all() and (nm == pm[lp]).all()
    assert (t.perm(0) == perm0[lp]).all()
    inv = t.perm(2)
    assert (t.perm(0)[inv] == np.arange(10000)).all()
    # ordered view returns the original masses
    assert (nm[inv] == np.asarray(m, dtype=t.F)).all()


@pytest.mark.parametrize("fp", FPS)
@pytest.mark.parametrize("mac", MACS)
def test_update_masses(oracle_mod, fp, mac):
    # test/update_masses.cpp:50-94
    m, x, y, z = oracle_mod.Rng(2).uniform_particles(10000, 1.0, fp=fp)
    t = oracle_mod.OracleTree(x, y, z, m, box_size=10, fp=fp, mac=mac)
    n0 = t.nodes()
    pm = t.parts()[3]
    t.update_masses(pm)
    assert (t.nodes() == n0).all()
    t.update_masses(pm * 2)
    n1 = t.nodes()
    assert (n1["props"][:, 3] == n0["props"][:, 3] * 2).all()
    assert (n1["props"][:, :3] == n0["props"][:, :3]).all()
    t.update_masses(pm * 0)
    n2 = t.nodes()
    assert (n2["props"][:, 3] == 0).all()
    for i in range(0, len(n2), 17):
        assert (n2["props"][i, :3] == t.node_centre(n2["code"][i]).astype(t.F)).all()
    with pytest.raises(oracle_mod.OracleError):
        bad = pm.copy()
        bad[5] = np.inf
        t.update_masses(bad)
    assert t.nparts == 0  # exception => tree cleared (update_masses.cpp:134-139)


@pytest.mark.parametrize("mac", MACS)
def test_accuracy_double_vs_exact(oracle_mod, mac):
    # test/accuracy_acc.cpp:111-118 (acc < 5e-10), accuracy_pot.cpp:96 (pot < 1e-10): theta = 0.001
    for N, mln, nc in ((10, 1, 1), (100, 2, 16), (1000, 8, 128), (2000, 16, 256)):
        m, x, y, z = oracle_mod.Rng(3).uniform_particles(N, 1.0, fp=64)
        t = oracle_mod.OracleTree(x, y, z, m, box_size=10, max_leaf_n=mln, ncrit=nc, fp=64, mac=mac)
        out, _ = t.acc_pot(2, 0.001)
        for i in range(0, N, max(1, N // 50)):
            e = t.exact(i)
            for j in range(3):
                assert abs((out[j][i] - e[j]) / e[j]) < 5e-10
            assert abs((out[3][i] - e[3]) / e[3]) < 1e-10


@pytest.mark.parametrize("fp", FPS)
def test_g_constant_and_zero_masses(oracle_mod, fp):
    # test/g_constant_acc.cpp:65-88: G=0 -> 0; accs(G=2) == 2*accs(G=1) bit-exact; zero_masses.cpp:53-74
    m, x, y, z = oracle_mod.Rng(4).uniform_particles(10000, 1.0, fp=fp)
    t = oracle_mod.OracleTree(x, y, z, m, box_size=10, fp=fp)
    a1, _ = t.acc_pot(0, 0.75)
    a0, _ = t.acc_pot(0, 0.75, G=0)
    a2, _ = t.acc_pot(0, 0.75, G=2)
    for j in range(3):
        assert (a0[j] == 0).all()
        assert (a2[j] == a1[j] * 2).all()
    t = oracle_mod.OracleTree(x, y, z, m * 0, box_size=10, fp=fp)
    out, _ = t.acc_pot(2, 0.75)
    for o in out:
        assert np.isfinite(o).all() and (o == 0).all()


def test_domain_errors(oracle_mod):
    m, x, y, z = oracle_mod.Rng(5).uniform_particles(100, 1.0)
    t = oracle_mod.OracleTree(x, y, z, m, box_size=10)
    for kw, msg in ((dict(theta=0.0), "The MAC value must be finite and positive"),
                    (dict(theta=0.5, eps=-1.0), "The softening length must be finite and non-negative"),
                    (dict(theta=0.5, G=np.inf), "The value of the gravitational constant G must be finite")):
        with pytest.raises(oracle_mod.OracleError) as e:
            t.acc_pot(0, **kw)
        assert e.value.code == 2 and msg in str(e.value)


def test_plummer_generator_is_deterministic(oracle_mod):
    # benchmark/common.hpp:96-126, std::mt19937 default seed: first values pinned so that a libstdc++ change
    # that alters the stream is caught (the GPU box ships the same image).
    m, x, y, z = oracle_mod.plummer(1000)
    m2, x2, y2, z2 = oracle_mod.plummer(1000)
    assert (m == m2).all() and (x == x2).all()
    assert m.min() >= 0.1 and m.max() < 1.9
    r = np.sqrt(x.astype(np.float64) ** 2 + y ** 2 + z ** 2)
    assert 0.5 < np.median(r) < 2.5  # Plummer half-mass radius ~1.3a
