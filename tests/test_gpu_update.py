"""update_particles_u / update_masses_u paths (sync, tree.hpp:3678-3743; 3782-3805) on the GPU vs the oracle."""
import numpy as np
import pytest

from gpu_util import assert_same_tree, build_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fp", [32, 64])
def test_noop_and_rotation(oracle_mod, rk, fp):
    # test/update.cpp:65-155
    m, x, y, z = oracle_mod.Rng(1).uniform_particles(10000, 1.0, fp=fp)
    o, g = build_pair(oracle_mod, rk, x, y, z, m, fp=fp, box_size=10.0)
    perm0, parts0 = g.perm(0), g.parts()
    g.update_positions()
    o.update_positions()
    assert (g.perm(0) == perm0).all() and (g.perm(1) == np.arange(10000)).all()
    for a, b in zip(g.parts(), parts0):
        assert (a == b).all()
    assert_same_tree(o, g, 5e-7 if fp == 32 else 1e-15)
    px, py, pz, _ = g.parts()
    g.update_positions(py, pz, px)
    o.update_positions(py, pz, px)
    assert_same_tree(o, g, 5e-7 if fp == 32 else 1e-15)
    lp = g.perm(1)
    assert (g.parts()[0] == py[lp]).all()
    assert (g.perm(0) == perm0[lp]).all()


def test_box_rededuction(oracle_mod, rk):
    # test/auto_box_size.cpp:34-72
    for fp in (32, 64):
        g = rk.Octree(fp=fp)
        g.build([0, 1, 2, 3], [-4, -5, -6, -7], [4, 5, 3, 1], np.ones(4), max_leaf_n=1, ncrit=1)
        F = g.F
        assert F(g.box_size) == F(14) + F(0.7)
        x, y, z, _ = g.parts()
        g.update_positions(x * 2, y * 2, z * 2)
        assert F(g.box_size) == F(28) + F(1.4)
        x, y, z, _ = g.parts()
        g.update_positions(x / 4, y / 4, z / 4)
        assert F(g.box_size) == F(7) + F(0.35)
        x, y, z, _ = g.parts()
        ip = g.perm(2)
        assert list(x[ip]) == [0, 0.5, 1, 1.5] and list(y[ip]) == [-2, -2.5, -3, -3.5] and list(z[ip]) == [2, 2.5, 1.5, 0.5]


def test_leapfrog_like_drift(oracle_mod, rk):
    """A few drift steps: GPU and oracle stay identical (codes/perms/topology) after every rebuild."""
    m, x, y, z = oracle_mod.plummer(30000)
    o, g = build_pair(oracle_mod, rk, x, y, z, m)
    rng = np.random.default_rng(0)
    for step in range(3):
        px, py, pz, _ = g.parts()
        d = [rng.normal(scale=0.05, size=px.size).astype(np.float32) for _ in range(3)]
        nx, ny, nz = px + d[0], py + d[1], pz + d[2]
        g.update_positions(nx, ny, nz)
        o.update_positions(nx, ny, nz)
        assert_same_tree(o, g, 5e-7)


def test_update_out_of_box_clears_tree(oracle_mod, rk):
    m, x, y, z = oracle_mod.Rng(2).uniform_particles(1000, 1.0)
    g = rk.Octree()
    g.build(x, y, z, m, box_size=2.0)
    px, py, pz, _ = g.parts()
    px[3] = 50.0
    with pytest.raises(rk.RakauError) as e:
        g.update_positions(px, py, pz)
    assert "outside the allowed bounds" in str(e.value)
    assert g.nparts == 0  # tree.hpp:3760-3764


@pytest.mark.parametrize("fp", [32, 64])
@pytest.mark.parametrize("mac", ["bh", "bh_geom"])
def test_update_masses(oracle_mod, rk, fp, mac):
    # test/update_masses.cpp:50-94, 134-148
    m, x, y, z = oracle_mod.Rng(2).uniform_particles(10000, 1.0, fp=fp)
    o, g = build_pair(oracle_mod, rk, x, y, z, m, fp=fp, mac=mac, box_size=10.0)
    n0 = g.nodes()
    pm = g.parts()[3]
    g.update_masses(pm)
    assert (g.nodes() == n0).all()
    g.update_masses(pm * 2)
    n1 = g.nodes()
    assert (n1["props"][:, 3] == n0["props"][:, 3] * 2).all()
    assert (n1["props"][:, :3] == n0["props"][:, :3]).all()  # COM bit-identical
    for f in ("begin", "end", "n_children", "code", "level"):
        assert (n1[f] == n0[f]).all()
    g.update_masses(pm * 0)
    n2 = g.nodes()
    assert (n2["props"][:, 3] == 0).all()
    for i in range(0, len(n2), 13):
        assert (n2["props"][i, :3] == o.node_centre(n2["code"][i]).astype(g.F)).all()  # == get_node_centre
    bad = pm.copy()
    bad[5] = np.inf
    with pytest.raises(rk.RakauError):
        g.update_masses(bad)
    assert g.nparts == 0
