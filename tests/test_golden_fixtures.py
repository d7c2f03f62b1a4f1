"""Golden fixtures generated from the UNMODIFIED reference header (tests/golden/make_golden.py; the reference does
not exist on the GPU box, the fixtures travel). CPU: the oracle reproduces them bit for bit. GPU: the CUDA path
reproduces codes / permutations / particle order / topology bit for bit, node properties within the bound of
gpu_util.assert_same_tree, and accelerations + potentials within the north_star tolerance (fp32: median relative
error <= 1e-6, max <= 1e-4; fp64: max <= 1e-12) when it traverses the reference's own node array."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "plummer*.npz")))
TOPO = ("begin", "end", "n_children", "code", "level")


def _load(path):
    d = dict(np.load(path))
    name = os.path.basename(path)[:-4]
    fp = 32 if "_fp32_" in name else 64
    mac = "bh_geom" if name.endswith("bh_geom") else "bh"
    return d, fp, mac


def _nodes(d, dtype):
    out = np.zeros(d["node_begin"].size, dtype=dtype)
    for f in dtype.names:
        out[f] = d["node_" + f]
    return out


def test_fixtures_exist():
    assert len(FIXTURES) >= 4


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_oracle_reproduces_the_reference_fixture(oracle_mod, path):
    d, fp, mac = _load(path)
    o = oracle_mod.OracleTree(d["x"], d["y"], d["z"], d["m"], fp=fp, mac=mac, max_leaf_n=int(d["max_leaf_n"]),
                              ncrit=int(d["ncrit"]))
    assert o.box_size == float(d["box_size"])
    assert (o.codes() == d["codes"]).all()
    for w, key in enumerate(("perm", "last_perm", "inv_perm")):
        assert (o.perm(w) == d[key]).all(), key
    for a, key in zip(o.parts(), ("px", "py", "pz", "pm")):
        assert (a == d[key]).all(), key
    on = o.nodes()
    for f in on.dtype.names:
        assert (on[f] == d["node_" + f]).all(), f  # every field, masses and centres of mass included: bit for bit
    out, _ = o.acc_pot(2, float(d["theta"]), G=float(d["G"]), eps=float(d["eps"]))
    for a, key in zip(out, ("ax", "ay", "az", "pot")):
        assert (a == d[key]).all(), key
    n = d["x"].size
    for row, i in zip(d["exact"], (0, n // 2, n - 1)):
        assert (np.asarray(o.exact(i, G=float(d["G"]), eps=float(d["eps"]))) == row).all()


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_cuda_path_reproduces_the_reference_fixture(rk, path):
    d, fp, mac = _load(path)
    F = np.float32 if fp == 32 else np.float64
    theta, G, eps, ncrit = float(d["theta"]), float(d["G"]), float(d["eps"]), int(d["ncrit"])
    g = rk.Octree(fp=fp, mac=mac)
    g.build(d["x"], d["y"], d["z"], d["m"], max_leaf_n=int(d["max_leaf_n"]), ncrit=ncrit)
    # ---- build: integers bit-exact ----
    assert F(g.box_size) == F(d["box_size"])
    assert (g.codes() == d["codes"]).all()
    for w, key in enumerate(("perm", "last_perm", "inv_perm")):
        assert (g.perm(w) == d[key]).all(), key
    for a, key in zip(g.parts(), ("px", "py", "pz", "pm")):
        assert (a == d[key]).all(), key
    gn = g.nodes()
    assert len(gn) == d["node_begin"].size
    for f in TOPO:
        assert (gn[f] == d["node_" + f]).all(), f
    assert (gn["dim"] == d["node_dim"]).all()
    # node properties: bounded by the reference's own sequential-summation error (see gpu_util.assert_same_tree)
    cnt = (d["node_end"] - d["node_begin"]).astype(np.float64)
    bound = np.maximum(4 * np.finfo(F).eps, (cnt + 8.0) * float(np.finfo(F).eps))
    rp, gp = d["node_props"].astype(np.float64), gn["props"].astype(np.float64)
    mass_err = np.abs(gp[:, 3] - rp[:, 3]) / np.maximum(np.abs(rp[:, 3]), 1e-300)
    assert (mass_err <= bound).all()
    size = float(d["box_size"]) / (2.0 ** d["node_level"].astype(np.float64))
    com_err = np.abs(gp[:, :3] - rp[:, :3]).max(axis=1) / (np.abs(rp[:, :3]).max(axis=1) + size)
    assert (com_err <= bound).all()
    # ---- traversal of the reference's own node array (the cuda_acc_pot_impl seam): values within tolerance ----
    mac_value = F(1) / (F(theta) * F(theta)) if mac == "bh" else F(1) / F(theta)
    out, info = rk.traverse_external_tree(_nodes(d, rk.NODE_DTYPE[fp]), [d["px"], d["py"], d["pz"], d["pm"]], d["codes"], 2,
                                          mac_value, G=G, eps2=float(F(eps) * F(eps)), mac=mac, fp=fp, ncrit=ncrit)
    ref_a = np.stack([d["ax"], d["ay"], d["az"]], 1).astype(np.float64)
    got_a = np.stack(out[:3], 1).astype(np.float64)
    ea = np.linalg.norm(got_a - ref_a, axis=1) / np.linalg.norm(ref_a, axis=1)
    rp_, gp_ = d["pot"].astype(np.float64), out[3].astype(np.float64)
    nz = rp_ != 0
    assert (gp_[~nz] == 0).all()  # massless targets: the potential is exactly zero
    ep = np.abs(gp_[nz] - rp_[nz]) / np.abs(rp_[nz])
    for e in (ea, ep):
        assert np.isfinite(e).all()
        if fp == 32:
            assert np.median(e) <= 1e-6 and e.max() <= 1e-4, (np.median(e), e.max())
        else:
            assert e.max() <= 1e-12, e.max()
    # ---- and the tree built on the device gives the same answer up to MAC decisions within an ulp ----
    go = g.acc_pot(2, theta, G=G, eps=eps)
    eg = np.linalg.norm(np.stack(go[:3], 1).astype(np.float64) - ref_a, axis=1) / np.linalg.norm(ref_a, axis=1)
    assert np.median(eg) <= (1e-6 if fp == 32 else 1e-12)
