import numpy as np


def build_pair(oracle_mod, rk, x, y, z, m, fp=32, mac="bh", **kw):
    o = oracle_mod.OracleTree(x, y, z, m, fp=fp, mac=mac, **kw)
    o._ncrit = kw.get("ncrit", 128)
    g = rk.Octree(fp=fp, mac=mac)
    g.build(x, y, z, m, **kw)
    return o, g


def assert_same_tree(o, g, props_tol):
    """T2 of SURVEY §8c: codes, perms, topology and critical nodes bit-exact; mass/COM within tolerance."""
    F = o.F
    assert F(g.box_size) == F(o.box_size)
    assert g.nparts == o.nparts
    assert (g.codes() == o.codes()).all()
    for w in range(3):
        assert (g.perm(w) == o.perm(w)).all(), w
    for a, b in zip(g.parts(), o.parts()):
        assert (a == b).all()
    gn, on = g.nodes(), o.nodes()
    assert len(gn) == len(on)
    for f in ("begin", "end", "n_children", "code", "level"):
        assert (gn[f] == on[f]).all(), f
    assert (gn["dim"] == on["dim"]).all()  # dim2 / dim: pure function of level and box
    # Node properties. The reference sums each node sequentially in F (tree.hpp:1162-1168), the GPU reduces in
    # fp64 with a fixed tree and rounds once, so the difference is bounded by the REFERENCE's own summation
    # error: n * eps_F relative for the mass, n * eps_F * (|com| + node size) for each COM component.
    # props_tol caps the bound for small nodes (a few ulps).
    cnt = (on["end"] - on["begin"]).astype(np.float64)
    eps = float(np.finfo(F).eps)
    bound = np.maximum(props_tol, (cnt + 8.0) * eps)
    mass_err = np.abs(gn["props"][:, 3].astype(np.float64) - on["props"][:, 3]) / np.maximum(np.abs(on["props"][:, 3]), 1e-300)
    assert (mass_err <= bound).all(), (mass_err / bound).max()
    size = o.box_size / (2.0 ** on["level"].astype(np.float64))
    scale = np.abs(on["props"][:, :3].astype(np.float64)).max(axis=1) + size
    com_err = np.abs(gn["props"][:, :3].astype(np.float64) - on["props"][:, :3]).max(axis=1) / scale
    assert (com_err <= bound).all(), (com_err / bound).max()
    gc, (oc, _) = g.crit(), o.crit()
    assert gc.shape == oc.shape and (gc == oc).all()


def rel_err_vec(ga, oa):
    ga = np.stack(ga, 1).astype(np.float64)
    oa = np.stack(oa, 1).astype(np.float64)
    return np.linalg.norm(ga - oa, axis=1) / np.linalg.norm(oa, axis=1)
