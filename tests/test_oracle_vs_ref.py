"""Pins the oracle against the reference ITSELF: oracle/_ref/libref_scalar.so is the unmodified
/root/reference/include/rakau/tree.hpp compiled against dependency stand-ins (oracle/ref_shim) with
-DRAKAU_DISABLE_SIMD (the reference's own CI configuration gcc7_debug_nosimd) and a stable sort.
The oracle must agree with it BIT FOR BIT. CPU only; skipped when _ref was not built."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def ref(oracle_mod):
    if not oracle_mod.ref_available("scalar"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return oracle_mod


@pytest.mark.parametrize("fp", [32, 64])
@pytest.mark.parametrize("mac", ["bh", "bh_geom"])
def test_oracle_is_bit_exact_with_reference(ref, fp, mac):
    m, x, y, z = ref.plummer(20000, fp=fp)
    for kw in (dict(), dict(max_leaf_n=4, ncrit=32), dict(box_size=2000.0, max_leaf_n=1, ncrit=1)):
        o = ref.OracleTree(x, y, z, m, fp=fp, mac=mac, **kw)
        r = ref.RefTree(x, y, z, m, fp=fp, mac=mac, variant="scalar", **kw)
        assert "unmodified" not in r.variant() and "scalar" in r.variant()
        assert o.box_size == r.box_size
        assert (o.codes() == r.codes()).all()
        for w in range(3):
            assert (o.perm(w) == r.perm(w)).all()
        for a, b in zip(o.parts(), r.parts()):
            assert (a == b).all()
        on, rn = o.nodes(), r.nodes()
        assert len(on) == len(rn)
        for f in ("begin", "end", "n_children", "code", "level", "props", "dim", "delta"):
            assert (on[f] == rn[f]).all(), f
    o = ref.OracleTree(x, y, z, m, fp=fp, mac=mac)
    r = ref.RefTree(x, y, z, m, fp=fp, mac=mac, variant="scalar")
    for Q in (0, 1, 2):
        for theta, G, eps in ((0.75, 1.0, 0.0), (0.4, 2.5, 0.01)):
            oo, _ = o.acc_pot(Q, theta, G=G, eps=eps)
            ro = r.acc_pot(Q, theta, G=G, eps=eps)
            for a, b in zip(oo, ro):
                assert (a == b).all(), (Q, theta)
    for i in (0, 7, 19999):
        assert (o.exact(i, G=2.0, eps=0.1) == r.exact(i, G=2.0, eps=0.1)).all()


@pytest.mark.parametrize("fp", [32, 64])
def test_updates_bit_exact_with_reference(ref, fp):
    m, x, y, z = ref.Rng(9).uniform_particles(5000, 1.0, fp=fp)
    o = ref.OracleTree(x, y, z, m, fp=fp, box_size=10.0)
    r = ref.RefTree(x, y, z, m, fp=fp, box_size=10.0, variant="scalar")
    px, py, pz, pm = o.parts()
    o.update_positions(py * 0.5, pz + 1, px)
    r.update_positions(py * 0.5, pz + 1, px)
    for w in range(3):
        assert (o.perm(w) == r.perm(w)).all()
    assert (o.nodes() == r.nodes()).all()
    pm = o.parts()[3]
    o.update_masses(pm * 3)
    r.update_masses(pm * 3)
    assert (o.nodes() == r.nodes()).all()


def test_reference_error_messages(ref):
    c = np.array([-10, 1, 2, 10.0])
    with pytest.raises(ref.OracleError) as e:
        ref.RefTree(c, c, c, np.ones(4), box_size=3.0, max_leaf_n=4, ncrit=5)
    assert "produced the floating-point value" in str(e.value)
    t = ref.RefTree(c, c, c, np.ones(4))
    assert t.box_size == 21.0
    with pytest.raises(ref.OracleError) as e:
        t.acc_pot(0, 0.0)
    assert "The MAC value must be finite and positive" in str(e.value)


def test_simd_reference_close_to_oracle(ref):
    """The reference's SIMD + rsqrt path differs from its scalar path only by rounding."""
    v = ref.best_ref_variant()
    if v is None:
        pytest.skip("no SIMD variant built / supported by this CPU")
    m, x, y, z = ref.plummer(30000)
    o = ref.OracleTree(x, y, z, m)
    r = ref.RefTree(x, y, z, m, variant=v)
    on, rn = o.nodes(), r.nodes()
    for f in ("begin", "end", "n_children", "code", "level"):
        assert (on[f] == rn[f]).all()
    oo, _ = o.acc_pot(0, 0.75)
    ro = r.acc_pot(0, 0.75)
    same = o.codes()[1:] != o.codes()[:-1]  # ignore tied particles (unstable sort in this variant)
    keep = np.concatenate([[True], same]) & np.concatenate([same, [True]])
    a, b = np.stack(oo, 1)[keep].astype(np.float64), np.stack(ro, 1)[keep].astype(np.float64)
    rel = np.linalg.norm(a - b, axis=1) / np.linalg.norm(a, axis=1)
    assert np.median(rel) < 2e-6 and rel.max() < 1e-3
