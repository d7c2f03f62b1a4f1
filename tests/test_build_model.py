"""The non-recursive build formulation implemented by the CUDA kernels (tests/build_model.py is its numpy
statement) must reproduce the oracle's recursive build exactly. CPU only."""
import numpy as np
import pytest

from build_model import build


def _check(oracle_mod, x, y, z, m, mln, nc, box=0.0):
    t = oracle_mod.OracleTree(x, y, z, m, max_leaf_n=mln, ncrit=nc, box_size=box)
    nd = t.nodes()
    _, ci = t.crit()
    b = build(t.codes(), mln, nc)
    assert len(nd) == len(b["begin"])
    for f in ("begin", "end", "n_children", "code", "level"):
        assert (nd[f] == b[f]).all(), f
    assert (np.nonzero(b["iscrit"])[0] == ci).all()


@pytest.mark.parametrize("N", [1, 2, 5, 17, 100, 1000, 3000])
@pytest.mark.parametrize("mln,nc", [(1, 1), (2, 16), (16, 128), (8, 4), (16, 16), (3, 1000)])
def test_plummer(oracle_mod, N, mln, nc):
    m, x, y, z = oracle_mod.plummer(N)
    _check(oracle_mod, x, y, z, m, mln, nc)


@pytest.mark.parametrize("mln,nc", [(1, 1), (16, 128), (4, 2), (8, 600)])
def test_duplicates_reach_max_depth(oracle_mod, mln, nc):
    rng = np.random.default_rng(1)
    N = 3000
    x, y, z = (rng.normal(size=N).astype(np.float32) for _ in range(3))
    x[:500], y[:500], z[:500] = x[0], y[0], z[0]
    x[600:700] = x[600] + 1e-7
    _check(oracle_mod, x, y, z, np.ones(N), mln, nc, box=20.0)
