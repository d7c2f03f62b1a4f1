"""The oracle reproduces the UNMODIFIED reference's tree at BASELINE.json's full size (4 M particles): its codes,
permutation, node topology, node properties and critical nodes hash to the digests that
tests/golden/make_golden_sizes.py took from oracle/_ref/libref_scalar.so. CPU only (the `-m gpu` counterpart,
tests/test_gpu_sizes.py, checks the CUDA build and traversal against the same fixture)."""
import pytest

from test_gpu_sizes import _digest, _load, sha


@pytest.mark.parametrize("name", ["config1_fp32_accs", "config3_fp64_accs"])
def test_oracle_tree_digests_at_4m(oracle_mod, name):
    fix, _ = _load()
    c = fix[name]
    m, x, y, z = oracle_mod.plummer(c["nparts"], fp=c["fp"])
    o = oracle_mod.OracleTree(x, y, z, m, fp=c["fp"], max_leaf_n=c["max_leaf_n"], ncrit=c["ncrit"])
    d, nodes = _digest(o, o.crit()[0])
    assert d == {k: v for k, v in c["tree"].items() if k != "node_props"}
    assert sha(nodes["props"]) == c["tree"]["node_props"]
