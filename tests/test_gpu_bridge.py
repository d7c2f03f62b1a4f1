"""The seam-level drop-in as compiled code: oracle/_ref/libref_bridge.so is the UNMODIFIED
reference header built with RAKAU_WITH_CUDA whose cuda_acc_pot_impl / cuda_device_count / cuda_min_size are defined by
integration/rakau_b200_bridge.cpp on top of librakau_b200.so. The reference's own accs_u(..., split = {0, 1}) - its
accelerator branch, tree.hpp:3131-3257 - therefore runs on this library (rk_traverse_external_tree).

Also: oracle/_ref/libref_cuda.so = the same header with the reference's OWN CUDA backend (src/rakau_cuda.cu compiled
with nvcc for sm_100a): it must agree with its CPU path, which makes it a valid second baseline for bench.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = np.stack(a[:3], 1).astype(np.float64)
    b = np.stack(b[:3], 1).astype(np.float64)
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


@pytest.mark.parametrize("fp,mac,Q", [(32, "bh", 0), (32, "bh_geom", 2), (64, "bh", 2)])
def test_reference_split_lands_in_this_library(oracle_mod, rk, fp, mac, Q):
    if not oracle_mod.ref_available("bridge"):
        pytest.skip("oracle/_ref/libref_bridge.so was not built (needs /root/reference at build time)")
    n = 60000
    m, x, y, z = oracle_mod.plummer(n, fp=fp)
    theta, G, eps = (0.75, 2.5, 0.01) if fp == 32 else (0.5, 1.0, 0.0)
    o = oracle_mod.OracleTree(x, y, z, m, fp=fp, mac=mac)
    want, cnt = o.acc_pot(Q, theta, G=G, eps=eps, nthreads=8)
    t = oracle_mod.RefTree(x, y, z, m, fp=fp, mac=mac, variant="bridge")
    assert "rakau_b200_bridge" in t.variant()
    l0 = rk.kernel_launch_count()
    # all the work on the accelerator: CPU share 0
    got = t.acc_pot(Q, theta, G=G, eps=eps, split=[0.0, 1.0])
    assert rk.kernel_launch_count() > l0  # the reference's call reached the CUDA kernels of librakau_b200.so
    e = _rel(got, want) if Q != 1 else None
    if fp == 32:
        assert np.median(e) <= 1e-6 and e.max() <= 1e-4, (np.median(e), e.max())
    else:
        assert e.max() <= 1e-12, e.max()
    if Q == 2:
        pe = np.abs(got[3].astype(np.float64) - want[3]) / np.abs(want[3].astype(np.float64))
        assert pe.max() <= (1e-4 if fp == 32 else 1e-12)
    # a mixed split: the reference evaluates the first critical nodes on its CPU path, the rest through the seam
    mixed = t.acc_pot(Q, theta, G=G, eps=eps, split=[1.0, 2.0])
    e = _rel(mixed, want)
    assert e.max() <= (1e-4 if fp == 32 else 1e-12)
    # the reference's own validation of `split` still applies (tree.hpp:3134-3141)
    with pytest.raises(oracle_mod.OracleError, match="accelerators"):
        t.acc_pot(Q, theta, split=[0.0] + [1.0] * (rk.device_count() + 1))


def test_reference_cuda_backend_builds_and_agrees_with_its_cpu_path(oracle_mod):
    if not oracle_mod.ref_available("cuda"):
        pytest.skip("oracle/_ref/libref_cuda.so was not built")
    n = 60000
    m, x, y, z = oracle_mod.plummer(n)
    t = oracle_mod.RefTree(x, y, z, m, variant="cuda")
    assert "src/rakau_cuda.cu" in t.variant()
    cpu = t.acc_pot(0, 0.75)
    gpu = t.acc_pot(0, 0.75, split=[0.0, 1.0])
    # the reference's GPU kernel applies the MAC per particle, its CPU path per group of <= ncrit particles: the two
    # agree at the level of the Barnes-Hut approximation itself, not bit for bit
    e = _rel(gpu, cpu)
    assert np.isfinite(e).all() and np.median(e) < 5e-3, np.median(e)
