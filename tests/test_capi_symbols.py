"""The C-ABI library loads and exports every symbol include/rakau_b200.h declares. No compute calls (CPU box)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rakau_b200.h")).read()
    return sorted(set(re.findall(r"RK_API\s+[\w\s\*]+?\b(rk_\w+)\s*\(", src)))


def test_header_declares_the_expected_surface(rk):
    syms = declared_symbols()
    assert len(syms) >= 25
    assert sorted(rk.SYMBOLS) == syms


def test_library_exports_every_symbol(rk):
    L = ctypes.CDLL(rk.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), s


def test_min_size_matches_reference(rk):
    # cuda_min_size(), src/rakau_cuda.cu:26-29
    assert rk.lib().rk_min_size() == 1000


def test_no_cpu_fallback(rk):
    """Without a CUDA device construction must fail loudly (never route to a CPU path)."""
    if rk.device_count() == 0:
        import pytest
        with pytest.raises(RuntimeError) as e:
            rk.Octree()
        assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    for dp, _, fs in os.walk(os.path.join(ROOT, "rakau_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, f
    for dp, _, fs in os.walk(os.path.join(ROOT, "include")):
        for f in fs:
            txt = open(os.path.join(dp, f)).read()
            assert "liboracle" not in txt and "rakau_oracle" not in txt, f
