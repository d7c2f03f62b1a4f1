"""The C++17 drop-in header (include/rakau/tree.hpp): the reference's own test programs, restated in
tests/cpp/, compile on the CPU box and pass on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
BINS = ["readme_example", "test_basic", "test_update", "test_accuracy"]


def build():
    subprocess.check_call(["make", "-C", CPP, "-j", "4"], stdout=subprocess.DEVNULL)


def test_header_and_tests_compile(rk):
    """Needs the C-ABI library to link against, not a GPU."""
    build()
    for b in BINS:
        assert os.path.exists(os.path.join(CPP, "bin", b))


@pytest.mark.gpu
@pytest.mark.parametrize("name", BINS)
def test_reference_test_programs_pass(rk, name):
    build()
    r = subprocess.run([os.path.join(CPP, "bin", name)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
