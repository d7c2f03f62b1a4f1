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


BENCH = os.path.join(ROOT, "benchmark")


def test_benchmark_cli_compiles_and_validates_options(rk):
    """benchmark/benchmark_acc.cpp: the reference's command line (benchmark/common.hpp:143-229). Option checks run
    before any device work, so they are testable without a GPU."""
    subprocess.check_call(["make", "-C", BENCH], stdout=subprocess.DEVNULL)
    exe = os.path.join(BENCH, "bin", "benchmark_acc")
    r = subprocess.run([exe, "--nparts", "0"], capture_output=True, text=True)
    assert r.returncode == 1 and "The number of particles cannot be zero" in r.stderr
    r = subprocess.run([exe, "--nparts", "10", "--idx", "10"], capture_output=True, text=True)
    assert r.returncode == 1 and "less-than the total number of particles (10)" in r.stderr
    r = subprocess.run([exe, "--fp_type", "half"], capture_output=True, text=True)
    assert r.returncode == 1 and "Only the 'float' and 'double' floating-point types are supported" in r.stderr
    r = subprocess.run([exe, "--mac_type", "foo"], capture_output=True, text=True)
    assert r.returncode == 1 and "'foo' is not a valid MAC type" in r.stderr
    r = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--max_leaf_n" in r.stdout and "--split" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["--ordered"], ["--fp_type", "double", "--mac_type", "bh_geom", "--mac_value", "0.5"],
                                   ["--parinit", "--ncrit", "64", "--max_leaf_n", "8", "--bsize", "100"]])
def test_benchmark_acc_runs(rk, extra):
    subprocess.check_call(["make", "-C", BENCH], stdout=subprocess.DEVNULL)
    exe = os.path.join(BENCH, "bin", "benchmark_acc")
    r = subprocess.run([exe, "--nparts", "200000", "--idx", "77"] + extra, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    vec = [ln for ln in r.stdout.strip().splitlines() if ln.count(",") == 2][-2:]
    tree_acc, exact = ([float(v) for v in ln.split(",")] for ln in vec)
    num = sum((a - b) ** 2 for a, b in zip(tree_acc, exact)) ** 0.5
    den = sum(b ** 2 for b in exact) ** 0.5
    assert num / den < 5e-2, (tree_acc, exact)
