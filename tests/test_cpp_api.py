"""The C++17 drop-in header (include/rakau/tree.hpp): the reference's own test programs, restated in
tests/cpp/, compile on the CPU box and pass on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")
BINS = ["readme_example", "test_basic", "test_update", "test_accuracy", "test_split"]


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


def test_benchmark_leapfrog_cli_validates_options(rk):
    """benchmark/benchmark_leapfrog.cpp: the reference's option checks (benchmark_leapfrog.cpp:146-171)."""
    subprocess.check_call(["make", "-C", BENCH], stdout=subprocess.DEVNULL)
    exe = os.path.join(BENCH, "bin", "benchmark_leapfrog")
    for args, msg in ((["--nparts", "0"], "The number of particles cannot be zero"),
                      (["--timestep", "-1"], "The integration timestep must be finite and positive"),
                      (["--a", "0"], "The Plummer core radius must be finite and positive"),
                      (["--mac_type", "x"], "'x' is not a valid MAC type")):
        r = subprocess.run([exe] + args, capture_output=True, text=True)
        assert r.returncode == 1 and msg in r.stderr, (args, r.stderr)
    for name in ("benchmark_pot", "benchmark_acc_pot"):
        r = subprocess.run([os.path.join(BENCH, "bin", name), "--help"], capture_output=True, text=True)
        assert r.returncode == 0 and "--mac_value" in r.stdout


@pytest.mark.gpu
def test_benchmark_move_runs(rk):
    """benchmark/benchmark_move.cpp (reference: benchmark/benchmark_move.cpp:44-76): evaluate, shift all particles
    through update_particles_u, evaluate again. A uniform shift inside the fixed box leaves the forces unchanged up to
    rounding, and both tree results agree with the direct sums printed beside them."""
    subprocess.check_call(["make", "-C", BENCH], stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(BENCH, "bin", "benchmark_move"), "--nparts", "200000", "--idx", "11", "--bsize",
                        "400", "--mac_value", "0.5"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    vec = [[float(v) for v in ln.split(",")] for ln in r.stdout.strip().splitlines() if ln.count(",") == 2]
    assert len(vec) == 4
    rel = lambda a, b: sum((u - v) ** 2 for u, v in zip(a, b)) ** 0.5 / sum(v * v for v in b) ** 0.5
    assert rel(vec[0], vec[1]) < 2e-2 and rel(vec[2], vec[3]) < 2e-2
    assert rel(vec[3], vec[1]) < 1e-3  # exact sums before / after the shift


@pytest.mark.gpu
@pytest.mark.parametrize("name,nlines", [("benchmark_pot", 1), ("benchmark_acc_pot", 3)])
def test_benchmark_pot_and_acc_pot_run(rk, name, nlines):
    """The printed tree result on particle --idx agrees with the printed direct sum (a smoke check of the programs; the
    parity statements are tests/test_gpu_traverse.py and tests/test_gpu_sizes.py)."""
    subprocess.check_call(["make", "-C", BENCH], stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(BENCH, "bin", name), "--nparts", "200000", "--idx", "5", "--mac_value", "0.5"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if not ln.startswith("Elapsed") and "|" not in ln]
    if name == "benchmark_pot":
        tree_v, exact_v = float(lines[-2]), float(lines[-1])
    else:
        tree_v, exact_v = float(lines[-2]), float(lines[-1])  # potentials follow the two acceleration lines
        ta, ea = ([float(v) for v in ln.split(",")] for ln in lines[-4:-2])
        assert sum((a - b) ** 2 for a, b in zip(ta, ea)) ** 0.5 / sum(b * b for b in ea) ** 0.5 < 2e-2
    assert abs(tree_v - exact_v) / abs(exact_v) < 2e-2


@pytest.mark.gpu
def test_benchmark_leapfrog_host_and_device_loops_agree(rk):
    """The reference's loop with host functors (update_particles_u) and the device-resident integrator perform the same
    fused multiply-adds on the same accelerations: their conserved quantities agree step by step."""
    subprocess.check_call(["make", "-C", BENCH], stdout=subprocess.DEVNULL)
    exe = os.path.join(BENCH, "bin", "benchmark_leapfrog")
    outs = []
    for mode in ([], ["--device"]):
        r = subprocess.run([exe, "--nparts", "100000", "--steps", "4", "--track-integrals", "--timestep", "0.001"] + mode,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        e = [float(ln.split(":")[1]) for ln in r.stdout.splitlines() if ln.startswith("Total energy")]
        assert len(e) == 4 and "Average time per step" in r.stdout
        outs.append(e)
    for a, b in zip(*outs):
        assert abs(a - b) <= 2e-5 * abs(b), outs
    assert abs(outs[1][-1] - outs[1][0]) <= 1e-3 * abs(outs[1][0])


@pytest.mark.gpu
def test_phase_timers_like_the_reference(rk, tmp_path):
    """-DRAKAU_WITH_TIMER: the reference's 'Elapsed time for ...' lines with its phase names
    (include/rakau/detail/simple_timer.hpp:21-47; tree.hpp:934, 1270, 1333, 1440, 1460, 3297)."""
    exe = str(tmp_path / "readme_timer")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-DRAKAU_WITH_TIMER", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(CPP, "readme_example.cpp"), "-o", exe, "-L" + os.path.join(ROOT, "rakau_b200", "lib"),
                           "-lrakau_b200", "-Wl,-rpath," + os.path.join(ROOT, "rakau_b200", "lib")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    for phase in ("overall tree construction", "morton encoding", "indirect code sorting", "permute", "node building",
                  "vector accs/pots computation"):
        assert f"Elapsed time for '{phase}': " in r.stdout, (phase, r.stdout[-1500:])
