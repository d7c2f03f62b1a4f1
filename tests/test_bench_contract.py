"""bench.py contract, CPU side: the reference arm (the reference's own CPU path, oracle/_ref when built, else the
oracle port) prints one JSON line with the agreed keys; our arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line(rk):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--nparts", "200000",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, RK_BENCH_PRINT_MAPS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "librakau_b200" not in r.stderr and "liboracle" in r.stderr  # the arm maps the checker, never the product
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Ginteractions/s" and d["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "Ginteractions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "plummer_200000_fp32_theta0.75_accs"
    assert d["steps"] == 1 and d["steps_requested"] == 1  # the steps actually timed
    # both arms describe the workload with the same keys and values
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    ns = argparse.Namespace(gpus=1, nparts=200000, theta=0.75, max_leaf_n=16, ncrit=128)
    assert d["config"] == bench.workload_config(ns, 200000)


def test_reference_arm_other_ranks_exit_without_work(rk):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_needs_a_gpu(rk):
    import torch
    if torch.cuda.is_available():
        return  # measured by the driver on the GPU box
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--nparts", "1000"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0  # fails loudly: there is no CPU path
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
