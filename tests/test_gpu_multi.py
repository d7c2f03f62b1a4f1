"""Multi-GPU path on hardware (needs >= 2 visible GPUs; skipped otherwise): bench.py's N > 1 arm is launched as two
ranks over NCCL and must report that the sharded tree (fingerprints of every device array) and the gathered
accelerations equal a single-GPU build + evaluation of the same particles bit for bit - also when the particles move
between steps, i.e. when the cost-weighted cuts have to survive a rebuild."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _two_ranks(extra):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2",
           "--warmup", "3", "--nparts", "3000000"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_two_rank_result_equals_single_gpu():
    d = _two_ranks([])
    assert d["n_gpus"] == 2 and d["parity_checked"] is True, d.get("parity")
    assert d["parity"]["tree_arrays_differing"] == [] and d["parity"]["accelerations_bit_equal"] is True
    assert d["parity"]["max_abs_diff"] == 0.0


def test_two_rank_cuts_survive_moving_particles():
    d = _two_ranks(["--perturb"])
    assert d["parity_checked"] is True, d.get("parity")
    assert d["tree"]["shard_cost_imbalance"] is not None and d["tree"]["shard_cost_imbalance"] < 1.2
