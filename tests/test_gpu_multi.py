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


def _two_ranks(extra, env=None):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2",
           "--warmup", "3", "--nparts", "3000000"] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_two_rank_result_equals_single_gpu():
    d = _two_ranks([])
    assert d["n_gpus"] == 2 and d["parity_checked"] is True, d.get("parity")
    assert d["parity"]["tree_arrays_differing"] == [] and d["parity"]["accelerations_bit_equal"] is True
    assert d["parity"]["max_abs_diff"] == 0.0


@pytest.mark.parametrize("env,mode", [({"RK_MULTICAST": "1", "RK_MULTICAST_CODES": "1"}, "in-kernel multicast stores"),
                                      ({"RK_MIRROR_EXCHANGE": "0"}, "copy-engine pushes of 4 launches")])
def test_two_rank_other_exchange_schemes(env, mode):
    """The output exchange has three implementations (in-kernel stores to each peer: the default between two ranks;
    NVSwitch multicast stores: the default for more; copy-engine pushes of four chunked launches: the fallback without
    peer-mapped stores); all of them - and the multicast gather of the codes - must reproduce the single-GPU result."""
    d = _two_ranks(["--perturb"], env)
    if d["tree"]["output_exchange"] != mode:
        pytest.skip("this platform has no NVSwitch multicast: " + str(d["tree"]["output_exchange"]))
    assert d["parity_checked"] is True, d.get("parity")
    assert d["parity"]["accelerations_bit_equal"] is True and d["parity"]["tree_arrays_differing"] == []


def test_two_rank_cuts_survive_moving_particles():
    d = _two_ranks(["--perturb"])
    assert d["parity_checked"] is True, d.get("parity")
    assert d["tree"]["shard_cost_imbalance"] is not None and d["tree"]["shard_cost_imbalance"] < 1.2


def test_split_across_devices_in_one_process(rk):
    """The reference's `split` kwarg with one share per device (tree.hpp:3147-3198), one process driving them: equal to
    the one-device evaluation bit for bit, ordered and unordered outputs, counters included."""
    import numpy as np
    ndev = rk.device_count()
    if ndev < 2:
        pytest.skip("needs two GPUs")
    m, x, y, z = rk.plummer(400000)
    g = rk.Octree()
    g.build(x, y, z, m)
    for ordered in (False, True):
        want = g.acc_pot(2, 0.75, G=2.5, eps=0.01, ordered=ordered)
        wi = g.eval_info.asdict()
        for split in ([0.0] + [1.0] * ndev, [0.5, 2.0] + [1.0] * (ndev - 1)):
            got = g.acc_pot(2, 0.75, G=2.5, eps=0.01, ordered=ordered, split=split)
            gi = g.eval_info.asdict()
            for a, b in zip(got, want):
                assert (a == b).all(), (ordered, split)
            for k in ("mac_tests", "accepted", "p2p_pairs", "self_pairs", "interactions", "n_groups"):
                assert gi[k] == wi[k], (k, gi[k], wi[k])
    # moved particles: the mirrors on the other devices follow the rebuild
    px, py, pz, pm = g.parts()
    g.update_positions(px + np.float32(1e-3), py, pz)
    want = g.acc_pot(0, 0.75, ordered=True)
    got = g.acc_pot(0, 0.75, ordered=True, split=[0.0] + [1.0] * ndev)
    for a, b in zip(got, want):
        assert (a == b).all()
