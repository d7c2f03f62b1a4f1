"""GPU check + timing of the experimental bottom-up node properties (RK_PROPS_BOTTOMUP=1, build.cu). Not collected
by pytest: the variant was written without GPU time left in round 1 and has NOT been run yet.

  gpurun -- python tests/studies/props_bottomup_check.py

Runs the tree-build parity test of tests/gpu_util.assert_same_tree against the oracle with the flag on, then times
the properties phase with the flag off and on at 4 M and 32 M particles."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, oracle, rakau_b200 as rk
from gpu_util import build_pair, assert_same_tree
for fp, mac, n in ((32, "bh", 200000), (64, "bh_geom", 100000), (32, "bh_geom", 50000)):
    m, x, y, z = oracle.plummer(n, fp=fp)
    o, g = build_pair(oracle, rk, x, y, z, m, fp=fp, mac=mac)
    assert_same_tree(o, g, 4 * float(np.finfo(o.F).eps))
    print("parity ok", fp, mac, n, flush=True)
for n in (4_000_000, 32_000_000):
    m, x, y, z = rk.plummer(n)
    g = rk.Octree()
    t = []
    for it in range(4):
        t.append(g.build(x, y, z, m).ms_props)
    print("n", n, "ms_props", round(min(t), 3), flush=True)
''' % (ROOT, ROOT)
for flag in ("0", "1"):
    print("RK_PROPS_BOTTOMUP =", flag, flush=True)
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, RK_PROPS_BOTTOMUP=flag))
    if r.returncode:
        sys.exit(r.returncode)
