"""Generates the golden fixtures of this directory from the UNMODIFIED reference header
(/root/reference/include/rakau/tree.hpp compiled by oracle/Makefile into oracle/_ref/libref_scalar.so: the
reference's gcc7_debug_nosimd configuration, -ffp-contract=off, stable sort).

Run where /root/reference exists:   make -C oracle ref && python tests/golden/make_golden.py
The fixtures travel with the repository, so the parity tests do not need the reference at run time:
tests/test_golden_fixtures.py checks the oracle (CPU) and the CUDA path (GPU) against them.

Each plummer<N>_fp<bits>_<mac>.npz holds: the inputs (x, y, z, m in original order, max_leaf_n, ncrit, theta, G,
eps), and the reference's outputs: box_size, codes, perm, last_perm, inv_perm, particles in tree order, the node
array (one array per field), accelerations + potentials (accs_pots_u), and exact_* on three particles."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

CASES = [(3000, 32, "bh", 16, 128, 0.75, 2.5, 0.01), (3000, 32, "bh_geom", 8, 64, 0.75, 1.0, 0.0),
         (2000, 64, "bh", 16, 128, 0.5, 1.0, 0.0), (2000, 64, "bh_geom", 4, 32, 0.5, 6.674e-11, 0.05)]


def main():
    assert oracle.ref_available("scalar"), "build oracle/_ref first (make -C oracle ref)"
    for n, fp, mac, mln, ncrit, theta, G, eps in CASES:
        m, x, y, z = oracle.plummer(n, fp=fp)
        m = m.copy()
        m[::97] = 0  # a few massless particles
        if eps > 0:  # one coincident pair => equal Morton codes (stable order matters); unsoftened it would be NaN
            x[5], y[5], z[5] = x[6], y[6], z[6]
        t = oracle.RefTree(x, y, z, m, max_leaf_n=mln, ncrit=ncrit, mac=mac, fp=fp, variant="scalar")
        nodes = t.nodes()
        out = t.acc_pot(2, theta, G=G, eps=eps)
        px, py, pz, pm = t.parts()
        d = dict(x=x, y=y, z=z, m=m, max_leaf_n=mln, ncrit=ncrit, theta=theta, G=G, eps=eps, box_size=t.box_size,
                 codes=t.codes(), perm=t.perm(0), last_perm=t.perm(1), inv_perm=t.perm(2), px=px, py=py, pz=pz, pm=pm,
                 ax=out[0], ay=out[1], az=out[2], pot=out[3],
                 exact=np.stack([t.exact(i, G=G, eps=eps) for i in (0, n // 2, n - 1)]))
        for f in nodes.dtype.names:
            d["node_" + f] = nodes[f]
        path = os.path.join(HERE, f"plummer{n}_fp{fp}_{mac}.npz")
        np.savez_compressed(path, **d)
        print(path, os.path.getsize(path), "bytes;", len(nodes), "nodes; variant:", t.variant())


if __name__ == "__main__":
    main()
