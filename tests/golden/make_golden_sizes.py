"""Compact fixtures at BASELINE.json's full sizes (4 M particles), generated from the UNMODIFIED reference header
(oracle/_ref/libref_scalar.so, see make_golden.py) - hashes and counts instead of 200 MB arrays.

Run where /root/reference exists:   make -C oracle ref && python tests/golden/make_golden_sizes.py
Writes tests/golden/baseline_sizes.json + baseline_sizes_samples.npz. Per configuration (BASELINE.json configs
1-3): sha256 of the reference's codes, perm, node begin/end/n_children/code/level, box size; the interaction
counters of the evaluation (the reference has none: they come from the oracle, after the oracle's own arrays were
checked equal to the reference's here); the reference's accelerations/potentials on 4096 sampled particles."""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

N = 4_000_000
#         name    fp  Q  theta  G    eps
CONFIGS = [("config1_fp32_accs", 32, 0, 0.75, 1.0, 0.0), ("config2_fp32_accs_pots", 32, 2, 0.75, 2.5, 0.01),
           ("config3_fp64_accs", 64, 0, 0.5, 1.0, 0.0)]
NODE_FIELDS = ("begin", "end", "n_children", "code", "level")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def tree_record(t):
    nodes = t.nodes()
    rec = {"box_size_hex": float(t.box_size).hex(), "n_nodes": int(len(nodes)), "codes": sha(t.codes()),
           "perm": sha(t.perm(0))}
    for f in NODE_FIELDS:
        rec["node_" + f] = sha(nodes[f])
    return rec, nodes


def main():
    assert oracle.ref_available("scalar"), "build oracle/_ref first (make -C oracle ref)"
    out, samples, trees = {}, {}, {}
    rng = np.random.default_rng(20261017)
    sample_idx = np.sort(rng.choice(N, 4096, replace=False)).astype(np.int64)  # positions in MORTON order
    samples["sample_idx"] = sample_idx
    for name, fp, Q, theta, G, eps in CONFIGS:
        t0 = time.time()
        m, x, y, z = oracle.plummer(N, fp=fp)
        if fp not in trees:
            ref = oracle.RefTree(x, y, z, m, fp=fp, variant="scalar")
            rec, rnodes = tree_record(ref)
            orc = oracle.OracleTree(x, y, z, m, fp=fp)
            orec, onodes = tree_record(orc)
            assert orec == rec, "oracle tree differs from the reference's at 4M"
            assert (onodes["props"] == rnodes["props"]).all() and (onodes["dim"] == rnodes["dim"]).all()
            crit, _ = orc.crit()
            rec["n_crit"] = int(len(crit))
            rec["crit"] = sha(crit)
            rec["node_props"] = sha(rnodes["props"])
            trees[fp] = (ref, orc, rec)
        ref, orc, rec = trees[fp]
        racc = ref.acc_pot(Q, theta, G=G, eps=eps)
        oacc, cnt = orc.acc_pot(Q, theta, G=G, eps=eps, nthreads=os.cpu_count() or 1)
        for a, b in zip(racc, oacc):
            assert (a == b).all(), "oracle results differ from the reference's at 4M"
        out[name] = dict(nparts=N, fp=fp, mac="bh", max_leaf_n=16, ncrit=128, Q=Q, theta=theta, G=G, eps=eps, tree=rec,
                         counters={k: int(v) for k, v in cnt.items()}, results=[sha(a) for a in racc])
        for j, a in enumerate(racc):
            samples[f"{name}_out{j}"] = a[sample_idx]
        print(name, "done in %.1f s" % (time.time() - t0), cnt, flush=True)
    with open(os.path.join(HERE, "baseline_sizes.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "baseline_sizes_samples.npz"), **samples)


if __name__ == "__main__":
    main()
