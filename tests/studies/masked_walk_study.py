"""CPU design study (oracle tree, not product code): phase 1.5 = ONE masked walk of a run's frontier subtrees that
resolves, per node, the group MAC of every group still interested in it (bit mask), producing per-group source lists;
phase 2' then only streams its list. Asserts the per-group interaction count of the reference and prints the sizes that
decide whether the kernel is worth writing: masked-walk node visits and (node, group) tests per run vs the per-group
frontier walks of the current kernel, list lengths."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
W = 256
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 60
m, x, y, z = oracle.plummer(n)
t = oracle.OracleTree(x, y, z, m, fp=32, mac="bh", max_leaf_n=16, ncrit=128)
nodes = t.nodes(); crit, _ = t.crit()
px, py, pz, pm = t.parts()
P = np.stack([px, py, pz], 1).astype(np.float64)
box = t.box_size; theta2 = 0.75 ** 2
beg = nodes["begin"].astype(np.int64); end = nodes["end"].astype(np.int64); nd = nodes["n_children"].astype(np.int64)
lvl = nodes["level"].astype(np.int64); props = nodes["props"].astype(np.float64)
M = len(nodes)
first = {}
for i in range(M):
    first.setdefault((int(beg[i]), int(end[i])), i)
cn = np.array([first[(int(b), int(e))] for _, b, e in crit], dtype=np.int64)
cb = np.array([b for _, b, e in crit], dtype=np.int64)
cuts = np.flatnonzero(np.diff(cb // W)) + 1
runs = [list(r) for r in np.split(np.arange(len(cn)), cuts)]

def children(i):
    j = i + 1; e = i + 1 + nd[i]
    while j < e:
        yield j; j += nd[j] + 1

rng = np.random.default_rng(5)
S = dict(runs=0, groups=0, p1_visits=0, p2_visits=0, m_visits=0, m_tests=0, list_entries=0, inter_ref=0, inter_new=0, maxpop=0)
for ri in rng.choice(len(runs), NS, replace=False):
    r = runs[ri]; gs = cn[r]
    if len(gs) < 2:
        continue
    sb, se = beg[gs[0]], end[gs[-1]]
    T = [P[beg[g]:end[g]] for g in gs]; TA = P[sb:se]; lo = TA.min(0); hi = TA.max(0); nT = se - sb
    S["runs"] += 1; S["groups"] += len(gs)
    # reference per-group interactions
    for k, g in enumerate(gs):
        st = [0]
        while st:
            i = st.pop()
            if i == g: continue
            if beg[i] <= beg[g] and end[g] <= end[i]:
                st.extend(children(i)); continue
            size = box / 2.0 ** lvl[i]; mac_lh = size * size / theta2
            if (mac_lh < ((T[k] - props[i, :3]) ** 2).sum(1)).all(): S["inter_ref"] += len(T[k])
            elif nd[i] == 0: S["inter_ref"] += len(T[k]) * (end[i] - beg[i])
            else: st.extend(children(i))
    # phase 1 (union box)
    frontier = []; st = [0]; shared = 0
    while st:
        i = st.pop(); S["p1_visits"] += 1
        if beg[i] <= sb and se <= end[i]:
            st.extend(children(i)); continue
        if beg[i] < se and end[i] > sb:
            frontier.append(i); continue
        c = props[i, :3]; size = box / 2.0 ** lvl[i]; mac_lh = size * size / theta2
        gap = np.maximum(0, np.maximum(lo - c, c - hi)); far = np.maximum(np.abs(lo - c), np.abs(hi - c))
        if mac_lh < (gap ** 2).sum() * (1 - 2.0 ** -20): shared += 1
        elif mac_lh >= (far ** 2).sum() * (1 + 2.0 ** -20):
            if nd[i] == 0: shared += end[i] - beg[i]
            else: st.extend(children(i))
        else: frontier.append(i)
    S["inter_new"] += shared * nT
    # current phase 2 visits (per-group frontier walks), for comparison
    for k, g in enumerate(gs):
        st = list(frontier)
        while st:
            i = st.pop(); S["p2_visits"] += 1
            if i == g: continue
            if beg[i] <= beg[g] and end[g] <= end[i]:
                st.extend(children(i)); continue
            size = box / 2.0 ** lvl[i]; mac_lh = size * size / theta2
            if (mac_lh < ((T[k] - props[i, :3]) ** 2).sum(1)).all(): pass
            elif nd[i] != 0: st.extend(children(i))
    # phase 1.5: masked walk
    G = len(gs); full = (1 << G) - 1
    lists = [0] * G
    st = [(i, full) for i in frontier]
    while st:
        i, mask = st.pop(); S["m_visits"] += 1; S["maxpop"] = max(S["maxpop"], bin(mask).count("1"))
        c = props[i, :3]; size = box / 2.0 ** lvl[i]; mac_lh = size * size / theta2
        desc = 0
        for k in range(G):
            if not (mask >> k) & 1: continue
            g = gs[k]
            if i == g: continue
            if beg[i] <= beg[g] and end[g] <= end[i]:
                desc |= 1 << k; continue
            S["m_tests"] += 1
            if (mac_lh < ((T[k] - c) ** 2).sum(1)).all():
                lists[k] += 1; S["inter_new"] += len(T[k]); S["list_entries"] += 1
            else:
                desc |= 1 << k
        if desc:
            if nd[i] == 0:
                for k in range(G):
                    if (desc >> k) & 1:
                        S["inter_new"] += len(T[k]) * (end[i] - beg[i]); S["list_entries"] += 1
            else:
                st.extend((j, desc) for j in children(i))
assert S["inter_new"] == S["inter_ref"], (S["inter_new"], S["inter_ref"])
g, r = S["groups"], S["runs"]
print(f"n={n} runs sampled {r}, groups/run {g/r:.2f}; interactions equal: {S['inter_ref']}")
print(f"per run: phase-1 visits {S['p1_visits']/r:.0f}; current phase-2 visits {S['p2_visits']/r:.0f} ({S['p2_visits']/g:.0f} per group); "
      f"masked walk visits {S['m_visits']/r:.0f} with {S['m_tests']/r:.0f} (node, group) tests ({S['m_tests']/S['m_visits']:.2f} per visit); "
      f"list entries {S['list_entries']/r:.0f} ({S['list_entries']/g:.0f} per group); max groups per run {S['maxpop']}")
