"""CPU design study (oracle tree, not product code): two-phase traversal per run of sibling critical nodes.

Phase 1 (once per run S of consecutive critical nodes with one parent, <= TS_MAX targets): walk from the root with the
bounding box of S. A node accepted for the whole box is a source for ALL targets of S; a node every group of S rejects
(ancestor of S, or the MAC fails for every target / for one target of every group) is descended / opened for all of S.
Everything else is a FRONTIER node. Phase 2 (per group): the reference's group walk, started from the frontier.
Decisions per group are exactly the reference's (a group never sees a descendant of a node it would have accepted).
Prints node visits and the share of interactions evaluated against the whole run."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
TS_MAX = int(sys.argv[2]) if len(sys.argv) > 2 else 256
NS = int(sys.argv[3]) if len(sys.argv) > 3 else 150
MODE = sys.argv[4] if len(sys.argv) > 4 else 'parent'  # 'parent': runs share a parent; 'morton': any consecutive critical nodes
GMAX = int(sys.argv[5]) if len(sys.argv) > 5 else 8
m, x, y, z = oracle.plummer(n)
t = oracle.OracleTree(x, y, z, m, fp=32, mac="bh", max_leaf_n=16, ncrit=128)
nodes = t.nodes(); crit, _ = t.crit()
px, py, pz, pm = t.parts()
P = np.stack([px, py, pz], 1).astype(np.float64)
box = t.box_size; theta2 = 0.75 ** 2
beg = nodes["begin"].astype(np.int64); end = nodes["end"].astype(np.int64); nd = nodes["n_children"].astype(np.int64)
lvl = nodes["level"].astype(np.int64); props = nodes["props"].astype(np.float64)
M = len(nodes)
parent = np.full(M, -1, dtype=np.int64)
stack = []
for i in range(M):
    while stack and i > stack[-1] + nd[stack[-1]]:
        stack.pop()
    if stack:
        parent[i] = stack[-1]
    stack.append(i)
first = {}
for i in range(M):
    first.setdefault((int(beg[i]), int(end[i])), i)
cn = np.array([first[(int(b), int(e))] for _, b, e in crit], dtype=np.int64)
par = parent[cn]
# runs: consecutive critical nodes, same parent, <= TS_MAX targets, <= 8 groups
runs = []
cur = [0]
cb = np.array([b for _, b, e in crit], dtype=np.int64)
for j in range(1, len(cn) if MODE != 'window' else 0):
    tsz = sum(end[cn[k]] - beg[cn[k]] for k in cur)
    if (MODE == 'morton' or par[j] == par[cur[0]]) and len(cur) < GMAX and tsz + (end[cn[j]] - beg[cn[j]]) <= TS_MAX:
        cur.append(j)
    else:
        runs.append(cur); cur = [j]
runs.append(cur)
if MODE == 'window':  # groups whose first particle lies in the same TS_MAX-aligned window of particles
    w = cb // TS_MAX
    cuts = np.flatnonzero(np.diff(w)) + 1
    runs = [list(r) for r in np.split(np.arange(len(cn)), cuts)]
rl = np.array([len(r) for r in runs]); rt = np.array([sum(end[cn[k]] - beg[cn[k]] for k in r) for r in runs])
print(f"n={n} crit={len(cn)} runs={len(runs)} groups/run mean {rl.mean():.2f} targets/run mean {rt.mean():.1f} max {rt.max()}")

def children(i):
    j = i + 1; e = i + 1 + nd[i]
    while j < e:
        yield j; j += nd[j] + 1

rng = np.random.default_rng(3)
sel = rng.choice(len(runs), NS, replace=False)
for variant in ("box", "gbox", "exact"):
    S = dict(base_visits=0, p1_visits=0, p2_visits=0, frontier=0, inter=0, inter_shared=0, groups=0, runs=0, fmax=0,
             shared_src=0, p2_src=0)
    for ri in sel:
        r = runs[ri]; gs = cn[r]
        sb, se = beg[gs[0]], end[gs[-1]]
        T = [P[beg[g]:end[g]] for g in gs]; TA = P[sb:se]; lo = TA.min(0); hi = TA.max(0); nT = se - sb
        S["runs"] += 1; S["groups"] += len(gs)
        # baseline: separate walks
        for k, g in enumerate(gs):
            st = [0]
            while st:
                i = st.pop(); S["base_visits"] += 1
                if i == g: continue
                if beg[i] <= beg[g] and end[g] <= end[i]:
                    st.extend(children(i)); continue
                size = box / 2.0 ** lvl[i]; mac_lh = size * size / theta2
                d2 = ((T[k] - props[i, :3]) ** 2).sum(1)
                if (mac_lh < d2).all():
                    S["inter"] += len(T[k])
                elif nd[i] == 0:
                    S["inter"] += len(T[k]) * (end[i] - beg[i])
                else:
                    st.extend(children(i))
            S["inter"] += len(T[k]) * (len(T[k]) - 1)
        # phase 1
        frontier = []
        st = [0]
        while st:
            i = st.pop(); S["p1_visits"] += 1
            if beg[i] <= sb and se <= end[i]:
                st.extend(children(i)); continue
            c = props[i, :3]; size = box / 2.0 ** lvl[i]; mac_lh = size * size / theta2
            gap = np.maximum(0, np.maximum(lo - c, c - hi)); dmin2 = (gap ** 2).sum()
            if mac_lh < dmin2 * (1 - 2.0 ** -20):
                S["inter_shared"] += nT; S["shared_src"] += 1; continue
            if variant == "box":
                far = np.maximum(np.abs(lo - c), np.abs(hi - c)); dmax2 = (far ** 2).sum()
                rej_all = mac_lh >= dmax2 * (1 + 2.0 ** -20)
                acc_all = False
            elif variant == "gbox":
                # per-group bounding boxes: accepted by every group's box / failing at every group's farthest corner
                acc_all = rej_all = True
                for k in range(len(gs)):
                    glo, ghi = T[k].min(0), T[k].max(0)
                    gap = np.maximum(0, np.maximum(glo - c, c - ghi))
                    far = np.maximum(np.abs(glo - c), np.abs(ghi - c))
                    acc_all &= mac_lh < (gap ** 2).sum() * (1 - 2.0 ** -20)
                    rej_all &= mac_lh >= (far ** 2).sum() * (1 + 2.0 ** -20)
            else:
                # a node that is a group's own node or its ancestor is neither accepted nor tested by that group
                special = any(beg[i] <= beg[g] and end[g] <= end[i] for g in gs)
                if special:
                    rej_all = acc_all = False
                else:
                    fails = [bool((mac_lh >= ((T[k] - c) ** 2).sum(1)).any()) for k in range(len(gs))]
                    rej_all = all(fails); acc_all = not any(fails)
            if acc_all:
                S["inter_shared"] += nT; S["shared_src"] += 1
            elif rej_all:
                if nd[i] == 0:
                    S["inter_shared"] += nT * (end[i] - beg[i]); S["shared_src"] += end[i] - beg[i]
                else:
                    st.extend(children(i))
            else:
                frontier.append(i)
        S["frontier"] += len(frontier); S["fmax"] = max(S["fmax"], len(frontier))
        # phase 2
        for k, g in enumerate(gs):
            S["inter2"] = S.get("inter2", 0) + len(T[k]) * (len(T[k]) - 1)
            st = list(frontier)
            while st:
                i = st.pop(); S["p2_visits"] += 1
                if i == g: continue
                if beg[i] <= beg[g] and end[g] <= end[i]:
                    st.extend(children(i)); continue
                size = box / 2.0 ** lvl[i]; mac_lh = size * size / theta2
                d2 = ((T[k] - props[i, :3]) ** 2).sum(1)
                if (mac_lh < d2).all():
                    S["p2_src"] += 1; S["inter2"] = S.get("inter2", 0) + len(T[k])
                elif nd[i] == 0:
                    S["p2_src"] += end[i] - beg[i]; S["inter2"] = S.get("inter2", 0) + len(T[k]) * (end[i] - beg[i])
                else:
                    st.extend(children(i))
    S["inter2"] = S.get("inter2", 0) + S["inter_shared"]
    assert S["inter2"] == S["inter"], "the two-phase walk must reproduce the per-group interaction count exactly"
    print(variant, S)
    print(f"  visits per group: baseline {S['base_visits']/S['groups']:.0f}; two-phase {(S['p1_visits']+S['p2_visits'])/S['groups']:.0f}"
          f" (phase 1 {S['p1_visits']/S['runs']:.0f} per run, phase 2 {S['p2_visits']/S['groups']:.0f} per group)"
          f" ratio {(S['p1_visits']+S['p2_visits'])/S['base_visits']:.3f}")
    print(f"  frontier per run mean {S['frontier']/S['runs']:.0f} max {S['fmax']}; shared interactions share {S['inter_shared']/S['inter']:.3f};"
          f" shared sources/run {S['shared_src']/S['runs']:.0f}, phase-2 sources/group {S['p2_src']/S['groups']:.0f}")
