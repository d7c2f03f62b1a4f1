import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import oracle
n = 1_000_000
m, x, y, z = oracle.plummer(n)
t = oracle.OracleTree(x, y, z, m, fp=32, mac="bh", max_leaf_n=16, ncrit=128)
nodes = t.nodes(); crit, _ = t.crit()
px, py, pz, pm = t.parts()
P = np.stack([px, py, pz], 1).astype(np.float64)
box = t.box_size; theta2 = 0.75 ** 2
beg = nodes["begin"].astype(np.int64); end = nodes["end"].astype(np.int64); nd = nodes["n_children"].astype(np.int64); lvl = nodes["level"].astype(np.int64)
props = nodes["props"].astype(np.float64)
M = len(nodes)
# parent of every node (DFS order)
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
# super-groups = critical nodes with the same parent
order = np.argsort(par, kind="stable")
uniq, start, cnt = np.unique(par[order], return_index=True, return_counts=True)
sizes = np.array([ (end[cn[order[s:s+c]]] - beg[cn[order[s:s+c]]]).sum() for s, c in zip(start, cnt)])
print("critical nodes", len(cn), "super-groups", len(uniq), "groups per super-group mean", cnt.mean(), "hist", np.bincount(cnt)[:10])
print("targets per super-group: mean", sizes.mean(), "max", sizes.max(), "quantiles", np.quantile(sizes, [0.5, 0.9, 0.99]))
rng = np.random.default_rng(2)
sel = rng.choice(len(uniq), 120, replace=False)
tot = dict(sep_tests=0, shared_tests=0, inter=0, masked_waste=0, mixed_nodes=0, src_entries=0)
for si in sel:
    gs = cn[order[start[si]:start[si] + cnt[si]]]
    G = len(gs)
    T = [P[beg[g]:end[g]] for g in gs]
    nT = np.array([len(a) for a in T]); totT = nT.sum()
    gset = set(int(g) for g in gs)
    # shared walk: stack of (node, mask)
    full = (1 << G) - 1
    st = [(0, full)]
    while st:
        i, mask = st.pop()
        tot["shared_tests"] += 1
        tot["sep_tests"] += bin(mask).count("1")
        c = props[i, :3]; size = box / 2.0 ** lvl[i]; mac_lh = size * size / theta2
        acc_mask = 0; desc_mask = 0
        for k in range(G):
            if not (mask >> k) & 1: continue
            g = int(gs[k])
            if i == g:
                continue  # own node: self interactions, handled separately
            if beg[i] <= beg[g] and end[g] <= end[i]:
                desc_mask |= 1 << k; continue  # ancestor
            d2 = ((T[k] - c) ** 2).sum(1)
            if (mac_lh < d2).all(): acc_mask |= 1 << k
            else: desc_mask |= 1 << k
        if acc_mask:
            na = sum(nT[k] for k in range(G) if (acc_mask >> k) & 1)
            tot["inter"] += na; tot["masked_waste"] += totT - na; tot["src_entries"] += 1
            if acc_mask == full: tot["inter_full"] = tot.get("inter_full", 0) + na; tot["src_full"] = tot.get("src_full", 0) + 1
            else: tot["src_part_appends"] = tot.get("src_part_appends", 0) + bin(acc_mask).count("1")
        if acc_mask and desc_mask: tot["mixed_nodes"] += 1
        if desc_mask:
            if nd[i] == 0:
                cntp = end[i] - beg[i]
                na = sum(nT[k] for k in range(G) if (desc_mask >> k) & 1)
                tot["inter"] += na * cntp; tot["masked_waste"] += (totT - na) * cntp; tot["src_entries"] += cntp
                if desc_mask == full: tot["inter_full"] = tot.get("inter_full", 0) + na * cntp; tot["src_full"] = tot.get("src_full", 0) + cntp
                else: tot["src_part_appends"] = tot.get("src_part_appends", 0) + bin(desc_mask).count("1") * cntp
            else:
                j = i + 1; e = i + 1 + nd[i]
                while j < e:
                    st.append((j, desc_mask)); j += nd[j] + 1
print(tot)
print("walk tests: separate", tot["sep_tests"], "shared", tot["shared_tests"], "ratio", tot["shared_tests"] / tot["sep_tests"])
print("masked (wasted) lane-interactions / useful", tot["masked_waste"] / tot["inter"])
print("useful interactions from full-mask sources:", tot["inter_full"] / tot["inter"], "full-mask sources", tot["src_full"], "partial-mask per-group appends", tot["src_part_appends"])
print("mixed nodes / shared tests", tot["mixed_nodes"] / tot["shared_tests"])
