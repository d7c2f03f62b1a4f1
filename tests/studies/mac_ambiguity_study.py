import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import oracle
n = 1_000_000
m, x, y, z = oracle.plummer(n)
t = oracle.OracleTree(x, y, z, m, fp=32, mac="bh", max_leaf_n=16, ncrit=128)
nodes = t.nodes(); crit, _ = t.crit()
px, py, pz, pm = t.parts()
P = np.stack([px, py, pz], 1).astype(np.float64)
box = t.box_size
theta2 = 0.75 ** 2
beg = nodes["begin"]; end = nodes["end"]; nd = nodes["n_children"]; lvl = nodes["level"]; props = nodes["props"].astype(np.float64)
rng = np.random.default_rng(1)
sel = rng.choice(len(crit), 200, replace=False)
tot = dict(tests=0, sure_acc=0, sure_rej=0, amb=0, amb_acc=0, amb_rej=0, h6=0, h14=0, hc=0, h1=0, h2c=0)
dirs14 = np.array([[1,0,0],[-1,0,0],[0,1,0],[0,-1,0],[0,0,1],[0,0,-1]] + [[a,b,c] for a in (1,-1) for b in (1,-1) for c in (1,-1)], dtype=np.float64)
for ci in sel:
    gnode, gb, ge = [int(v) for v in crit[ci]]
    T = P[gb:ge]
    lo, hi = T.min(0), T.max(0)
    sup14 = T[np.argmax(T @ dirs14.T, axis=0)]  # support points
    sup6 = sup14[:6]
    cen = 0.5 * (lo + hi)
    near_c = T[np.argmin(((T - cen) ** 2).sum(1))][None]
    stack = [0]
    while stack:
        i = stack.pop()
        if i == gnode:
            continue
        anc = beg[i] <= gb and ge <= end[i]
        if not anc:
            tot["tests"] += 1
            c = props[i, :3]
            size = box / 2.0 ** lvl[i]
            mac_lh = size * size / theta2
            a = lo - c; b = c - hi
            gap = np.maximum(0, np.maximum(a, b)); far = np.maximum(-a, -b)
            dmin2 = (gap ** 2).sum(); dmax2 = (far ** 2).sum()
            if mac_lh < dmin2:
                tot["sure_acc"] += 1; continue
            if mac_lh >= dmax2:
                tot["sure_rej"] += 1
                rej = True
            else:
                tot["amb"] += 1
                d2 = ((T - c) ** 2).sum(1)
                rej = bool((mac_lh >= d2).any())
                if rej:
                    tot["amb_rej"] += 1
                    if (mac_lh >= ((sup6 - c) ** 2).sum(1)).any(): tot["h6"] += 1
                    if (mac_lh >= ((sup14 - c) ** 2).sum(1)).any(): tot["h14"] += 1
                    if (mac_lh >= ((near_c - c) ** 2).sum(1)).any(): tot["hc"] += 1
                    # support point in the node's own direction (needs a per-node argmax: not free) for reference
                    d = c - cen
                    sp = T[np.argmax(T @ d)]
                    if mac_lh >= ((sp - c) ** 2).sum(): tot["h1"] += 1
                    octant = (4 if d[0] < 0 else 0) + (2 if d[1] < 0 else 0) + (1 if d[2] < 0 else 0)
                    ax = int(np.argmax(np.abs(d))); axc = ax * 2 + (1 if d[ax] < 0 else 0)
                    c2 = np.stack([sup14[6 + octant], sup14[axc]])
                    if (mac_lh >= ((c2 - c) ** 2).sum(1)).any(): tot["h2c"] += 1
                    if mac_lh >= ((sup14[6 + octant] - c) ** 2).sum(): tot["hoct"] = tot.get("hoct", 0) + 1
                    if mac_lh >= ((sup14[axc] - c) ** 2).sum(): tot["hax"] = tot.get("hax", 0) + 1
                else:
                    tot["amb_acc"] += 1; continue
            if not rej:
                continue
            if nd[i] == 0:
                continue
        # descend: children of i
        j = i + 1; e = i + 1 + nd[i]
        while j < e:
            stack.append(j); j += nd[j] + 1
print(tot)
for k in ("sure_acc", "sure_rej", "amb"): print(k, tot[k] / tot["tests"])
for k in ("amb_acc", "amb_rej"): print(k, tot[k] / tot["amb"])
for k in ("h6", "h14", "hc", "h1", "h2c", "hoct", "hax"): print(k, tot[k] / tot["amb_rej"])
