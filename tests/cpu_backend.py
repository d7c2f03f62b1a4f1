"""Numpy stand-in for rakau_b200.Octree, for the world-size-2 gloo tests of rakau_b200.distributed.ShardedTree.

Test infrastructure only: it lets the multi-GPU ORCHESTRATION (sample-sort splitters, bucket exchange, padded
all-gathers, Morton-range cuts, output exchange, cost rebalancing) run on CPU tensors. It implements the handful of
Octree methods ShardedTree calls, with a plain 21-bit-per-dimension Morton code, a stable argsort, critical nodes of
`ncrit` consecutive particles and a synthetic 'evaluation' whose result is a known function of each particle."""
import numpy as np
import torch


def morton_codes(x, y, z, box):
    def disc(a):
        v = np.floor((a.astype(np.float64) / box + 0.5) * (1 << 21)).astype(np.int64)
        return np.clip(v, 0, (1 << 21) - 1)

    def spread(v):
        out = np.zeros_like(v)
        for b in range(21):
            out |= ((v >> b) & 1) << (3 * b)
        return out
    return spread(disc(x)) | (spread(disc(y)) << 1) | (spread(disc(z)) << 2)


class _Info:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def asdict(self):
        return dict(self.__dict__)


class CpuOctree:
    def __init__(self, fp=32, mac="bh", device=None):
        self.F = np.float32 if fp == 32 else np.float64
        self.eval_info = _Info()
        self._costs = None

    @staticmethod
    def _np(t):
        return t.numpy() if isinstance(t, torch.Tensor) else np.asarray(t)

    # ---- sample-sort building blocks ----
    def sort_shard(self, x, y, z, m, n, box_size, codes=None):
        x, y, z, m = (self._np(a)[:n] for a in (x, y, z, m))
        c = morton_codes(x, y, z, box_size) if codes is None else self._np(codes)[:n].astype(np.int64)
        p = np.argsort(c, kind="stable")
        self._codes, self._parts, self._lp = c[p], [a[p].copy() for a in (x, y, z, m)], p.astype(np.int32)

    def codes_device(self, out):
        out.copy_(torch.from_numpy(self._codes))

    def parts_device(self, x, y, z, m):
        for o, a in zip((x, y, z, m), self._parts):
            o.copy_(torch.from_numpy(a))

    def perm_device(self, out, which=None):
        out.copy_(torch.from_numpy(self._lp))

    # ---- replicated "tree": critical nodes = blocks of ncrit consecutive particles ----
    def build_presorted(self, x, y, z, m, codes, perm, n, box_size, max_leaf_n=16, ncrit=128, parts_ready_event=None):
        assert parts_ready_event is None
        self._codes = self._np(codes)[:n].copy()
        self._parts = [self._np(a)[:n].copy() for a in (x, y, z, m)]
        self._perm = self._np(perm)[:n].copy()
        self.n = n
        # blocks of ncrit particles, shifted by a build counter so that a REBUILD changes the critical nodes
        self._nbuilds = getattr(self, "_nbuilds", 0) + 1
        first = (17 * (self._nbuilds - 1)) % ncrit
        self._begin = np.unique(np.concatenate([[0], np.arange(first, n, ncrit, dtype=np.int64)]))
        self._costs = np.zeros(self._begin.size, dtype=np.uint64)
        return _Info(n_nodes=int(self._begin.size), n_crit=int(self._begin.size), box_size=float(box_size))

    @property
    def ncrit_nodes(self):
        return int(self._begin.size)

    def crit(self):
        end = np.append(self._begin[1:], self.n)
        return np.stack([np.arange(self._begin.size), self._begin, end], 1).astype(np.uint64)

    def crit_begin_at(self, idx):
        b = np.append(self._begin, self.n)
        return b[np.asarray(idx, dtype=np.int64)].astype(np.uint64)

    def to_original_order(self, arrays, out):
        for a, o in zip(arrays, out):
            o[torch.from_numpy(self._perm.astype(np.int64))] = a
        return out

    def crit_lower_bound(self, particle_idx):
        return np.searchsorted(self._begin, np.asarray(particle_idx, dtype=np.int64), side="left").astype(np.uint64)

    @staticmethod
    def expected(parts, j):
        x, y, z, m = parts
        return (x * (j + 1) + y - z * 0.5 + m).astype(x.dtype)

    def acc_pot(self, Q, theta, G=1.0, eps=0.0, ordered=False, split=None, out=None, where=None, crit_range=None):
        c0, c1 = crit_range
        b = np.append(self._begin, self.n)
        pb, pe = int(b[c0]), int(b[c1])
        for j, o in enumerate(out):
            o[pb:pe] = torch.from_numpy(self.expected([a[pb:pe] for a in self._parts], j))
        sizes = (b[1:] - b[:-1])[c0:c1]
        cost = sizes * (1 + (np.arange(c0, c1) % 7))  # uneven per-group cost
        self._costs[c0:c1] = cost.astype(np.uint64)
        self.eval_info = _Info(mac_tests=0, accepted=0, p2p_pairs=0, self_pairs=0, interactions=int(cost.sum()),
                               n_groups=int(c1 - c0), kernel_launches=1, ms_kernel=1.0, ms_total=1.0)
        return out

    def group_costs(self):
        return self._costs.copy()
