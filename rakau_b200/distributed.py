"""Multi-GPU orchestration (one process per GPU, torch.distributed / NCCL over NVLink).

The reference has no distributed mode (its multi-GPU split is one process driving several devices through
managed memory, src/rakau_cuda.cu:434-527, and its README admits poor scaling). Here:

* build: distributed sample sort. Every rank Morton-encodes and sorts its 1/P shard with the global box, the
  ranks agree on P-1 splitter codes from regular samples, exchange buckets with one all-to-all per array,
  sort their bucket (P sorted runs, rank order => the stable order of the single-GPU path), all-gather the
  sorted buckets and build the replicated tree from the globally sorted arrays. Per-rank sort work is N/P
  instead of N.
* traversal: contiguous Morton ranges of critical nodes, cut by the previous evaluation's interaction counts
  (rakau_b200.sharding); every rank broadcasts the output slice it owns.
"""
import numpy as np

from . import RK_DEVICE, RK_LAST_PERM, Octree, deduce_box, sharding


class ShardedTree:
    def __init__(self, dist, device, fp=32, mac="bh", max_leaf_n=16, ncrit=128, samples_per_rank=2048):
        import torch
        self.torch, self.dist, self.dev = torch, dist, device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.fp, self.mln, self.ncrit = fp, max_leaf_n, ncrit
        self.dt = torch.float32 if fp == 32 else torch.float64
        idx = device.index
        self.local, self.bucket, self.tree = (Octree(fp=fp, mac=mac, device=idx) for _ in range(3))
        s = torch.cuda.current_stream().cuda_stream
        for t in (self.local, self.bucket, self.tree):
            t.set_stream(s)
        self.nsamp = samples_per_rank
        self.cuts = None
        self.cut_particles = None  # first particle of every rank's range

    # ---- build -------------------------------------------------------------------------------------------
    def build(self, x, y, z, m, first_index):
        """x, y, z, m: this rank's shard (device tensors); first_index: global index of its first particle."""
        torch, dist = self.torch, self.dist
        n_loc = x.numel()
        amax = torch.stack([x.abs().max(), y.abs().max(), z.abs().max()]).max().double().reshape(1)
        dist.all_reduce(amax, op=dist.ReduceOp.MAX)
        box = deduce_box(float(amax.item()), self.fp)
        # 1. local sort of the shard
        self.local.sort_shard(x, y, z, m, n_loc, box)
        rows = self._rows_from_tree(self.local, n_loc, offset=int(first_index))
        codes = rows[:, 0:2].contiguous().view(torch.int64).reshape(-1)
        # 2. splitters from regular samples (codes are < 2^63, so int64 order == unsigned order)
        pos = (torch.arange(self.nsamp, device=self.dev, dtype=torch.int64) * max(n_loc - 1, 0)) // max(self.nsamp - 1, 1)
        samp = codes[pos] if n_loc else torch.full((self.nsamp,), 2 ** 62, dtype=torch.int64, device=self.dev)
        allsamp = torch.empty(self.nsamp * self.world, dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(allsamp, samp)
        allsamp, _ = torch.sort(allsamp)
        split = allsamp[torch.arange(1, self.world, device=self.dev) * self.nsamp]
        # 3. bucket exchange: ONE all-to-all of 28-byte rows (code, x, y, z, m, original index)
        bounds = torch.searchsorted(codes, split, right=False)
        bounds = torch.cat([torch.zeros(1, dtype=torch.int64, device=self.dev), bounds,
                            torch.tensor([n_loc], dtype=torch.int64, device=self.dev)])
        send = (bounds[1:] - bounds[:-1])
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)
        send_l, recv_l = send.tolist(), recv.tolist()
        n_b = int(sum(recv_l))
        brows = torch.empty((n_b, 7), dtype=torch.int32, device=self.dev)
        dist.all_to_all_single(brows, rows, output_split_sizes=recv_l, input_split_sizes=send_l)
        # 4. sort the bucket (runs arrive in rank order => stable order of the single-GPU path)
        bc, bx, by, bz, bm, bi = self._cols(brows)
        self.bucket.sort_shard(bx, by, bz, bm, n_b, box, codes=bc)
        srows = self._rows_from_tree(self.bucket, n_b, gidx=bi)
        # 5. ONE all-gather of the sorted buckets, padded to the largest bucket
        sizes = torch.empty(self.world, dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(sizes, torch.tensor([n_b], dtype=torch.int64, device=self.dev))
        sizes_l = sizes.tolist()
        n, mx = int(sum(sizes_l)), int(max(sizes_l))
        if n_b < mx:
            srows = torch.cat([srows, torch.zeros((mx - n_b, 7), dtype=torch.int32, device=self.dev)])
        allrows = torch.empty((self.world, mx, 7), dtype=torch.int32, device=self.dev)
        dist.all_gather_into_tensor(allrows, srows)
        full = torch.cat([allrows[r, :sizes_l[r]] for r in range(self.world)]) if min(sizes_l) < mx \
            else allrows.reshape(-1, 7)
        fc, fx, fy, fz, fm, fi = self._cols(full)
        # 6. replicated topology + node properties
        self.n = n
        self.full_sorted = (fx, fy, fz, fm)  # keep alive
        bi_ = self.tree.build_presorted(fx, fy, fz, fm, fc, fi, n, box, self.mln, self.ncrit)
        self.cut_particles = None
        return bi_

    def _rows_from_tree(self, t, n, offset=0, gidx=None):
        """[n, 7] int32 rows (code lo, code hi, x, y, z, m, original index) of a sorted shard / bucket."""
        torch = self.torch
        rows = torch.empty((n, 7), dtype=torch.int32, device=self.dev)
        codes = torch.empty(n, dtype=torch.int64, device=self.dev)
        cols = [torch.empty(n, dtype=self.dt, device=self.dev) for _ in range(4)]
        lp = torch.empty(n, dtype=torch.int32, device=self.dev)
        t.codes_device(codes)
        t.parts_device(*cols)
        t.perm_device(lp, RK_LAST_PERM)
        rows[:, 0:2] = codes.view(torch.int32).reshape(n, 2)
        for j in range(4):
            rows[:, 2 + j] = cols[j].view(torch.int32)
        rows[:, 6] = (lp + offset) if gidx is None else gidx[lp.long()]
        return rows

    def _cols(self, rows):
        torch = self.torch
        n = rows.shape[0]
        codes = rows[:, 0:2].contiguous().view(torch.int64).reshape(n)
        x, y, z, m = (rows[:, 2 + j].contiguous().view(self.dt) for j in range(4))
        return codes, x, y, z, m, rows[:, 6].contiguous()

    # ---- traversal -----------------------------------------------------------------------------------------
    def _ensure_cuts(self):
        C = self.tree.ncrit_nodes
        if self.cuts is None or self.cuts[-1] != C:
            # first evaluation of this tree shape: equal particle counts (tree.hpp:3147-3178)
            cr = self.tree.crit()[:, 1].astype(np.int64)
            self.cuts = sharding.cuts_by_particles(cr, self.n, self.world)
            self.cut_particles = None
        if self.cut_particles is None:
            self.cut_particles = self.tree.crit_begin_at(self.cuts).astype(np.int64)

    def acc_pot(self, Q, theta, out, G=1.0, eps=0.0, exchange=True):
        """Evaluate this rank's Morton range into `out` (device tensors of n elements, Morton order); with
        exchange=True every rank ends with the full result."""
        self._ensure_cuts()
        c0, c1 = self.cuts[self.rank], self.cuts[self.rank + 1]
        self.tree.acc_pot(Q, theta, G=G, eps=eps, out=out, where=RK_DEVICE, crit_range=(c0, c1))
        info = self.tree.eval_info.asdict()
        if exchange:
            # every rank owns one contiguous slice: zero the rest and sum (x + 0 is exact), one collective per
            # output array instead of world x broadcasts
            pb, pe = int(self.cut_particles[self.rank]), int(self.cut_particles[self.rank + 1])
            for o in out:
                o[:pb].zero_()
                o[pe:].zero_()
                self.dist.all_reduce(o)
        return info

    def rebalance(self):
        """Cost-weighted cuts from the last evaluation's per-group interaction counts."""
        costs = sharding.allreduce_costs(self.tree.group_costs(), self.dist, self.dev)
        self.cuts = sharding.cuts_by_cost(costs, self.world)
        self.cut_particles = None
        return sharding.imbalance(costs, self.cuts)
