"""Multi-GPU orchestration (one process per GPU, torch.distributed / NCCL over NVLink).

The reference has no distributed mode (its multi-GPU split is one process driving several devices through
managed memory, src/rakau_cuda.cu:434-527, and its README admits poor scaling). Here:

* build: distributed sample sort. Every rank Morton-encodes and sorts its 1/P shard with the global box, the
  ranks agree on P-1 splitter codes from regular samples, exchange buckets with one all-to-all per array,
  sort their bucket (P sorted runs, rank order => the stable order of the single-GPU path), all-gather the
  sorted buckets and build the replicated tree from the globally sorted arrays. Per-rank sort work is N/P
  instead of N.
* traversal: contiguous Morton ranges of critical nodes, cut by the previous evaluation's interaction counts
  (rakau_b200.sharding); the output slices are all-gathered (padded to the largest slice).
"""
import os

import numpy as np

from . import RK_DEVICE, RK_LAST_PERM, Octree, deduce_box, sharding


def _multicast_ptr(handle):
    """NVSwitch multicast address of a symmetric-memory buffer (0: the platform has none)."""
    try:
        return int(getattr(handle, "multicast_ptr", 0) or 0)
    except Exception:  # noqa: BLE001
        return 0


def world_fits_mirrors(world):
    """rk_tree_set_output_mirrors holds 8 mirrors: the peers of the rank plus, possibly, one pinned host slice."""
    return world - 1 + 1 <= 8


class ShardedTree:
    def __init__(self, dist, device, fp=32, mac="bh", max_leaf_n=16, ncrit=128, samples_per_rank=2048,
                 octree_factory=None):
        """octree_factory(fp=, mac=, device=): the per-rank tree backend, rakau_b200.Octree unless given. The CPU
        tests (tests/test_sharding_gloo.py) run this class over gloo with a numpy stand-in for the backend, so the
        orchestration below - splitters, bucket exchange, gathers, range cuts, output exchange - is exercised
        without a GPU; the product always uses the CUDA library."""
        import torch
        self.torch, self.dist, self.dev = torch, dist, device
        self.cuda = device.type == "cuda"
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.fp, self.mln, self.ncrit = fp, max_leaf_n, ncrit
        self.dt = torch.float32 if fp == 32 else torch.float64
        make = octree_factory or Octree
        self.local, self.bucket, self.tree = (make(fp=fp, mac=mac, device=device.index) for _ in range(3))
        if self.cuda:
            s = torch.cuda.current_stream().cuda_stream
            for t in (self.local, self.bucket, self.tree):
                t.set_stream(s)
        self.nsamp = samples_per_rank
        self._bufs = {}
        self.cuts = None           # critical-node index where every rank's range starts (world + 1 entries)
        self.cut_particles = None  # first particle of every rank's range
        # The cuts are remembered as PARTICLE indices: a rebuild with moved particles changes the critical nodes, so
        # the cost-weighted boundaries are re-snapped to the new critical nodes (as tree.hpp:3168-3178 snaps the
        # reference's split) instead of being thrown away.
        self.cut_pidx = None
        self._build_id, self._cuts_build_id = 0, -1
        # output exchange by stores from inside the traversal kernel (True) or by copy-engine pushes of 4 chunked launches
        self.mirror_exchange = os.environ.get("RK_MIRROR_EXCHANGE", "1") != "0" and world_fits_mirrors(self.world)
        # NVSwitch multicast stores where the platform has them and there are enough peers for them to pay: between two
        # ranks a multicast store is just a slower store (measured: codes gather 1.24 ms against 0.51 ms for the copy
        # engine, evaluation + exchange 18.45 against 18.28 ms)
        mc_env = os.environ.get("RK_MULTICAST", "auto")
        self.multicast = mc_env == "1" or (mc_env == "auto" and self.world > 2)
        self.multicast_codes = os.environ.get("RK_MULTICAST_CODES", "0") == "1"
        self.exchange_mode = self.codes_gather_mode = None
        self._side = None  # stream of the particle all-gather that runs underneath the topology build
        self._push = None  # per-peer copy streams of the output exchange
        self._peer = None  # two sets of (capacity, buffers, per-buffer list of every rank's device pointer), or False
        self._peer_flip = 0
        self._bar = None
        self._bpeer = None  # symmetric-memory bucket arrays of the exchange, or False
        self._gpeer = None  # symmetric-memory full arrays of the build (codes, x, y, z, m, original index), or False

    # ---- build -------------------------------------------------------------------------------------------
    def build(self, x, y, z, m, first_index):
        """x, y, z, m: this rank's shard (device tensors); first_index: global index of its first particle."""
        torch, dist = self.torch, self.dist
        n_loc = x.numel()
        self._ev = [('start', self._rec())]
        if n_loc:
            amax = torch.stack([x.abs().max(), y.abs().max(), z.abs().max()]).max().double().reshape(1)
        else:  # a rank may hold no particles at all
            amax = torch.zeros(1, dtype=torch.float64, device=self.dev)
        dist.all_reduce(amax, op=dist.ReduceOp.MAX)
        box = deduce_box(float(amax.item()), self.fp)
        partition = hasattr(self.local, "partition_shard") and self.world <= 256
        if partition:
            # 1. no local pre-sort: encode only; the splitter buckets are formed by ONE stable radix pass below
            self.local.encode_shard(x, y, z, m, n_loc, box)
            codes = torch.empty(n_loc, dtype=torch.int64, device=self.dev)
            self.local.codes_device(codes)
        else:
            # 1. local sort of the shard
            self.local.sort_shard(x, y, z, m, n_loc, box)
            codes, sx, sy, sz, sm, lp = self._sorted_arrays(self.local, n_loc)
            gidx = lp + int(first_index)  # original (global) index of each locally sorted particle
        self._ev.append(('local_sort', self._rec()))
        # 2. splitters from regular samples (codes are < 2^63, so int64 order == unsigned order); of the input order
        # when the shard was only encoded - particles arrive unordered, so that is a random sample
        pos = (torch.arange(self.nsamp, device=self.dev, dtype=torch.int64) * max(n_loc - 1, 0)) // max(self.nsamp - 1, 1)
        samp = codes[pos] if n_loc else torch.full((self.nsamp,), 2 ** 62, dtype=torch.int64, device=self.dev)
        allsamp = torch.empty(self.nsamp * self.world, dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(allsamp, samp)
        allsamp, _ = torch.sort(allsamp)
        split = allsamp[torch.arange(1, self.world, device=self.dev) * self.nsamp].contiguous()
        self._ev.append(('splitters', self._rec()))
        # 3. bucket exchange (the arrays stay SoA and contiguous)
        if partition:
            cnt = self.local.partition_shard(split)  # bucket r = codes in [split[r-1], split[r]), input order kept
            codes, sx, sy, sz, sm, lp = self._sorted_arrays(self.local, n_loc)
            gidx = lp + int(first_index)
            bounds = torch.from_numpy(np.concatenate([[0], np.cumsum(cnt.astype(np.int64))])).to(self.dev)
        else:
            bounds = torch.searchsorted(codes, split, right=False)
            bounds = torch.cat([torch.zeros(1, dtype=torch.int64, device=self.dev), bounds,
                                torch.tensor([n_loc], dtype=torch.int64, device=self.dev)])
        send = (bounds[1:] - bounds[:-1])
        exchanged = self._push_a2a(send, bounds, (codes, sx, sy, sz, sm, gidx)) if self.cuda and self.world > 1 else None
        if exchanged is not None:
            n_b, (bc, bx, by, bz, bm, bi) = exchanged
        else:
            recv = torch.empty_like(send)
            dist.all_to_all_single(recv, send)
            send_l, recv_l = send.tolist(), recv.tolist()
            n_b = int(sum(recv_l))

            def a2a(t):
                out = torch.empty(n_b, dtype=t.dtype, device=self.dev)
                dist.all_to_all_single(out, t, output_split_sizes=recv_l, input_split_sizes=send_l)
                return out
            bc, bx, by, bz, bm, bi = (a2a(t) for t in (codes, sx, sy, sz, sm, gidx))
        self._ev.append(('all_to_all', self._rec()))
        # 4. sort the bucket (runs arrive in rank order => stable order of the single-GPU path)
        self.bucket.sort_shard(bx, by, bz, bm, n_b, box, codes=bc)
        bc, bx, by, bz, bm, blp = self._sorted_arrays(self.bucket, n_b)
        bi = bi[blp.long()]
        self._ev.append(('bucket_sort', self._rec()))
        # 5. uneven all-gather of the sorted buckets straight into the full arrays (grouped NCCL send/recv)
        sizes = torch.empty(self.world, dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(sizes, torch.tensor([n_b], dtype=torch.int64, device=self.dev))
        offs = np.concatenate([[0], np.cumsum(sizes.tolist())]).astype(np.int64)
        n = int(offs[-1])
        # padded all_gather_into_tensor (NCCL's tuned all-gather: ~600 GB/s per rank on this NVSwitch box, the
        # grouped send/recv form measured 2x slower here) + one compaction pass
        sizes_l = [int(offs[r + 1] - offs[r]) for r in range(self.world)]
        mx = max(sizes_l)

        def gather(key, t):
            pad = self._persistent('p' + key, mx, t.dtype)
            pad[:n_b] = t
            allb = self._persistent('a' + key, mx * self.world, t.dtype)
            dist.all_gather_into_tensor(allb, pad)
            if min(sizes_l) == mx:
                return allb
            full = self._persistent('f' + key, n, t.dtype)
            for r in range(self.world):
                full[int(offs[r]):int(offs[r + 1])] = allb[r * mx:r * mx + sizes_l[r]]
            return full
        # The topology needs only the codes: gather them first, then gather the particle arrays on a side stream
        # while rk_tree_build_presorted builds the topology on this one (it waits for `ready` before the node
        # properties).
        pushed = self._push_gather(n, n_b, offs, bc, (bx, by, bz, bm, bi)) if self.cuda and self.world > 1 else None
        if pushed is not None:
            fc, (fx, fy, fz, fm, fi), ready = pushed
            self._ev.append(('all_gather_codes', self._rec()))
            self.n = n
            self.full_sorted = (fx, fy, fz, fm)
            bi_ = self.tree.build_presorted(fx, fy, fz, fm, fc, fi, n, box, self.mln, self.ncrit,
                                            parts_ready_event=ready.cuda_event)
            torch.cuda.current_stream().wait_stream(self._side)
            self._ev.append(('topology_props_and_particle_gather', self._rec()))
            self.cut_particles = None
            self._build_id += 1
            return bi_
        fc = gather('c', bc)
        self._ev.append(('all_gather_codes', self._rec()))
        ready = None
        if self.cuda:
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.dev)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                fx, fy, fz, fm, fi = (gather(k, t) for k, t in zip('xyzmi', (bx, by, bz, bm, bi)))
                ready = torch.cuda.Event()
                ready.record(self._side)
        else:
            fx, fy, fz, fm, fi = (gather(k, t) for k, t in zip('xyzmi', (bx, by, bz, bm, bi)))
        # 6. replicated topology + node properties
        self.n = n
        self.full_sorted = (fx, fy, fz, fm)  # keep alive
        bi_ = self.tree.build_presorted(fx, fy, fz, fm, fc, fi, n, box, self.mln, self.ncrit,
                                        parts_ready_event=ready.cuda_event if ready is not None else None)
        if self.cuda:
            main.wait_stream(self._side)
        self._ev.append(('topology_props_and_particle_gather', self._rec()))
        self.cut_particles = None
        self._build_id += 1
        return bi_

    def _push_a2a(self, send, bounds, arrays):
        """Bucket exchange with copy-engine pushes: rank s writes its slice for rank r straight into r's bucket arrays
        (symmetric memory) at the offset of the runs of the ranks before s, so the runs arrive in rank order as the
        stable order needs. One all-gather of the P x P count matrix, one stream-ordered barrier. Returns None when
        peer memory is not available."""
        torch, dist = self.torch, self.dist
        if self._bpeer is False:
            return None
        counts = torch.empty(self.world * self.world, dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(counts, send.contiguous())
        M = np.asarray(counts.tolist(), dtype=np.int64).reshape(self.world, self.world)  # M[s][r]: s sends r
        n_b_all = M.sum(axis=0)
        need = int(n_b_all.max())
        if self._bpeer is None or self._bpeer['cap'] < need:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                cap = int(need * 1.25) + 16  # the same value on every rank: the rendezvous is collective
                b = {'cap': cap}
                for k, t in zip('cxyzmi', arrays):
                    buf = symm_mem.empty(cap, dtype=t.dtype, device=self.dev)
                    h = symm_mem.rendezvous(buf, dist.group.WORLD)
                    b[k] = (buf, [int(p) for p in h.buffer_ptrs])
                self._bpeer = b
            except Exception as exc:  # noqa: BLE001
                self._bpeer = False
                self._peer_error = repr(exc)
                return None
        from . import device_copy_async
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        if self._push is None:
            self._push = [torch.cuda.Stream(device=self.dev) for _ in range(self.world)]
        for s in self._push:
            s.wait_stream(main)
        bnd = [int(v) for v in bounds.tolist()]
        off_at = np.cumsum(M, axis=0) - M  # off_at[s][r]: where the run of s starts in the bucket of r
        out = []
        for k, t in zip('cxyzmi', arrays):
            buf, ptrs = self._bpeer[k]
            esz = t.element_size()
            for d in range(self.world):
                r = (self.rank + d) % self.world
                cnt = int(M[self.rank][r])
                if cnt:
                    device_copy_async(ptrs[r] + int(off_at[self.rank][r]) * esz, t.data_ptr() + bnd[r] * esz, cnt * esz,
                                      self._push[r].cuda_stream)
            out.append(buf[:int(n_b_all[self.rank])])
        for s in self._push:
            self._side.wait_stream(s)
        self._stream_barrier(self._side)
        main.wait_stream(self._side)
        self._keep_alive_a2a = arrays
        return int(n_b_all[self.rank]), tuple(out)

    def _push_gather(self, n, n_b, offs, bc, parts):
        """Gather of the sorted buckets with COPY-ENGINE pushes into peer memory instead of NCCL all-gathers: every
        rank copies its bucket into the same slice of every peer's (symmetric-memory) full arrays, one stream per peer,
        the codes first. An NCCL all-gather kernel competes with the topology kernels for the SMs (measured in round 1:
        12.8 + 7.5 ms run one after the other, 18.4 ms overlapped); the copy engines do not. Two tiny barriers (NCCL, on
        the side stream) tell the ranks that all codes / all particle arrays have arrived. Returns None when peer
        memory is not available (then the collective path is used).
        Reuse of the buffers across steps is safe: a rank starts pushing step k+1 only after the barrier that ends the
        evaluation of step k, by which time every peer has finished reading the arrays of step k."""
        torch, dist = self.torch, self.dist
        if self._gpeer is False:
            return None
        keys = ('c', 'x', 'y', 'z', 'm', 'i')
        dts = (torch.int64, self.dt, self.dt, self.dt, self.dt, torch.int32)
        if self._gpeer is None or self._gpeer['cap'] < n:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                cap = int(n * 1.02) + 16
                g = {'cap': cap}
                for k, dtp in zip(keys, dts):
                    t = symm_mem.empty(cap, dtype=dtp, device=self.dev)
                    h = symm_mem.rendezvous(t, dist.group.WORLD)
                    g[k] = (t, [int(p) for p in h.buffer_ptrs])
                    g['mc_' + k] = _multicast_ptr(h)
                self._gpeer = g
            except Exception as exc:  # noqa: BLE001
                self._gpeer = False
                self._peer_error = repr(exc)
                return None
        from . import device_copy_async
        g = self._gpeer
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        if self._push is None:
            self._push = [torch.cuda.Stream(device=self.dev) for _ in range(self.world)]
        side = self._side
        for s in self._push:
            s.wait_stream(main)
        off = int(offs[self.rank])

        def push(key, t):
            buf, ptrs = g[key]
            esz = t.element_size()
            for d in range(self.world):  # d = 0: this rank's own copy; then staggered over the peers
                r = (self.rank + d) % self.world
                if n_b:
                    device_copy_async(ptrs[r] + off * esz, t.data_ptr(), n_b * esz, self._push[r].cuda_stream)
            return buf[:n]

        if self.multicast_codes and g['mc_c'] and n_b:
            # the codes: ONE kernel stores this rank's bucket through the NVSwitch multicast address of the full array -
            # every word leaves the GPU once and lands in every rank's copy. Measured at 8 GPUs (128 MB per rank): 3.1 ms,
            # plain stores to each of the 7 peers 4.3 ms, the copy engines 2.6 ms - so this is off unless asked for.
            from . import device_bcast_copy
            side.wait_stream(main)
            device_bcast_copy([g['mc_c'] + off * 8], bc.data_ptr(), n_b * 8, side.cuda_stream, multicast=True)
            fc = g['c'][0][:n]
            self.codes_gather_mode = "multicast kernel"
        else:
            self.codes_gather_mode = "copy-engine pushes"
            fc = push('c', bc)
            for s in self._push:
                side.wait_stream(s)
        self._stream_barrier(side)
        codes_ready = torch.cuda.Event()
        codes_ready.record(side)
        full = tuple(push(k, t) for k, t in zip(keys[1:], parts))
        for s in self._push:
            side.wait_stream(s)
        self._stream_barrier(side)
        ready = torch.cuda.Event()
        ready.record(side)
        main.wait_event(codes_ready)
        self._keep_alive = (bc, parts)  # the sources of the pushes must outlive them
        return fc, full, ready

    def _stream_barrier(self, stream):
        """A barrier in STREAM order that does not block the host: a one-element all-reduce on `stream`. (dist.barrier()
        synchronises the device, so the host could not enqueue the work that is meant to overlap the copies: measured
        at 8 GPUs, the particle pushes then ran before the topology kernels instead of underneath them.)"""
        if self._bar is None:
            self._bar = self.torch.zeros(1, dtype=self.torch.float32, device=self.dev)
        with self.torch.cuda.stream(stream):
            self.dist.all_reduce(self._bar)

    def _rec(self):
        if not self.cuda:
            return None
        e = self.torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def phase_ms(self):
        """CUDA-event time of the build phases of the last build() on this rank."""
        if not self.cuda:
            return {}
        self.torch.cuda.synchronize()
        return {self._ev[i][0]: self._ev[i - 1][1].elapsed_time(self._ev[i][1]) for i in range(1, len(self._ev))}

    def _sorted_arrays(self, t, n):
        """codes, x, y, z, m, last_perm of a sorted shard / bucket as contiguous device tensors."""
        torch = self.torch
        codes = torch.empty(n, dtype=torch.int64, device=self.dev)
        cols = [torch.empty(n, dtype=self.dt, device=self.dev) for _ in range(4)]
        lp = torch.empty(n, dtype=torch.int32, device=self.dev)
        t.codes_device(codes)
        t.parts_device(*cols)
        t.perm_device(lp, RK_LAST_PERM)
        return (codes, *cols, lp)

    def _persistent(self, key, n, dtype):
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < n or buf.dtype != dtype:
            buf = self.torch.empty(int(n * 1.02) + 16, dtype=dtype, device=self.dev)
            self._bufs[key] = buf
        return buf[:n]

    def _allgather_uneven(self, mine, offs, full=None, copy_in=False):
        """All-gather of unequal contiguous slices: rank r owns full[offs[r]:offs[r+1]]. One grouped NCCL
        send/recv round (every pair exchanges directly over NVLink), no padding and no staging copies."""
        torch, dist = self.torch, self.dist
        if full is None or copy_in:
            if full is None:
                full = torch.empty(int(offs[-1]), dtype=mine.dtype, device=self.dev)
            full[int(offs[self.rank]):int(offs[self.rank + 1])] = mine
            mine = full[int(offs[self.rank]):int(offs[self.rank + 1])]
        ops = []
        for r in range(self.world):
            if r == self.rank:
                continue
            if mine.numel():
                ops.append(dist.P2POp(dist.isend, mine, r))
            if offs[r + 1] > offs[r]:
                ops.append(dist.P2POp(dist.irecv, full[int(offs[r]):int(offs[r + 1])], r))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return full

    # ---- traversal -----------------------------------------------------------------------------------------
    def _snap(self, pidx):
        """Critical-node cuts for the given particle boundaries (first critical node starting at or after each)."""
        inner = [int(v) for v in pidx[1:-1]]
        cuts = []
        for i in range(0, len(inner), 16):  # rk_tree_crit_lower_bound takes 16 values per call
            cuts += [int(v) for v in self.tree.crit_lower_bound(inner[i:i + 16])]
        return [0] + cuts + [int(self.tree.ncrit_nodes)]

    def ensure_cuts(self):
        """Range cuts (self.cuts, self.cut_particles) for the current tree; called by acc_pot."""
        if self.cut_pidx is None or int(self.cut_pidx[-1]) != self.n or len(self.cut_pidx) != self.world + 1:
            # first evaluation: equal particle counts (tree.hpp:3147-3178)
            self.cut_pidx = [(r * self.n) // self.world for r in range(self.world)] + [self.n]
            self._cuts_build_id = -1
        if self._cuts_build_id != self._build_id:
            self.cuts = self._snap(self.cut_pidx)
            self.cut_particles = self.tree.crit_begin_at(self.cuts).astype(np.int64)
            self._cuts_build_id = self._build_id

    def _peer_outputs(self, nres):
        """Result buffers in CUDA peer memory (torch symmetric memory: every rank maps every other rank's buffer),
        allocated once for n particles. None if the platform cannot provide it (then NCCL does the exchange)."""
        if self._peer is False:
            return None
        if self._peer is None or self._peer[0][0] < self.n or len(self._peer[0][1]) < nres:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                cap = int(self.n * 1.02) + 16
                sets = []
                self._peer_mc = []
                for _ in range(2):
                    bufs, ptrs, mcs = [], [], []
                    for _ in range(max(nres, 3)):
                        t = symm_mem.empty(cap, dtype=self.dt, device=self.dev)
                        h = symm_mem.rendezvous(t, self.dist.group.WORLD)
                        bufs.append(t)
                        ptrs.append([int(p) for p in h.buffer_ptrs])
                        mcs.append(_multicast_ptr(h))
                    sets.append((cap, bufs, ptrs))
                    self._peer_mc.append(mcs)
                self._peer = sets
            except Exception as exc:  # noqa: BLE001 - no peer memory on this platform: use the collective path
                self._peer = False
                self._peer_error = repr(exc)
                return None
        # Two sets, alternating between evaluations: a fast rank pushes the slices of evaluation k+1 into set B while a
        # slow peer's kernels may still read the results of evaluation k in set A. Set A is written again by
        # evaluation k+2, whose pushes start after the barrier that ends evaluation k+1 - and every rank enters that
        # barrier in stream order behind its readers of set A.
        self._peer_flip ^= 1
        return self._peer[self._peer_flip]

    def acc_pot(self, Q, theta, out=None, G=1.0, eps=0.0, exchange=True, chunks=(0.4, 0.7, 0.9), host_out=None,
                ordered=False):
        """ordered=True: after the exchange every rank re-orders the complete result from the Morton order to the
        ORIGINAL (global) particle order, as the reference's `_o` functions do (tree.hpp:3320-3330); see _acc_pot."""
        info, res = self._acc_pot(Q, theta, out=out, G=G, eps=eps, exchange=exchange, chunks=chunks, host_out=host_out)
        if ordered:
            assert exchange, "the original order needs the complete result"
            dst = [self._persistent('ord%d' % j, self.n, self.dt) for j in range(len(res))]
            self.tree.to_original_order([r[:self.n] for r in res], dst)
            res = dst
        return info, res

    def _acc_pot(self, Q, theta, out=None, G=1.0, eps=0.0, exchange=True, chunks=(0.4, 0.7, 0.9), host_out=None):
        """Evaluate this rank's Morton range; with exchange=True every rank ends with the full result (Morton order).
        Returns (info, outputs): outputs are `out` if given, else library-owned peer-memory buffers.

        Peer-memory path (out=None): the range is evaluated in 4 launches (40/30/20/10 % of its critical nodes); as
        soon as a launch has finished, its contiguous output slice is pushed into every peer's buffer with
        copy-engine copies on a side stream (rk_device_copy_async) while the next launch occupies the SMs, so only the
        last tenth of the exchange is exposed. An NCCL kernel could not overlap: the traversal's persistent CTAs
        fill every SM. Collective path (out given, or no peer memory): padded all-gather of the owned slices.
        host_out: optional pinned host tensors (one per result, at least as long as this rank's particle range); the
        slice of every finished launch is copied to them while the next launch runs (peer-memory path)."""
        torch, dist = self.torch, self.dist
        self.ensure_cuts()
        c0, c1 = int(self.cuts[self.rank]), int(self.cuts[self.rank + 1])
        nres = {0: 3, 1: 1, 2: 4}[Q]
        peer = self._peer_outputs(nres) if (out is None and exchange and self.world > 1 and self.cuda) else None
        if peer is None:
            if out is None:
                out = [self._persistent('o%d' % j, self.n, self.dt) for j in range(nres)]
            self.tree.acc_pot(Q, theta, G=G, eps=eps, out=out, where=RK_DEVICE, crit_range=(c0, c1))
            info = self.tree.eval_info.asdict()
            if exchange and self.world > 1:
                # padded all-gather of the slices (one tuned NCCL collective per array) + one copy per peer slice
                cp = [int(v) for v in self.cut_particles]
                sizes = [cp[r + 1] - cp[r] for r in range(self.world)]
                mx = max(sizes)
                pb, pe = cp[self.rank], cp[self.rank + 1]
                for j, o in enumerate(out):
                    pad = self._persistent('xp%d' % j, mx, o.dtype)
                    pad[:pe - pb] = o[pb:pe]
                    allb = self._persistent('xa%d' % j, mx * self.world, o.dtype)
                    dist.all_gather_into_tensor(allb, pad)
                    for r in range(self.world):
                        if r != self.rank and sizes[r]:
                            o[cp[r]:cp[r + 1]] = allb[r * mx:r * mx + sizes[r]]
            return info, out
        _, bufs, ptrs = peer
        out = [b[:self.n] for b in bufs[:nres]]
        mc = self._peer_mc[self._peer_flip][:nres] if self.multicast else []
        use_mc = bool(mc) and all(mc)
        # (stores to 7 peers from inside the kernel cost more than they save - measured at 8 GPUs: the kernel 23.5 ms
        # instead of 19.3 - so without multicast the in-kernel exchange is used between two ranks only)
        if self.mirror_exchange and (use_mc or self.world <= 2):
            # ONE launch over the whole range (a single tail, work stealing included) whose kernel stores every final
            # result into its own buffer AND into each peer's (and into the pinned host slice, if asked for): the
            # exchange rides on NVLink store by store underneath the arithmetic (rk_tree_set_output_mirrors).
            esz = out[0].element_size()
            if use_mc:
                mirrors, mask = [list(mc)], 1  # one multicast store per result: the switch writes every rank's copy
            else:
                mirrors, mask = [[ptrs[j][r] for j in range(nres)] for r in range(self.world) if r != self.rank], 0
            if host_out is not None:
                pb = int(self.cut_particles[self.rank])
                mirrors.append([host_out[j].data_ptr() - pb * esz for j in range(nres)])
            self.tree.set_output_mirrors(mirrors, mask)
            self.exchange_mode = "in-kernel multicast stores" if use_mc else "in-kernel peer stores"
            try:
                if c1 > c0:
                    self.tree.acc_pot(Q, theta, G=G, eps=eps, out=out, where=RK_DEVICE, crit_range=(c0, c1))
                    info = self.tree.eval_info.asdict()
                else:
                    from . import EvalInfo
                    info = EvalInfo().asdict()
            finally:
                self.tree.set_output_mirrors([])
            main = torch.cuda.current_stream()
            self._stream_barrier(main)  # every rank's launch (and with it its stores) has completed
            return info, out
        self.exchange_mode = "copy-engine pushes of 4 launches"
        from . import device_copy_async
        nc = c1 - c0
        ccuts = sorted({c0, c1, *[c0 + int(nc * f) for f in chunks]})
        pcuts = [int(v) for v in self.tree.crit_begin_at(ccuts)]
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        if self._push is None:
            # one stream per peer: the copies to different peers run on different copy engines / NVLink ports
            self._push = [torch.cuda.Stream(device=self.dev) for _ in range(self.world)]
        side = self._side
        for s in self._push:
            s.wait_stream(main)
        esz = out[0].element_size()
        info = None
        if len(ccuts) < 2:  # an empty range (more ranks than critical nodes): nothing to evaluate, still exchange
            from . import EvalInfo
            info = EvalInfo().asdict()
        for k in range(len(ccuts) - 1):
            self.tree.acc_pot(Q, theta, G=G, eps=eps, out=out, where=RK_DEVICE, crit_range=(ccuts[k], ccuts[k + 1]))
            part = self.tree.eval_info.asdict()  # (the call returns after its launch has finished)
            if info is None:
                info = part
            else:
                for key in ("mac_tests", "accepted", "p2p_pairs", "self_pairs", "interactions", "n_groups",
                            "kernel_launches", "ms_kernel", "ms_total"):
                    info[key] += part[key]
            b, e = pcuts[k], pcuts[k + 1]
            if e > b and host_out is not None:
                with torch.cuda.stream(self._push[self.rank]):  # (this rank's own stream carries no peer pushes)
                    for j in range(nres):
                        host_out[j][b - pcuts[0]:e - pcuts[0]].copy_(out[j][b:e], non_blocking=True)
            if e > b:
                for d in range(1, self.world):
                    r = (self.rank + d) % self.world  # staggered: at any time every rank receives from one peer
                    for j in range(nres):
                        device_copy_async(ptrs[j][r] + b * esz, ptrs[j][self.rank] + b * esz, (e - b) * esz,
                                          self._push[r].cuda_stream)
        # every rank's pushes are ordered before its part of the barrier, so after it all slices have arrived
        for s in self._push:
            side.wait_stream(s)
        self._stream_barrier(side)
        main.wait_stream(side)
        return info, out

    def rebalance(self, kernel_ms=None):
        """Cost-weighted cuts from the last evaluation's per-group interaction counts. With kernel_ms (this rank's
        traversal time of that evaluation) the counts of every rank's range are rescaled by its measured time per
        interaction, so that ranges whose groups run at lower lane utilisation get fewer of them."""
        torch = self.torch
        local = self.tree.group_costs().astype(np.float64)
        if kernel_ms is not None and self.cuts is not None:
            c0, c1 = self.cuts[self.rank], self.cuts[self.rank + 1]
            tot = local[c0:c1].sum()
            if tot > 0:
                local[c0:c1] *= kernel_ms * 1e6 / tot  # cost unit: nanoseconds of this rank
        t = torch.from_numpy(local).to(self.dev)
        self.dist.all_reduce(t)
        costs = t.cpu().numpy()
        self.cuts = sharding.cuts_by_cost(costs, self.world)
        self.cut_particles = self.tree.crit_begin_at(self.cuts).astype(np.int64)
        self.cut_pidx = [int(v) for v in self.cut_particles]  # survives the next rebuild (re-snapped in _ensure_cuts)
        self._cuts_build_id = self._build_id
        return sharding.imbalance(costs, self.cuts)

    def check_against_single_gpu(self, shard, first_index, theta, outs, G=1.0, eps=0.0):
        """Hardware parity of the sharded path: every rank gathers the whole input, builds the single-GPU tree and
        evaluates ALL critical nodes with it, then compares (1) the fingerprints of the two trees' device arrays
        (codes, permutation, particles, node topology and properties, critical nodes: rk_tree_digest) and (2) the
        accelerations the sharded evaluation left on this rank, bit for bit. shard: this rank's x, y, z, m of the
        last build; outs: the outputs of the last acc_pot (Morton order, complete on every rank)."""
        torch, dist = self.torch, self.dist
        n_loc = shard[0].numel()
        sizes = torch.empty(self.world, dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(sizes, torch.tensor([n_loc], dtype=torch.int64, device=self.dev))
        sizes = [int(v) for v in sizes.tolist()]
        mx = max(sizes)
        full = []
        for t in shard:
            pad = torch.zeros(mx, dtype=t.dtype, device=self.dev)
            pad[:n_loc] = t
            allb = torch.empty(mx * self.world, dtype=t.dtype, device=self.dev)
            dist.all_gather_into_tensor(allb, pad)
            full.append(torch.cat([allb[r * mx:r * mx + sizes[r]] for r in range(self.world)]))
            del pad, allb
        n = full[0].numel()
        single = Octree(fp=self.fp, mac="bh" if self.tree.mac == 0 else "bh_geom", device=self.dev.index)
        single.set_stream(torch.cuda.current_stream().cuda_stream)
        single.build(full[0], full[1], full[2], full[3], max_leaf_n=self.mln, ncrit=self.ncrit, where=RK_DEVICE, n=n)
        d1, d2 = single.digest(), self.tree.digest()
        names = ("codes", "perm", "particles", "node_topology", "node_properties", "crit_nodes", "crit_begins", "sizes")
        differing = [nm for nm, a, b in zip(names, d1, d2) if int(a) != int(b)]
        ref = [torch.empty(n, dtype=self.dt, device=self.dev) for _ in range(len(outs))]
        Q = {3: 0, 1: 1, 4: 2}[len(outs)]
        single.acc_pot(Q, theta, G=G, eps=eps, out=ref, where=RK_DEVICE)
        acc_equal = all(bool(torch.equal(a[:n], b)) for a, b in zip(outs, ref))
        max_diff = max(float((a[:n] - b).abs().max().item()) for a, b in zip(outs, ref))
        # ... and in the ORIGINAL particle order (tree.hpp:3320-3330): the single-GPU tree's own ordered evaluation
        # against this rank's re-ordered copy of the sharded result
        ref_o = [torch.empty(n, dtype=self.dt, device=self.dev) for _ in range(len(outs))]
        single.acc_pot(Q, theta, G=G, eps=eps, out=ref_o, where=RK_DEVICE, ordered=True)
        mine_o = self.tree.to_original_order([a[:n] for a in outs], [torch.empty_like(r) for r in ref_o])
        ordered_equal = all(bool(torch.equal(a, b)) for a, b in zip(mine_o, ref_o))
        acc_equal = acc_equal and ordered_equal
        del ref_o, mine_o
        ok = torch.tensor([1 if (not differing and acc_equal) else 0], dtype=torch.int64, device=self.dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        single.close()
        return {"ok": bool(ok.item()), "nparts": n, "tree_arrays_differing": differing,
                "accelerations_bit_equal": acc_equal, "original_order_bit_equal": ordered_equal, "max_abs_diff": max_diff,
                "what": "sharded tree fingerprints and gathered accelerations vs a single-GPU build + full evaluation "
                        "of the same particles on every rank"}
