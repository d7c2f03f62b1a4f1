"""rakau_b200 — thin ctypes plumbing over librakau_b200.so (C ABI in include/rakau_b200.h).

The product is the CUDA library plus the C++17 header include/rakau/tree.hpp (the reference's own
`rakau::octree<F, MAC>` API). This module only lets Python drive the C ABI for tests, bench.py and
multi-GPU orchestration with torch.distributed. There is NO CPU fallback: if the shared library is
missing, or no CUDA device is present, construction fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RK_LIB") or os.path.join(_HERE, "lib", "librakau_b200.so")  # RK_LIB: tuning builds
_LIB = None

RK_HOST, RK_DEVICE = 0, 1
RK_PERM, RK_LAST_PERM, RK_INV_PERM = 0, 1, 2
FDT = {32: np.float32, 64: np.float64}

NODE_DTYPE = {
    32: np.dtype([("begin", "<u8"), ("end", "<u8"), ("n_children", "<u8"), ("code", "<u8"), ("level", "<u8"),
                  ("props", "<f4", (4,)), ("dim", "<f4"), ("delta", "<f4")]),
    64: np.dtype([("begin", "<u8"), ("end", "<u8"), ("n_children", "<u8"), ("code", "<u8"), ("level", "<u8"),
                  ("props", "<f8", (4,)), ("dim", "<f8"), ("delta", "<f8")]),
}

# Every symbol declared in include/rakau_b200.h (checked by tests/test_capi_symbols.py).
SYMBOLS = [
    "rk_device_count", "rk_min_size", "rk_tree_create", "rk_tree_destroy", "rk_last_error", "rk_create_error",
    "rk_tree_set_stream", "rk_tree_synchronize", "rk_tree_build", "rk_tree_update_positions",
    "rk_tree_update_masses", "rk_tree_clear", "rk_tree_nparts", "rk_tree_nnodes", "rk_tree_ncrit",
    "rk_tree_box_size", "rk_tree_get_parts", "rk_tree_get_codes", "rk_tree_get_perm", "rk_tree_get_nodes",
    "rk_tree_get_crit", "rk_tree_acc_pot", "rk_tree_acc_pot_range", "rk_tree_get_group_costs", "rk_tree_exact",
    "rk_traverse_external_tree", "rk_tree_group_costs_device", "rk_kernel_launch_count", "rk_measure_fp32_peak", "rk_device_copy_async", "rk_device_bcast_copy",
    "rk_plummer", "rk_tree_clone", "rk_plummer_leapfrog", "rk_tree_get_parts_device", "rk_tree_get_perm_device",
    "rk_tree_sort_shard", "rk_tree_get_codes_device", "rk_tree_build_presorted", "rk_deduce_box", "rk_tree_crit_begin_at",
    "rk_tree_crit_lower_bound", "rk_tree_digest", "rk_tree_last_kernel", "rk_measure_fp64_peak", "rk_tree_set_option", "rk_tree_set_output_mirrors",
    "rk_tree_leapfrog_init", "rk_tree_leapfrog_step", "rk_tree_leapfrog_get", "rk_tree_encode_shard",
    "rk_tree_partition_shard", "rk_tree_to_original_order",
]


class BuildInfo(C.Structure):
    _fields_ = [("box_size", C.c_double), ("n_nodes", C.c_uint64), ("n_crit", C.c_uint64),
                ("max_group", C.c_uint64), ("sort_passes", C.c_uint32), ("ms_total", C.c_float),
                ("ms_encode", C.c_float), ("ms_sort", C.c_float), ("ms_permute", C.c_float),
                ("ms_topology", C.c_float), ("ms_props", C.c_float)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class EvalInfo(C.Structure):
    _fields_ = [("mac_tests", C.c_uint64), ("accepted", C.c_uint64), ("p2p_pairs", C.c_uint64),
                ("self_pairs", C.c_uint64), ("interactions", C.c_uint64), ("n_groups", C.c_uint64),
                ("kernel_launches", C.c_uint32), ("ms_kernel", C.c_float), ("ms_total", C.c_float)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class LeapfrogInfo(C.Structure):
    _fields_ = [("ms_step", C.c_float), ("ms_integrals", C.c_float), ("ms_kick_drift", C.c_float),
                ("ms_rebuild", C.c_float), ("ms_traverse", C.c_float), ("ms_reindex", C.c_float),
                ("interactions", C.c_uint64), ("n_nodes", C.c_uint64), ("com", C.c_double * 3),
                ("com_v", C.c_double * 3), ("energy", C.c_double)]

    def asdict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["com"], d["com_v"] = list(self.com), list(self.com_v)
        return d


class RakauError(Exception):
    """status: 1 invalid_argument, 2 domain_error, 3 overflow_error, 4 runtime_error, 5 bad_alloc."""

    def __init__(self, status, msg):
        super().__init__(msg)
        self.status = status


def lib():
    """Load librakau_b200.so. Raises if it has not been built — there is no fallback path."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `make` or __graft_entry__.build(); "
                           "rakau_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, sz, dbl, i32 = C.c_void_p, C.c_size_t, C.c_double, C.c_int
    L.rk_device_count.restype = C.c_uint
    L.rk_min_size.restype = C.c_uint
    L.rk_tree_create.restype = vp
    L.rk_tree_create.argtypes = [i32, i32, i32]
    L.rk_tree_destroy.argtypes = [vp]
    L.rk_last_error.restype = C.c_char_p
    L.rk_last_error.argtypes = [vp]
    L.rk_create_error.restype = C.c_char_p
    L.rk_tree_set_stream.argtypes = [vp, vp]
    L.rk_tree_synchronize.argtypes = [vp]
    L.rk_tree_build.argtypes = [vp, vp, vp, vp, vp, sz, i32, dbl, i32, sz, sz, C.POINTER(BuildInfo)]
    L.rk_tree_update_positions.argtypes = [vp, vp, vp, vp, vp, i32, C.POINTER(BuildInfo)]
    L.rk_tree_clone.argtypes = [vp, vp]
    L.rk_tree_update_masses.argtypes = [vp, vp, i32]
    L.rk_tree_clear.argtypes = [vp]
    for f in ("rk_tree_nparts", "rk_tree_nnodes", "rk_tree_ncrit"):
        getattr(L, f).restype = sz
        getattr(L, f).argtypes = [vp]
    L.rk_tree_box_size.restype = dbl
    L.rk_tree_box_size.argtypes = [vp]
    L.rk_tree_get_parts.argtypes = [vp, vp, vp, vp, vp]
    L.rk_tree_get_codes.argtypes = [vp, vp]
    L.rk_tree_get_perm.argtypes = [vp, i32, vp]
    L.rk_tree_get_nodes.argtypes = [vp, vp]
    L.rk_tree_get_crit.argtypes = [vp, vp]
    L.rk_tree_get_group_costs.argtypes = [vp, vp]
    L.rk_tree_acc_pot.argtypes = [vp, i32, i32, dbl, dbl, dbl, vp, sz, C.POINTER(vp), i32, C.POINTER(EvalInfo)]
    L.rk_tree_acc_pot_range.argtypes = [vp, i32, i32, dbl, dbl, dbl, sz, sz, C.POINTER(vp), i32,
                                        C.POINTER(EvalInfo)]
    L.rk_tree_exact.argtypes = [vp, sz, i32, dbl, dbl, vp]
    L.rk_traverse_external_tree.argtypes = [i32, i32, i32, C.POINTER(vp), vp, sz, vp, sz, C.POINTER(vp), vp, sz,
                                            dbl, dbl, dbl, i32, sz, C.POINTER(EvalInfo), C.c_char_p, sz]
    L.rk_tree_group_costs_device.restype = vp
    L.rk_tree_group_costs_device.argtypes = [vp]
    L.rk_kernel_launch_count.restype = C.c_ulonglong
    L.rk_device_copy_async.argtypes = [vp, vp, sz, vp]
    L.rk_device_bcast_copy.argtypes = [C.POINTER(C.c_void_p), C.c_uint, vp, sz, vp, C.c_int]
    L.rk_measure_fp32_peak.argtypes = [i32, C.POINTER(dbl), C.POINTER(dbl)]
    L.rk_plummer.argtypes = [i32, sz, sz, sz, dbl, dbl, i32, sz, i32, vp, vp, vp, vp]
    L.rk_plummer_leapfrog.argtypes = [i32, sz, dbl, vp, vp, vp, vp, vp, vp, C.POINTER(sz)]
    L.rk_tree_get_parts_device.argtypes = [vp, vp, vp, vp, vp]
    L.rk_tree_get_perm_device.argtypes = [vp, i32, vp]
    L.rk_tree_sort_shard.argtypes = [vp, vp, vp, vp, vp, vp, sz, dbl]
    L.rk_tree_get_codes_device.argtypes = [vp, vp]
    L.rk_tree_encode_shard.argtypes = [vp, vp, vp, vp, vp, sz, dbl]
    L.rk_tree_partition_shard.argtypes = [vp, vp, C.c_uint, vp]
    L.rk_tree_build_presorted.argtypes = [vp, vp, vp, vp, vp, vp, vp, sz, dbl, sz, sz, vp, C.POINTER(BuildInfo)]
    L.rk_tree_crit_begin_at.argtypes = [vp, vp, sz, vp]
    L.rk_tree_crit_lower_bound.argtypes = [vp, vp, sz, vp]
    L.rk_tree_digest.argtypes = [vp, vp]
    L.rk_tree_to_original_order.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp)]
    L.rk_tree_set_option.argtypes = [vp, C.c_char_p, C.c_longlong]
    L.rk_tree_set_output_mirrors.argtypes = [vp, C.c_uint, C.POINTER(C.c_void_p), C.c_uint]
    L.rk_tree_leapfrog_init.argtypes = [vp, vp, vp, vp, i32, dbl, dbl, dbl, i32]
    L.rk_tree_leapfrog_step.argtypes = [vp, dbl, C.POINTER(LeapfrogInfo)]
    L.rk_tree_leapfrog_get.argtypes = [vp, i32, vp, vp, vp, i32]
    L.rk_tree_last_kernel.restype = C.c_char_p
    L.rk_tree_last_kernel.argtypes = [vp]
    L.rk_measure_fp64_peak.argtypes = [i32, C.POINTER(dbl), C.POINTER(dbl)]
    L.rk_deduce_box.restype = dbl
    L.rk_deduce_box.argtypes = [i32, dbl]
    _LIB = L
    return L


def _ptr(a):
    """Host numpy array, integer device pointer, or None -> c_void_p."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):  # torch tensor
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


class Octree:
    """Device-resident octree: mirrors the calls rakau::octree<F, MAC> makes into the C ABI."""

    def __init__(self, fp=32, mac="bh", device=0):
        self.fp = fp
        self.F = FDT[fp]
        self.L = lib()
        self.mac = 0 if mac == "bh" else 1
        self.device = device
        self.h = self.L.rk_tree_create(fp, self.mac, device)
        if not self.h:
            raise RuntimeError(self.L.rk_create_error().decode())
        self.build_info = BuildInfo()
        self.eval_info = EvalInfo()

    def close(self):
        if getattr(self, "h", None):
            self.L.rk_tree_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise RakauError(rc, self.L.rk_last_error(self.h).decode())

    def _prep(self, a, where):
        if where == RK_HOST and a is not None:
            return np.ascontiguousarray(a, dtype=self.F)
        return a

    def set_stream(self, stream_ptr):
        """cudaStream_t handle as an integer; 0 = legacy default stream (torch's default); -1 = the tree's own."""
        self._check(self.L.rk_tree_set_stream(self.h, C.c_void_p(stream_ptr if stream_ptr >= 0 else 2 ** 64 - 1)))

    def synchronize(self):
        self._check(self.L.rk_tree_synchronize(self.h))

    def set_output_mirrors(self, mirrors, multicast_mask=0):
        """mirrors: list (<= 8) of lists of integer device pointers / tensors, one per result array: further copies of
        the outputs that the following device-output evaluations write from inside the kernel ([] = off).
        multicast_mask: bit r set = mirror r holds NVSwitch multicast addresses."""
        flat = (C.c_void_p * (4 * max(len(mirrors), 1)))()
        for r, arrs in enumerate(mirrors):
            for j, a in enumerate(arrs):
                flat[4 * r + j] = _ptr(a)
        self._check(self.L.rk_tree_set_output_mirrors(self.h, len(mirrors), flat, int(multicast_mask)))

    def set_option(self, name, value):
        self._check(self.L.rk_tree_set_option(self.h, name.encode(), int(value)))

    def build(self, x, y, z, m, box_size=0.0, max_leaf_n=16, ncrit=128, where=RK_HOST, n=None):
        arrs = [self._prep(a, where) for a in (x, y, z, m)]
        if n is None:
            n = arrs[0].size if where == RK_HOST else arrs[0].numel()
        deduce = 1 if not box_size else 0
        self._keep = arrs
        self._check(self.L.rk_tree_build(self.h, *[_ptr(a) for a in arrs], n, where, float(box_size or 0.0), deduce,
                                         max_leaf_n, ncrit, C.byref(self.build_info)))
        return self.build_info

    def clone(self):
        other = Octree(fp=self.fp, mac="bh" if self.mac == 0 else "bh_geom", device=self.device)
        other._check(self.L.rk_tree_clone(other.h, self.h))
        return other

    def update_positions(self, x=None, y=None, z=None, m=None, where=RK_HOST):
        arrs = [self._prep(a, where) for a in (x, y, z, m)]
        self._check(self.L.rk_tree_update_positions(self.h, *[_ptr(a) for a in arrs], where,
                                                    C.byref(self.build_info)))
        return self.build_info

    def update_masses(self, m, where=RK_HOST):
        m = self._prep(m, where)
        self._check(self.L.rk_tree_update_masses(self.h, _ptr(m), where))

    def clear(self):
        self._check(self.L.rk_tree_clear(self.h))

    @property
    def nparts(self):
        return self.L.rk_tree_nparts(self.h)

    @property
    def nnodes(self):
        return self.L.rk_tree_nnodes(self.h)

    @property
    def ncrit_nodes(self):
        return self.L.rk_tree_ncrit(self.h)

    @property
    def box_size(self):
        return self.L.rk_tree_box_size(self.h)

    def parts(self):
        out = [np.empty(self.nparts, dtype=self.F) for _ in range(4)]
        self._check(self.L.rk_tree_get_parts(self.h, *[_ptr(a) for a in out]))
        return out

    def parts_device(self, x, y, z, m):
        """Morton-ordered SoA into device buffers (torch tensors / device pointers; None = skip)."""
        self._check(self.L.rk_tree_get_parts_device(self.h, _ptr(x), _ptr(y), _ptr(z), _ptr(m)))

    def perm_device(self, out, which=RK_PERM):
        """perm / last_perm / inv_perm as uint32 into a device buffer of nparts elements."""
        self._check(self.L.rk_tree_get_perm_device(self.h, which, _ptr(out)))

    def sort_shard(self, x, y, z, m, n, box_size, codes=None):
        """Sort a device-resident shard with the global box (multi-GPU sample sort building block)."""
        self._check(self.L.rk_tree_sort_shard(self.h, _ptr(x), _ptr(y), _ptr(z), _ptr(m), _ptr(codes), n,
                                              float(box_size)))

    def encode_shard(self, x, y, z, m, n, box_size):
        """Pack + Morton-encode a device-resident shard with the global box (codes in input order: codes_device)."""
        self._check(self.L.rk_tree_encode_shard(self.h, _ptr(x), _ptr(y), _ptr(z), _ptr(m), n, float(box_size)))

    def partition_shard(self, splitters):
        """Group the encoded shard by splitter bucket (one stable radix pass); returns the bucket sizes."""
        nsplit = splitters.numel() if hasattr(splitters, "numel") else len(splitters)
        counts = np.zeros(nsplit + 1, dtype=np.uint64)
        self._check(self.L.rk_tree_partition_shard(self.h, _ptr(splitters), nsplit, _ptr(counts)))
        return counts

    def codes_device(self, out):
        self._check(self.L.rk_tree_get_codes_device(self.h, _ptr(out)))

    def build_presorted(self, x, y, z, m, codes, perm, n, box_size, max_leaf_n=16, ncrit=128, parts_ready_event=None):
        """Tree (topology + node properties) from globally sorted device arrays. parts_ready_event: cudaEvent_t
        handle (int) after which x, y, z, m, perm are valid; the topology is built from the codes before it."""
        self._check(self.L.rk_tree_build_presorted(self.h, _ptr(x), _ptr(y), _ptr(z), _ptr(m), _ptr(codes), _ptr(perm),
                                                   n, float(box_size), max_leaf_n, ncrit,
                                                   C.c_void_p(parts_ready_event) if parts_ready_event else None,
                                                   C.byref(self.build_info)))
        return self.build_info

    def codes(self):
        out = np.empty(self.nparts, dtype=np.uint64)
        self._check(self.L.rk_tree_get_codes(self.h, _ptr(out)))
        return out

    def perm(self, which=RK_PERM):
        out = np.empty(self.nparts, dtype=np.uint64)
        self._check(self.L.rk_tree_get_perm(self.h, which, _ptr(out)))
        return out

    def nodes(self):
        out = np.empty(self.nnodes, dtype=NODE_DTYPE[self.fp])
        self._check(self.L.rk_tree_get_nodes(self.h, _ptr(out)))
        return out

    def crit(self):
        out = np.empty((self.ncrit_nodes, 3), dtype=np.uint64)
        self._check(self.L.rk_tree_get_crit(self.h, _ptr(out)))
        return out

    def crit_begin_at(self, idx):
        """First particle of the given critical nodes (index ncrit_nodes -> nparts)."""
        idx = np.ascontiguousarray(idx, dtype=np.uint64)
        out = np.empty(idx.size, dtype=np.uint64)
        self._check(self.L.rk_tree_crit_begin_at(self.h, _ptr(idx), idx.size, _ptr(out)))
        return out

    def crit_lower_bound(self, particle_idx):
        """First critical node whose first particle is >= each given particle index (<= 16 values)."""
        idx = np.ascontiguousarray(particle_idx, dtype=np.uint64)
        out = np.empty(idx.size, dtype=np.uint64)
        self._check(self.L.rk_tree_crit_lower_bound(self.h, _ptr(idx), idx.size, _ptr(out)))
        return out

    def to_original_order(self, arrays, out):
        """out[j][perm[i]] = arrays[j][i] for device arrays (torch tensors / pointers) in the tree's Morton order."""
        n = len(arrays)
        a = (C.c_void_p * 4)(*([_ptr(t) for t in arrays] + [None] * (4 - n)))
        o = (C.c_void_p * 4)(*([_ptr(t) for t in out] + [None] * (4 - n)))
        self._check(self.L.rk_tree_to_original_order(self.h, n, a, o))
        return out

    def digest(self):
        """Fingerprints of the device arrays (codes, perm, particles, nodeB, nodeA, crit nodes, crit begins, sizes)."""
        out = np.zeros(8, dtype=np.uint64)
        self._check(self.L.rk_tree_digest(self.h, _ptr(out)))
        return out

    def last_kernel(self):
        return self.L.rk_tree_last_kernel(self.h).decode()

    def group_costs_device_ptr(self):
        return self.L.rk_tree_group_costs_device(self.h)

    def group_costs(self):
        out = np.empty(self.ncrit_nodes, dtype=np.uint64)
        self._check(self.L.rk_tree_get_group_costs(self.h, _ptr(out)))
        return out

    def acc_pot(self, Q, theta, G=1.0, eps=0.0, ordered=False, split=None, out=None, where=RK_HOST,
                crit_range=None):
        """Q: 0 accs, 1 pots, 2 accs+pots. Returns the list of output arrays (host numpy unless `out` given)."""
        nres = {0: 3, 1: 1, 2: 4}[Q]
        if out is None:
            assert where == RK_HOST
            out = [np.zeros(self.nparts, dtype=self.F) for _ in range(nres)]
        ptrs = (C.c_void_p * 4)(*([_ptr(a) for a in out] + [None] * (4 - nres)))
        if crit_range is not None:
            rc = self.L.rk_tree_acc_pot_range(self.h, Q, int(ordered), float(theta), float(G), float(eps),
                                              crit_range[0], crit_range[1], ptrs, where, C.byref(self.eval_info))
        else:
            sp = None if split is None else np.ascontiguousarray(split, dtype=np.float64)
            rc = self.L.rk_tree_acc_pot(self.h, Q, int(ordered), float(theta), float(G), float(eps), _ptr(sp),
                                        0 if sp is None else sp.size, ptrs, where, C.byref(self.eval_info))
        self._check(rc)
        return out

    # ---- device-resident leapfrog (benchmark_leapfrog.cpp:252-384) ----
    def leapfrog_init(self, vx, vy, vz, theta, G=1.0, eps=0.0, track_integrals=False, where=RK_HOST):
        arrs = [self._prep(a, where) for a in (vx, vy, vz)]
        self._check(self.L.rk_tree_leapfrog_init(self.h, *[_ptr(a) for a in arrs], where, float(theta), float(G),
                                                 float(eps), int(track_integrals)))

    def leapfrog_step(self, dt):
        info = LeapfrogInfo()
        self._check(self.L.rk_tree_leapfrog_step(self.h, float(dt), C.byref(info)))
        return info

    def leapfrog_get(self, what):
        """what: 0 velocities, 1 accelerations, 2 kicked velocities (order before the last rebuild), 3 potentials."""
        out = [np.empty(self.nparts, dtype=self.F) for _ in range(1 if what == 3 else 3)]
        ptrs = [_ptr(a) for a in out] + [None] * (3 - len(out))
        self._check(self.L.rk_tree_leapfrog_get(self.h, what, *ptrs, RK_HOST))
        return out

    def exact(self, idx, G=1.0, eps=0.0, ordered=False):
        out = np.zeros(4, dtype=np.float64)
        self._check(self.L.rk_tree_exact(self.h, idx, int(ordered), float(G), float(eps), _ptr(out)))
        return out


def device_count():
    return lib().rk_device_count()


def deduce_box(absmax, fp=32):
    """Box size the library deduces for a given max |coordinate| (tree.hpp:1309-1312)."""
    return lib().rk_deduce_box(fp, float(absmax))


def device_copy_async(dst_ptr, src_ptr, nbytes, stream_ptr):
    """Copy-engine D2D copy (dst may be mapped peer memory) on the given cudaStream_t handle."""
    rc = lib().rk_device_copy_async(C.c_void_p(dst_ptr), C.c_void_p(src_ptr), nbytes, C.c_void_p(stream_ptr))
    if rc:
        raise RakauError(rc, "rk_device_copy_async failed")


def device_bcast_copy(dst_ptrs, src_ptr, nbytes, stream_ptr, multicast=False):
    """One kernel on the given cudaStream_t that copies nbytes from src to each of the (<= 8) destinations; multicast:
    the single destination is an NVSwitch multicast address."""
    arr = (C.c_void_p * len(dst_ptrs))(*[C.c_void_p(int(p)) for p in dst_ptrs])
    rc = lib().rk_device_bcast_copy(arr, len(dst_ptrs), C.c_void_p(src_ptr), nbytes, C.c_void_p(stream_ptr),
                                    1 if multicast else 0)
    if rc:
        raise RakauError(rc, "rk_device_bcast_copy failed")


def kernel_launch_count():
    return lib().rk_kernel_launch_count()


def measure_fp32_peak(device=0):
    """Measured FP32-pipe peak (TFLOP/s, FFMA = 2 flop) on `device`."""
    tf, ms = C.c_double(), C.c_double()
    rc = lib().rk_measure_fp32_peak(device, C.byref(tf), C.byref(ms))
    if rc:
        raise RuntimeError("rk_measure_fp32_peak failed")
    return tf.value


def measure_fp64_peak(device=0):
    """Measured FP64-pipe peak (TFLOP/s, DFMA = 2 flop) on `device`."""
    tf, ms = C.c_double(), C.c_double()
    rc = lib().rk_measure_fp64_peak(device, C.byref(tf), C.byref(ms))
    if rc:
        raise RuntimeError("rk_measure_fp64_peak failed")
    return tf.value


def plummer(n_total, first=0, count=None, a=1.0, size=0.0, fp=32, chunk=0, nthreads=None, out=None):
    """Plummer sphere of the reference's benchmarks (benchmark/common.hpp:39-126). chunk == 0: the sequential
    branch; chunk > 0: deterministic chunked form, shard [first, first+count). Returns m, x, y, z (numpy, or
    the given `out` arrays/tensors)."""
    count = n_total if count is None else count
    if out is None:
        out = [np.empty(count, dtype=FDT[fp]) for _ in range(4)]
    nthreads = nthreads or os.cpu_count() or 1
    rc = lib().rk_plummer(fp, n_total, first, count, a, size, 1 if chunk else 0, chunk, nthreads,
                          *[_ptr(o) for o in out])
    if rc:
        raise ValueError("rk_plummer: invalid arguments")
    return out


def traverse_external_tree(nodes, parts, codes, Q, mac_value, G=1.0, eps2=0.0, mac="bh", fp=32, first=0, ncrit=128,
                           offset_output=True):
    """rk_traverse_external_tree: the literal replacement of the reference's cuda_acc_pot_impl
    (src/rakau_cuda.cu:348-528). `nodes`: DFS AoS (NODE_DTYPE[fp]); `parts`: x, y, z, m in Morton order.
    Returns (outputs, EvalInfo dict)."""
    L = lib()
    F = FDT[fp]
    n = parts[0].size
    nres = {0: 3, 1: 1, 2: 4}[Q]
    out = [np.zeros(n if offset_output else n - first, dtype=F) for _ in range(nres)]
    optr = (C.c_void_p * 4)(*([_ptr(a) for a in out] + [None] * (4 - nres)))
    pa = [np.ascontiguousarray(a, dtype=F) for a in parts]
    pptr = (C.c_void_p * 4)(*[_ptr(a) for a in pa])
    nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE[fp])
    split = np.array([first, n], dtype=np.uint64)
    codes = np.ascontiguousarray(codes, dtype=np.uint64)
    info = EvalInfo()
    err = C.create_string_buffer(512)
    rc = L.rk_traverse_external_tree(fp, 0 if mac == "bh" else 1, Q, optr, _ptr(split), 2, _ptr(nodes), nodes.size,
                                     pptr, _ptr(codes), n, float(mac_value), float(G), float(eps2),
                                     1 if offset_output else 0, ncrit, C.byref(info), err, 512)
    if rc:
        raise RakauError(rc, err.value.decode())
    return out, info.asdict()


def plummer_leapfrog(n, a=1.0, fp=32):
    """Initial conditions of the reference's benchmark_leapfrog (positions + velocities, clipped at 10a).
    Returns x, y, z, vx, vy, vz (numpy, length = particles kept)."""
    out = [np.empty(n, dtype=FDT[fp]) for _ in range(6)]
    kept = C.c_size_t(0)
    rc = lib().rk_plummer_leapfrog(fp, n, a, *[_ptr(o) for o in out], C.byref(kept))
    if rc:
        raise ValueError("rk_plummer_leapfrog: invalid arguments")
    return [o[:kept.value] for o in out]


Leapfrog = True  # the device-resident integrator is part of this build (bench.py checks for it)


def leapfrog_benchmark(n, steps, theta=0.75, max_leaf_n=16, ncrit=128, device=0, dt=1e-4, e2e_steps=3):
    """BASELINE config 4: the reference's benchmark_leapfrog (Plummer sphere with velocities clipped at 10 core radii,
    equal masses, eps = 0.45 N^-0.73, kick-drift-kick with a tree rebuild every step).
    device_resident: rk_tree_leapfrog_step (nothing crosses PCIe). e2e_host_functors: the structure of the reference's
    own loop - positions read back (p_its_u), kicked/drifted by host code, rk_tree_update_positions from host buffers,
    accelerations written to host buffers, velocities updated on the host through last_perm."""
    import statistics
    import time
    t0 = time.time()
    x, y, z, vx, vy, vz = plummer_leapfrog(n)
    gen_s = time.time() - t0
    kept = x.size
    m = np.full(kept, np.float32(1) / np.float32(kept), dtype=np.float32)
    eps = float(np.float32(0.45) * np.float32(kept) ** np.float32(-0.73))
    t = Octree(fp=32, mac="bh", device=device)
    t.build(x, y, z, m, max_leaf_n=max_leaf_n, ncrit=ncrit)
    t.leapfrog_init(vx, vy, vz, theta, eps=eps)
    infos = [t.leapfrog_step(dt).asdict() for _ in range(steps + 2)][2:]
    med = lambda k: float(statistics.median(i[k] for i in infos))  # noqa: E731
    res = {"nparts_requested": n, "nparts": int(kept), "steps": steps, "dt": dt, "eps": eps, "generator_s": gen_s,
           "device_resident": {"ms_per_step": med("ms_step"), "ms_kick_drift": med("ms_kick_drift"),
                               "ms_rebuild": med("ms_rebuild"), "ms_traverse": med("ms_traverse"),
                               "ms_reindex": med("ms_reindex"), "interactions_per_step": infos[-1]["interactions"],
                               "kernel": t.last_kernel()}}
    # ---- the reference's structure: host functors around update_particles_u ----
    acc = [np.empty(kept, dtype=np.float32) for _ in range(3)]
    v = [a.copy() for a in t.leapfrog_get(0)]
    t.acc_pot(0, theta, eps=eps, out=acc)
    half = np.float32(dt / 2)
    ts = []
    for s in range(e2e_steps + 1):
        t0 = time.time()
        kv = [a * half + b for a, b in zip(acc, v)]
        pos = t.parts()
        t.update_positions(*[k * np.float32(dt) + p for k, p in zip(kv, pos[:3])])
        t.acc_pot(0, theta, eps=eps, out=acc)
        lp = t.perm(RK_LAST_PERM).astype(np.int64)
        v = [a * half + k[lp] for a, k in zip(acc, kv)]
        ts.append(time.time() - t0)
    res["e2e_host_functors"] = {"ms_per_step": 1e3 * float(statistics.median(ts[1:])), "steps": e2e_steps,
                                "h2d_bytes_per_step": 12 * int(kept), "d2h_bytes_per_step": (16 + 12 + 8) * int(kept),
                                "note": "host kick/drift/velocity update in numpy (single thread) between "
                                        "rk_tree_get_parts, rk_tree_update_positions and rk_tree_acc_pot with host buffers"}
    t.close()
    return res
