"""Parity oracle for rakau_b200 — TEST INFRASTRUCTURE ONLY.

ctypes bindings over oracle/liboracle.so (the scalar CPU restatement of the reference's
Barnes-Hut path, see rakau_oracle.cpp) and, when present, oracle/_ref/libref.so (the unmodified
reference header compiled against dependency stand-ins). Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this package; nothing under
rakau_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NODE_DTYPE = {
    32: np.dtype([("begin", "<u8"), ("end", "<u8"), ("n_children", "<u8"), ("code", "<u8"), ("level", "<u8"),
                  ("props", "<f4", (4,)), ("dim", "<f4"), ("delta", "<f4")]),
    64: np.dtype([("begin", "<u8"), ("end", "<u8"), ("n_children", "<u8"), ("code", "<u8"), ("level", "<u8"),
                  ("props", "<f8", (4,)), ("dim", "<f8"), ("delta", "<f8")]),
}
FDT = {32: np.float32, 64: np.float64}


def build(force=False):
    """Compile liboracle.so (and _ref/libref.so when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, "rakau_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        vp, u64p, sz, dbl, i32 = C.c_void_p, C.POINTER(C.c_uint64), C.c_size_t, C.c_double, C.c_int
        L.orc_create.restype = vp
        L.orc_create.argtypes = [i32, i32]
        L.orc_destroy.argtypes = [vp]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [vp]
        L.orc_build.restype = i32
        L.orc_build.argtypes = [vp, vp, vp, vp, vp, sz, dbl, i32, sz, sz]
        for f in ("orc_nparts", "orc_nnodes", "orc_ncrit_nodes", "orc_node_stride"):
            getattr(L, f).restype = sz
            getattr(L, f).argtypes = [vp]
        L.orc_box_size.restype = dbl
        L.orc_box_size.argtypes = [vp]
        L.orc_get_codes.argtypes = [vp, vp]
        L.orc_get_perm.argtypes = [vp, i32, vp]
        L.orc_get_parts.argtypes = [vp, vp, vp, vp, vp]
        L.orc_get_nodes.argtypes = [vp, vp]
        L.orc_get_crit.argtypes = [vp, vp, vp]
        L.orc_acc_pot.restype = i32
        L.orc_acc_pot.argtypes = [vp, i32, dbl, dbl, dbl, vp, vp, vp, vp, vp, vp, i32]
        L.orc_acc_pot_sample.restype = i32
        L.orc_acc_pot_sample.argtypes = [vp, i32, dbl, dbl, dbl, vp, vp, vp, vp, vp, i32, sz, sz]
        L.orc_exact.restype = i32
        L.orc_exact.argtypes = [vp, sz, dbl, dbl, vp]
        L.orc_update_positions.restype = i32
        L.orc_update_positions.argtypes = [vp, vp, vp, vp]
        L.orc_update_masses.restype = i32
        L.orc_update_masses.argtypes = [vp, vp]
        L.orc_node_centre.argtypes = [vp, C.c_uint64, vp]
        L.orc_morton_encode.restype = C.c_uint64
        L.orc_morton_encode.argtypes = [C.c_uint64] * 3
        L.orc_morton_decode.argtypes = [C.c_uint64, vp]
        L.orc_plummer.argtypes = [i32, sz, dbl, dbl, vp]
        L.orc_plummer_chunked.argtypes = [i32, sz, dbl, dbl, sz, i32, vp]
        L.orc_rng_create.restype = vp
        L.orc_rng_create.argtypes = [C.c_uint32]
        L.orc_rng_destroy.argtypes = [vp]
        L.orc_uniform.argtypes = [i32, sz, dbl, vp, vp]
        _LIB = L
    return _LIB


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code  # 1 invalid_argument, 2 domain_error, 3 overflow


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleTree:
    """Scalar CPU restatement of rakau::octree<F, MAC> (hot path only)."""

    def __init__(self, x, y, z, m, box_size=0.0, max_leaf_n=16, ncrit=128, mac="bh", fp=32):
        self.fp = fp
        self.F = FDT[fp]
        self.L = lib()
        self.h = self.L.orc_create(fp, 0 if mac == "bh" else 1)
        arrs = [np.ascontiguousarray(a, dtype=self.F) for a in (x, y, z, m)]
        deduce = 1 if (box_size is None or box_size == 0.0) else 0
        rc = self.L.orc_build(self.h, *[_p(a) for a in arrs], arrs[0].size, float(box_size or 0.0), deduce,
                              max_leaf_n, ncrit)
        self._check(rc)

    def _check(self, rc):
        if rc:
            raise OracleError(rc, self.L.orc_last_error(self.h).decode())

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    @property
    def nparts(self):
        return self.L.orc_nparts(self.h)

    @property
    def box_size(self):
        return self.L.orc_box_size(self.h)

    def codes(self):
        out = np.empty(self.nparts, dtype=np.uint64)
        self.L.orc_get_codes(self.h, _p(out))
        return out

    def perm(self, which=0):
        out = np.empty(self.nparts, dtype=np.uint64)
        self.L.orc_get_perm(self.h, which, _p(out))
        return out

    def parts(self):
        out = [np.empty(self.nparts, dtype=self.F) for _ in range(4)]
        self.L.orc_get_parts(self.h, *[_p(a) for a in out])
        return out

    def nodes(self):
        n = self.L.orc_nnodes(self.h)
        out = np.empty(n, dtype=NODE_DTYPE[self.fp])
        assert out.dtype.itemsize == self.L.orc_node_stride(self.h)
        self.L.orc_get_nodes(self.h, _p(out))
        return out

    def crit(self):
        n = self.L.orc_ncrit_nodes(self.h)
        out = np.empty((n, 3), dtype=np.uint64)
        idx = np.empty(n, dtype=np.uint64)
        self.L.orc_get_crit(self.h, _p(out), _p(idx))
        return out, idx

    def acc_pot(self, Q, theta, G=1.0, eps=0.0, nthreads=1, per_group=False):
        """Returns (list of output arrays in Morton order, counters dict[, per-group interactions])."""
        nres = {0: 3, 1: 1, 2: 4}[Q]
        out = [np.zeros(self.nparts, dtype=self.F) for _ in range(nres)]
        ptrs = [_p(a) for a in out] + [None] * (4 - nres)
        cnt = np.zeros(6, dtype=np.uint64)
        pg = np.zeros(self.L.orc_ncrit_nodes(self.h), dtype=np.uint64) if per_group else None
        rc = self.L.orc_acc_pot(self.h, Q, float(theta), float(G), float(eps), *ptrs, _p(cnt), _p(pg), nthreads)
        self._check(rc)
        names = ("mac_tests", "accepted", "p2p_pairs", "self_pairs", "interactions", "leaves_opened")
        c = {k: int(v) for k, v in zip(names, cnt)}
        return (out, c, pg) if per_group else (out, c)

    def acc_pot_sample(self, Q, theta, stride, offset=0, G=1.0, eps=0.0, nthreads=1):
        """Evaluate every stride-th critical node only (CPU-baseline timing). Returns the counters."""
        nres = {0: 3, 1: 1, 2: 4}[Q]
        out = [np.zeros(self.nparts, dtype=self.F) for _ in range(nres)]
        ptrs = [_p(a) for a in out] + [None] * (4 - nres)
        cnt = np.zeros(6, dtype=np.uint64)
        self._check(self.L.orc_acc_pot_sample(self.h, Q, float(theta), float(G), float(eps), *ptrs, _p(cnt), nthreads,
                                              stride, offset))
        names = ("mac_tests", "accepted", "p2p_pairs", "self_pairs", "interactions", "leaves_opened")
        return {k: int(v) for k, v in zip(names, cnt)}

    def exact(self, idx, G=1.0, eps=0.0):
        out = np.zeros(4, dtype=np.float64)
        self._check(self.L.orc_exact(self.h, idx, float(G), float(eps), _p(out)))
        return out

    def update_positions(self, x=None, y=None, z=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=self.F) for a in (x, y, z)]
        self._check(self.L.orc_update_positions(self.h, *[_p(a) for a in arrs]))

    def update_masses(self, m):
        m = np.ascontiguousarray(m, dtype=self.F)
        self._check(self.L.orc_update_masses(self.h, _p(m)))

    def node_centre(self, code):
        out = np.zeros(3, dtype=np.float64)
        self.L.orc_node_centre(self.h, int(code), _p(out))
        return out


def morton_encode(x, y, z):
    return lib().orc_morton_encode(int(x), int(y), int(z))


def morton_decode(code):
    out = np.zeros(3, dtype=np.uint64)
    lib().orc_morton_decode(int(code), _p(out))
    return [int(v) for v in out]


def plummer(n, a=1.0, size=0.0, fp=32, chunk=0, nthreads=8):
    """Plummer sphere of benchmark/common.hpp (sequential branch if chunk == 0). Returns m, x, y, z."""
    out = np.empty(4 * n, dtype=FDT[fp])
    if chunk:
        lib().orc_plummer_chunked(fp, n, a, size, chunk, nthreads, _p(out))
    else:
        lib().orc_plummer(fp, n, a, size, _p(out))
    return out[:n], out[n:2 * n], out[2 * n:3 * n], out[3 * n:]


class Rng:
    """std::mt19937 with the uniform-particle fixture of test/test_utils.hpp:41-59."""

    def __init__(self, seed):
        self.r = lib().orc_rng_create(seed)

    def __del__(self):
        try:
            lib().orc_rng_destroy(self.r)
        except Exception:
            pass

    def uniform_particles(self, n, size, fp=32):
        out = np.empty(4 * n, dtype=FDT[fp])
        lib().orc_uniform(fp, n, float(size), self.r, _p(out))
        return out[:n], out[n:2 * n], out[2 * n:3 * n], out[3 * n:]


# ---------------------------------------------------------------------------------------------------------
# oracle/_ref: the UNMODIFIED reference header compiled against dependency stand-ins (oracle/ref_shim)
# ---------------------------------------------------------------------------------------------------------
_REF = {}


def ref_available(variant="scalar"):
    return os.path.exists(os.path.join(_HERE, "_ref", f"libref_{variant}.so"))


def best_ref_variant():
    """Fastest variant the host CPU can run: avx512 > avx2 (both: reference SIMD path); None if not built."""
    flags = ""
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        pass
    if "avx512f" in flags and "avx512dq" in flags and ref_available("avx512"):
        return "avx512"
    if "avx2" in flags and ref_available("avx2"):
        return "avx2"
    return None


def ref_lib(variant="scalar"):
    if variant in _REF:
        return _REF[variant]
    L = C.CDLL(os.path.join(_HERE, "_ref", f"libref_{variant}.so"))
    vp, sz, dbl, i32 = C.c_void_p, C.c_size_t, C.c_double, C.c_int
    L.ref_variant.restype = C.c_char_p
    L.ref_set_threads.argtypes = [i32]
    L.ref_create.restype = vp
    L.ref_create.argtypes = [i32, i32, vp, vp, vp, vp, sz, dbl, i32, sz, sz, C.c_char_p, sz]
    L.ref_destroy.argtypes = [vp]
    L.ref_last_error.restype = C.c_char_p
    L.ref_last_error.argtypes = [vp]
    for f in ("ref_nparts", "ref_nnodes"):
        getattr(L, f).restype = sz
        getattr(L, f).argtypes = [vp]
    L.ref_box_size.restype = dbl
    L.ref_box_size.argtypes = [vp]
    L.ref_get_codes.argtypes = [vp, vp]
    L.ref_get_perm.argtypes = [vp, i32, vp]
    L.ref_get_parts.argtypes = [vp, vp, vp, vp, vp]
    L.ref_get_nodes.argtypes = [vp, vp]
    L.ref_acc_pot.restype = i32
    L.ref_acc_pot.argtypes = [vp, i32, dbl, dbl, dbl, vp, vp, vp, vp]
    L.ref_acc_pot_split.restype = i32
    L.ref_acc_pot_split.argtypes = [vp, i32, dbl, dbl, dbl, vp, vp, vp, vp, vp, sz]
    L.ref_exact.restype = i32
    L.ref_exact.argtypes = [vp, sz, dbl, dbl, vp]
    L.ref_update_positions.restype = i32
    L.ref_update_positions.argtypes = [vp, vp, vp, vp]
    L.ref_update_masses.restype = i32
    L.ref_update_masses.argtypes = [vp, vp]
    _REF[variant] = L
    return L


class RefTree:
    """rakau::octree<F, MAC> of the unmodified reference header (shimmed dependencies)."""

    def __init__(self, x, y, z, m, box_size=0.0, max_leaf_n=16, ncrit=128, mac="bh", fp=32, variant="scalar",
                 nthreads=None):
        self.fp, self.F = fp, FDT[fp]
        self.L = ref_lib(variant)
        self.L.ref_set_threads(nthreads or os.cpu_count() or 1)
        arrs = [np.ascontiguousarray(a, dtype=self.F) for a in (x, y, z, m)]
        err = C.create_string_buffer(1024)
        deduce = 1 if not box_size else 0
        self.h = self.L.ref_create(fp, 0 if mac == "bh" else 1, *[_p(a) for a in arrs], arrs[0].size,
                                   float(box_size or 0.0), deduce, max_leaf_n, ncrit, err, 1024)
        if not self.h:
            raise OracleError(1, err.value.decode())

    def __del__(self):
        try:
            if self.h:
                self.L.ref_destroy(self.h)
        except Exception:
            pass

    def variant(self):
        return self.L.ref_variant().decode()

    @property
    def nparts(self):
        return self.L.ref_nparts(self.h)

    @property
    def box_size(self):
        return self.L.ref_box_size(self.h)

    def codes(self):
        out = np.empty(self.nparts, dtype=np.uint64)
        self.L.ref_get_codes(self.h, _p(out))
        return out

    def perm(self, which=0):
        out = np.empty(self.nparts, dtype=np.uint64)
        self.L.ref_get_perm(self.h, which, _p(out))
        return out

    def parts(self):
        out = [np.empty(self.nparts, dtype=self.F) for _ in range(4)]
        self.L.ref_get_parts(self.h, *[_p(a) for a in out])
        return out

    def nodes(self):
        out = np.empty(self.L.ref_nnodes(self.h), dtype=NODE_DTYPE[self.fp])
        self.L.ref_get_nodes(self.h, _p(out))
        return out

    def acc_pot(self, Q, theta, G=1.0, eps=0.0, split=None):
        """split: the reference's `split` kwarg (tree.hpp:3147-3198; CPU share first, then one share per accelerator).
        Only the variants built with RAKAU_WITH_CUDA ("bridge": cuda_acc_pot_impl = integration/rakau_b200_bridge.cpp,
        "cuda": the reference's own src/rakau_cuda.cu) accept more than one entry."""
        nres = {0: 3, 1: 1, 2: 4}[Q]
        out = [np.zeros(self.nparts, dtype=self.F) for _ in range(nres)]
        ptrs = [_p(a) for a in out] + [None] * (4 - nres)
        if split is not None:
            sp = np.ascontiguousarray(split, dtype=np.float64)
            rc = self.L.ref_acc_pot_split(self.h, Q, float(theta), float(G), float(eps), *ptrs, _p(sp), sp.size)
        else:
            rc = self.L.ref_acc_pot(self.h, Q, float(theta), float(G), float(eps), *ptrs)
        if rc:
            raise OracleError(rc, self.L.ref_last_error(self.h).decode())
        return out

    def exact(self, idx, G=1.0, eps=0.0):
        out = np.zeros(4, dtype=np.float64)
        rc = self.L.ref_exact(self.h, idx, float(G), float(eps), _p(out))
        if rc:
            raise OracleError(rc, self.L.ref_last_error(self.h).decode())
        return out

    def update_positions(self, x=None, y=None, z=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=self.F) for a in (x, y, z)]
        rc = self.L.ref_update_positions(self.h, *[_p(a) for a in arrs])
        if rc:
            raise OracleError(rc, self.L.ref_last_error(self.h).decode())

    def update_masses(self, m):
        m = np.ascontiguousarray(m, dtype=self.F)
        rc = self.L.ref_update_masses(self.h, _p(m))
        if rc:
            raise OracleError(rc, self.L.ref_last_error(self.h).decode())
