"""numpy model of the non-recursive (per-particle) octree build used by the CUDA kernels
(rakau_b200/csrc/build.cu). It exists so the formulation can be checked against the oracle's
recursive build (tree.hpp:724-833 semantics) on the CPU. Not used by the product."""
import numpy as np

CBITS = 21


def prefix_digits(a, b):
    """number of leading 3-bit digits (out of 21) shared by 63-bit codes a and b (21 if equal)."""
    x = np.bitwise_xor(a, b).astype(np.uint64)
    out = np.full(x.shape, CBITS, dtype=np.int64)
    nz = x != 0
    # index of highest set bit
    hb = np.floor(np.log2(x[nz].astype(np.float64))).astype(np.int64)
    # fix float rounding
    hb = np.where((np.uint64(1) << hb.astype(np.uint64)) > x[nz], hb - 1, hb)
    hb = np.where((np.uint64(1) << (hb + 1).astype(np.uint64)) <= x[nz], hb + 1, hb)
    out[nz] = CBITS - 1 - hb // 3
    return out


def window_level(codes, w):
    """W_w(i): deepest level at which particle i's cell holds more than w particles (-1 if none)."""
    N = codes.size
    P = np.full(N, -1, dtype=np.int64)
    if N > w:
        P[: N - w] = prefix_digits(codes[: N - w], codes[w:])
    # sliding max over a in [i-w, i]
    W = np.full(N, -1, dtype=np.int64)
    # doubling
    A = P.copy()
    width = 1
    target = w + 1
    k = 1
    while k * 2 <= target:
        B = A.copy()
        B[: N - k] = np.maximum(A[: N - k], A[k:]) if N > k else B[: N - k]
        A = B
        k *= 2
    # A[a] = max P[a .. a+k), k = largest pow2 <= target. window [i-w, i]: combine A[i-w] and A[i-k+1]
    idx = np.arange(N)
    lo = idx - w
    hi = idx - k + 1
    v1 = np.where(lo >= 0, A[np.clip(lo, 0, N - 1)], -1)
    # for lo<0 the window is [0, i]; cover with A[0] (if k <= i+1) and A[i-k+1]
    v1 = np.where(lo < 0, np.where(idx + 1 >= k, A[0], -1), v1)
    v2 = np.where(hi >= 0, A[np.clip(hi, 0, N - 1)], -1)
    # if the window [max(0,i-w), i] is shorter than k, fall back to direct evaluation
    short = (idx + 1) < k
    W = np.maximum(v1, v2)
    for i in np.nonzero(short)[0]:
        W[i] = P[: i + 1].max()
    return W


def build(codes, max_leaf_n, ncrit):
    """Returns dict with DFS-ordered arrays begin,end,n_children,code,level and the crit list."""
    codes = np.asarray(codes, dtype=np.uint64)
    N = codes.size
    delta = np.full(N, -1, dtype=np.int64)
    if N > 1:
        delta[1:] = prefix_digits(codes[:-1], codes[1:])
    Wm = window_level(codes, max_leaf_n)
    D = np.minimum(Wm + 1, CBITS)  # leaf level of each particle
    Wc = window_level(codes, max(ncrit, max_leaf_n))
    Lc = np.minimum(Wc + 1, CBITS)  # critical level of each particle
    lo = delta + 1
    cnt = np.maximum(0, D - lo + 1)
    base = np.concatenate([[0], np.cumsum(cnt)])
    M = int(base[-1])
    begin = np.zeros(M, dtype=np.uint64)
    level = np.zeros(M, dtype=np.uint64)
    for i in np.nonzero(cnt)[0]:
        for l in range(lo[i], D[i] + 1):
            k = base[i] + (l - lo[i])
            begin[k] = i
            level[k] = l
    end = np.zeros(M, dtype=np.uint64)
    code = np.zeros(M, dtype=np.uint64)
    nch = np.zeros(M, dtype=np.uint64)
    iscrit = np.zeros(M, dtype=bool)
    for k in range(M):
        i = int(begin[k]); l = int(level[k])
        sh = 3 * (CBITS - l)
        pre = int(codes[i]) >> sh
        # end = first j > i with (codes[j] >> sh) != pre
        if sh >= 64:
            e = N
        else:
            e = int(np.searchsorted(codes, np.uint64(((pre + 1) << sh) if sh < 64 else 0), side="left")) if l > 0 else N
        end[k] = e
        code[k] = (1 << (3 * l)) | pre
        nch[k] = base[e] - k - 1
        iscrit[k] = (l == Lc[i])
    return dict(begin=begin, end=end, n_children=nch, code=code, level=level, iscrit=iscrit)
