// build.cu — particle packing, box deduction, Morton encoding, permutation and the non-recursive octree
// build for sm_100a. All kernels here are HBM/L2-bound integer or byte work: coalesced, one pass each.
//
// Reference functions replaced (include/rakau/tree.hpp):
//   determine_box_size 1278-1319, disc_single_coord 381-429, morton_encoder 222-242 (libmorton sLUT),
//   apply_isort 484-507, perm_to_inv_perm 1248-1262, build_tree 932-1111 (+ build_tree_ser_impl 724-833),
//   compute_node_properties 1116-1237, get_node_centre 450-482.
//
// Tree build formulation (validated against the recursive reference semantics by tests/build_model.py):
// with sorted codes c[i], let delta(i) = #leading 3-bit digits shared by c[i-1], c[i] (delta(0) = -1) and
// W_w(i) = deepest level at which the cell containing i holds more than w particles (-1 if none)
//        = max over j in [i, i+w], w <= j < N, of shared_digits(c[j-w], c[j]).
// Leaf level D(i) = min(W_maxleaf(i)+1, 21); critical level Lc(i) = min(W_max(ncrit,maxleaf)(i)+1, 21).
// The nodes that begin at particle i are exactly the levels delta(i)+1 .. D(i); DFS pre-order is
// (begin asc, level asc); the device keeps nodes level-major (BFS) so children are contiguous.

#include "common.cuh"
#include "scan.cuh"

#include <cmath>
#include <cstdlib>
#include <type_traits>

namespace rk
{

namespace
{

// ---------------------------------------------------------------------------------------------------
// pack / unpack / abs-max
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 abs_bits(float v) { return __float_as_uint(v) & 0x7fffffffu; }
__device__ __forceinline__ u64 abs_bits(double v)
{
    return static_cast<u64>(__double_as_longlong(v)) & 0x7fffffffffffffffull;
}

__device__ __forceinline__ void block_max_to_global(u64 v, u64 *out)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const u64 t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    if ((threadIdx.x & 31) == 0 && v) {
        atomicMax(out, v);
    }
}

// x,y,z,m SoA -> packed vec4; max |coord| as an unsigned bit pattern (non-negative IEEE values order like
// unsigned integers; Inf/NaN patterns sort above every finite value, which is how non-finite input is found).
template <typename F>
__global__ void __launch_bounds__(256) pack_absmax_kernel(const F *__restrict__ x, const F *__restrict__ y,
                                                          const F *__restrict__ z, const F *__restrict__ m,
                                                          vec4<F> *__restrict__ out, size_t n, u64 *__restrict__ absmax)
{
    u64 mx = 0;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const F a = x[i], b = y[i], c = z[i];
        out[i] = make_vec4<F>(a, b, c, m ? m[i] : F(0)); // NULL masses arrive with the gather (late_m)
        const u64 ba = abs_bits(a), bb = abs_bits(b), bc = abs_bits(c);
        mx = ba > mx ? ba : mx;
        mx = bb > mx ? bb : mx;
        mx = bc > mx ? bc : mx;
    }
    block_max_to_global(mx, absmax);
}

// Overwrite selected components (NULL = keep) of an already packed array and recompute the abs-max.
template <typename F>
__global__ void __launch_bounds__(256) set_coords_kernel(vec4<F> *__restrict__ p, const F *__restrict__ x,
                                                         const F *__restrict__ y, const F *__restrict__ z,
                                                         const F *__restrict__ m, size_t n, u64 *__restrict__ absmax)
{
    u64 mx = 0;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        vec4<F> v = p[i];
        if (x) {
            v.x = x[i];
        }
        if (y) {
            v.y = y[i];
        }
        if (z) {
            v.z = z[i];
        }
        if (m) {
            v.w = m[i];
        }
        p[i] = v;
        const u64 ba = abs_bits(v.x), bb = abs_bits(v.y), bc = abs_bits(v.z);
        mx = ba > mx ? ba : mx;
        mx = bb > mx ? bb : mx;
        mx = bc > mx ? bc : mx;
    }
    if (absmax) {
        block_max_to_global(mx, absmax);
    }
}

template <typename F>
__global__ void __launch_bounds__(256) unpack_kernel(const vec4<F> *__restrict__ in, F *__restrict__ x, F *__restrict__ y,
                                                     F *__restrict__ z, F *__restrict__ m, size_t n)
{
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i < n) {
        const vec4<F> v = in[i];
        if (x) {
            x[i] = v.x;
        }
        if (y) {
            y[i] = v.y;
        }
        if (z) {
            z[i] = v.z;
        }
        if (m) {
            m[i] = v.w;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// discretise + Morton encode (bit-exact with disc_single_coord + m3D_e_sLUT)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 spread3(u64 v)
{
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

__host__ __device__ inline u64 compact3(u64 v)
{
    v &= 0x1249249249249249ull;
    v = (v ^ (v >> 2)) & 0x10c30c30c30c30c3ull;
    v = (v ^ (v >> 4)) & 0x100f00f00f00f00full;
    v = (v ^ (v >> 8)) & 0x1f0000ff0000ffull;
    v = (v ^ (v >> 16)) & 0x1f00000000ffffull;
    v = (v ^ (v >> 32)) & 0x1fffffull;
    return v;
}

// Returns 0 on success, else the error class (1 non-finite, 2 fp out of bounds, 3 int out of bounds).
template <typename F>
__device__ __forceinline__ u32 disc_coord(F x, F inv_box, u64 &out)
{
    const F factor = F(2097152); // 2^21
    F tmp = rn_fma(x, inv_box, F(0.5));
    tmp = rn_mul(tmp, factor);
    if (!isfinite(tmp)) {
        return 1;
    }
    if (tmp < F(0) || tmp >= factor) {
        return 2;
    }
    out = static_cast<u64>(tmp); // truncation, value in [0, 2^21)
    return out >= 2097152ull ? 3u : 0u;
}

template <typename F>
__global__ void __launch_bounds__(256)
    encode_kernel(const vec4<F> *__restrict__ p, u64 *__restrict__ codes, size_t n, F inv_box, u64 *__restrict__ err)
{
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i >= n) {
        return;
    }
    const vec4<F> v = p[i];
    u64 dx = 0, dy = 0, dz = 0;
    u32 e = disc_coord(v.x, inv_box, dx), dim = 0;
    if (!e) {
        e = disc_coord(v.y, inv_box, dy);
        dim = 1;
    }
    if (!e) {
        e = disc_coord(v.z, inv_box, dz);
        dim = 2;
    }
    if (e) {
        // first offending particle wins: key = index << 8 | dim << 4 | class
        atomicMin(err, (static_cast<u64>(i) << 8) | (dim << 4) | e);
        codes[i] = 0;
        return;
    }
    codes[i] = spread3(dx) | (spread3(dy) << 1) | (spread3(dz) << 2);
}

template <typename F>
__global__ void __launch_bounds__(256)
    gather_kernel(const vec4<F> *__restrict__ pin, const u32 *__restrict__ idx, vec4<F> *__restrict__ pout, size_t n,
                  const F *__restrict__ late_m)
{
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i < n) {
        const u32 j = idx[i];
        vec4<F> v = pin[j];
        if (late_m) { // masses whose upload overlapped the sort
            v.w = late_m[j];
        }
        pout[i] = v;
    }
}

__global__ void __launch_bounds__(256)
    perm_first_kernel(const u32 *__restrict__ last_perm, u32 *__restrict__ perm, u32 *__restrict__ inv_perm, size_t n)
{
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i < n) {
        const u32 p = last_perm[i];
        perm[i] = p;
        if (inv_perm) {
            inv_perm[p] = static_cast<u32>(i);
        }
    }
}

// perm_to_inv_perm, tree.hpp:1248-1262 (run lazily: only exact_*_o and the inv_perm() getter need it; the
// scattered 4-byte writes cost 3 ms at 128 M particles).
__global__ void __launch_bounds__(256) perm_invert_kernel(const u32 *__restrict__ perm, u32 *__restrict__ inv_perm, size_t n)
{
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i < n) {
        inv_perm[perm[i]] = static_cast<u32>(i);
    }
}

// apply_isort(m_perm, m_last_perm) + perm_to_inv_perm, tree.hpp:3725-3732.
__global__ void __launch_bounds__(256)
    perm_compose_kernel(const u32 *__restrict__ old_perm, const u32 *__restrict__ last_perm, u32 *__restrict__ new_perm,
                        u32 *__restrict__ inv_perm, size_t n)
{
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i < n) {
        const u32 p = old_perm[last_perm[i]];
        new_perm[i] = p;
        if (inv_perm) {
            inv_perm[p] = static_cast<u32>(i);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// topology
// ---------------------------------------------------------------------------------------------------
// delta(i) and the two window arrays P'_w(j) = shared_digits(c[j-w], c[j]) (w <= j < n), else -1.
__global__ void __launch_bounds__(256) delta_window_kernel(const u64 *__restrict__ codes, size_t n, size_t w1, size_t w2,
                                                           i8 *__restrict__ delta, i8 *__restrict__ p1,
                                                           i8 *__restrict__ p2)
{
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i >= n) {
        return;
    }
    const u64 c = codes[i];
    delta[i] = i ? static_cast<i8>(shared_digits(codes[i - 1], c)) : i8(-1);
    p1[i] = (i >= w1) ? static_cast<i8>(shared_digits(codes[i - w1], c)) : i8(-1);
    if (p2) {
        p2[i] = (i >= w2) ? static_cast<i8>(shared_digits(codes[i - w2], c)) : i8(-1);
    }
}

// out[j] = max(in[j], in[j + step]) (missing = -1). 4 bytes per thread.
__global__ void __launch_bounds__(256) window_double_kernel(const i8 *__restrict__ in, i8 *__restrict__ out, size_t n,
                                                            size_t step)
{
    const size_t j = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (j >= n) {
        return;
    }
    const i8 a = in[j];
    const i8 b = (j + step < n) ? in[j + step] : i8(-1);
    out[j] = a > b ? a : b;
}

// lvl[i] = min(21, 1 + max(A[i], A[i + off])) where A holds window maxima of width 2^K and off = w+1-2^K.
__global__ void __launch_bounds__(256) window_final_kernel(const i8 *__restrict__ A, i8 *__restrict__ lvl, size_t n,
                                                           size_t off)
{
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i >= n) {
        return;
    }
    const i8 a = A[i];
    const i8 b = (i + off < n) ? A[i + off] : i8(-1);
    const int w = a > b ? a : b;
    lvl[i] = static_cast<i8>(w + 1 > CBITS ? CBITS : w + 1);
}


// Fused form of delta_window + window_double x log2(w) + window_final for windows that fit in shared memory:
// one CTA computes delta, the leaf level and the critical level of WIN_TILE particles from a staged tile of
// codes (halo of w2 on both sides); the sliding-window maxima are built by doubling in shared memory.
constexpr int WIN_TILE = 2048;
constexpr int WIN_MAXW = 2048;

// Word of four packed levels starting at byte 4 * wi + sh of the packed array a (bytes beyond the array read as -1).
__device__ __forceinline__ u32 shifted_word(const u32 *__restrict__ a, int wi, int sh, int nwords)
{
    const u32 lo = wi < nwords ? a[wi] : 0xffffffffu, hi = wi + 1 < nwords ? a[wi + 1] : 0xffffffffu;
    return sh == 0 ? lo : __byte_perm(lo, hi, 0x3210u + 0x1111u * static_cast<u32>(sh));
}

// lvl[i] = min(21, 1 + max_{j in [i, i + w], w <= j < n} shared_digits(c[j - w], c[j])) for the WIN_TILE particles of
// the tile: the sliding-window maxima are built by doubling in shared memory, four signed bytes per 32-bit word
// (__vmaxs4; a shift by k bytes is a word offset plus a byte permute), so a round costs 2 word operations per thread.
__device__ __forceinline__ void window_levels_smem(const u64 *__restrict__ sc, int w2, int w, size_t i0, size_t n,
                                                   u32 *__restrict__ A0, u32 *__restrict__ A1, i8 *__restrict__ out)
{
    const int len = WIN_TILE + w, nwords = (len + 3) / 4;
    for (int wi = threadIdx.x; wi < nwords; wi += blockDim.x) {
        u32 v = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int q = 4 * wi + b;
            const size_t j = i0 + q;
            const int sd = (q < len && j >= size_t(w) && j < n) ? shared_digits(sc[q + w2 - w], sc[q + w2]) : -1;
            v |= (static_cast<u32>(sd) & 0xffu) << (8 * b);
        }
        A0[wi] = v;
    }
    __syncthreads();
    int k = 1;
    u32 *in = A0, *ot = A1;
    while (2 * k <= w + 1) {
        for (int wi = threadIdx.x; wi < nwords; wi += blockDim.x) {
            ot[wi] = __vmaxs4(in[wi], shifted_word(in, wi + (k >> 2), k & 3, nwords));
        }
        __syncthreads();
        u32 *t = in;
        in = ot;
        ot = t;
        k *= 2;
    }
    const int off = w + 1 - k;
    for (int wi = threadIdx.x; wi < WIN_TILE / 4; wi += blockDim.x) {
        u32 v = __vmaxs4(in[wi], shifted_word(in, wi + (off >> 2), off & 3, nwords));
        v = __vminu4(__vadd4(v, 0x01010101u), 0x15151515u); // 1 + max (>= 0), capped at CBITS = 21
        const size_t i = i0 + 4 * size_t(wi);
        if (i + 4 <= n) {
            *reinterpret_cast<u32 *>(out + i) = v;
        } else {
            for (int b = 0; b < 4 && i + b < n; ++b) {
                out[i + b] = static_cast<i8>((v >> (8 * b)) & 0xffu);
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) window_fused_kernel(const u64 *__restrict__ codes, size_t n, int w1, int w2,
                                                           i8 *__restrict__ delta, i8 *__restrict__ lvl_leaf,
                                                           i8 *__restrict__ lvl_crit)
{
    extern __shared__ __align__(16) unsigned char win_smem[];
    u64 *sc = reinterpret_cast<u64 *>(win_smem); // codes of particles i0 - w2 .. i0 + WIN_TILE + w2
    u32 *A0 = reinterpret_cast<u32 *>(sc + WIN_TILE + 2 * w2);
    u32 *A1 = A0 + (WIN_TILE + w2 + 3) / 4 + 1;
    const size_t i0 = size_t(blockIdx.x) * WIN_TILE;
    {
        // the tile itself: all eight loads of a thread in flight before the first store; then the two halos
        u64 v[WIN_TILE / 256];
#pragma unroll
        for (int k = 0; k < WIN_TILE / 256; ++k) {
            const size_t j = i0 + size_t(k) * 256 + threadIdx.x;
            v[k] = j < n ? codes[j] : 0ull;
        }
        for (int s = threadIdx.x; s < 2 * w2; s += blockDim.x) {
            const long long j = s < w2 ? static_cast<long long>(i0) - w2 + s : static_cast<long long>(i0) + WIN_TILE + (s - w2);
            sc[s < w2 ? s : WIN_TILE + s] = (j >= 0 && static_cast<size_t>(j) < n) ? codes[j] : 0ull;
        }
#pragma unroll
        for (int k = 0; k < WIN_TILE / 256; ++k) {
            sc[w2 + k * 256 + threadIdx.x] = v[k];
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < WIN_TILE; q += blockDim.x) {
        const size_t i = i0 + q;
        if (i < n) {
            delta[i] = i ? static_cast<i8>(shared_digits(sc[q + w2 - 1], sc[q + w2])) : i8(-1);
        }
    }
    window_levels_smem(sc, w2, w1, i0, n, A0, A1, lvl_leaf);
    if (w2 != w1) {
        window_levels_smem(sc, w2, w2, i0, n, A0, A1, lvl_crit);
    } else {
        for (int q = threadIdx.x; q < WIN_TILE; q += blockDim.x) {
            const size_t i = i0 + q;
            if (i < n) {
                lvl_crit[i] = lvl_leaf[i];
            }
        }
    }
}

__global__ void __launch_bounds__(256) fill_i8_kernel(i8 *p, size_t n, i8 v)
{
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i < n) {
        p[i] = v;
    }
}

// Per tile of TOPO_TILE particles: number of nodes beginning in the tile at each level (rows 0..21) and the
// number of critical nodes beginning in the tile (row 22). tilecnt is row-major [NLEVELS+1][ntiles].
// Each warp only visits the levels that occur among its 32 particles (REDUX min/max): in a Plummer sphere the
// nodes beginning at neighbouring particles span 2-4 levels, not 22.
// Both topology kernels run TOPO_THREADS threads with TOPO_IT rows of 32 particles per warp, all the loads of a thread
// issued before the first one is used (measured at 4 M / 32 M particles, topology phase: 1 row 0.183 / 1.19 ms, 2 rows
// 0.176 / 1.11 ms, 4 rows 0.199 / 1.21 ms, 8 rows 0.213 / 1.32 ms).
constexpr int TOPO_THREADS = 256;
constexpr int TOPO_IT = TOPO_TILE / TOPO_THREADS;
static_assert(TOPO_TILE % TOPO_THREADS == 0 && TOPO_IT >= 1, "tile = whole rows per warp");

__global__ void __launch_bounds__(TOPO_THREADS)
    topo_count_kernel(const i8 *__restrict__ delta, const i8 *__restrict__ lvl_leaf, const i8 *__restrict__ lvl_crit,
                      size_t n, u32 ntiles, u32 *__restrict__ tilecnt)
{
    __shared__ u32 cnt[NLEVELS + 1];
    if (threadIdx.x < NLEVELS + 1) {
        cnt[threadIdx.x] = 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int lo[TOPO_IT], hi[TOPO_IT];
    bool critb[TOPO_IT];
#pragma unroll
    for (int j = 0; j < TOPO_IT; ++j) {
        const size_t i = size_t(blockIdx.x) * TOPO_TILE + size_t(w * TOPO_IT + j) * 32 + lane;
        const bool valid = i < n;
        lo[j] = valid ? delta[i] + 1 : 64;
        hi[j] = valid ? lvl_leaf[i] : -1;
        critb[j] = valid && (lo[j] <= lvl_crit[i]);
    }
#pragma unroll
    for (int j = 0; j < TOPO_IT; ++j) {
        const int wlo = __reduce_min_sync(0xffffffffu, lo[j]), whi = __reduce_max_sync(0xffffffffu, hi[j]);
        for (int l = wlo; l <= whi; ++l) {
            const u32 b = __ballot_sync(0xffffffffu, lo[j] <= l && l <= hi[j]);
            if (lane == 0 && b) {
                atomicAdd(&cnt[l], __popc(b));
            }
        }
        const u32 b = __ballot_sync(0xffffffffu, critb[j]);
        if (lane == 0 && b) {
            atomicAdd(&cnt[NLEVELS], __popc(b));
        }
    }
    __syncthreads();
    if (threadIdx.x < NLEVELS + 1) {
        tilecnt[size_t(threadIdx.x) * ntiles + blockIdx.x] = cnt[threadIdx.x];
    }
}

// Emits the nodes that begin in this tile (BFS positions from the scanned tile counts) and the critical
// nodes. nodeB.y (end) and the child count are filled by topo_finalize_kernel / topo_children_kernel.
__global__ void __launch_bounds__(TOPO_THREADS)
    topo_emit_kernel(const i8 *__restrict__ delta, const i8 *__restrict__ lvl_leaf, const i8 *__restrict__ lvl_crit,
                     size_t n, u32 ntiles, const u32 *__restrict__ tilecnt /* scanned */, level_table lt,
                     uint4 *__restrict__ nodeB, u32 *__restrict__ node_dfs, u32 *__restrict__ dfsbase,
                     u32 *__restrict__ crit_node, u32 *__restrict__ crit_begin, u32 n_nodes, u32 n_crit)
{
    constexpr int NR = TOPO_TILE / 32;   // rows of 32 particles in the tile
    __shared__ u32 wc[NLEVELS + 1][NR]; // per level: count of this row, then BFS base of this row
    __shared__ u32 wprefix[NR];         // nodes (all levels) beginning before this row's first particle
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 ltm = lanemask_lt();
    int lo[TOPO_IT], hi[TOPO_IT], lc[TOPO_IT];
    bool critb[TOPO_IT];
#pragma unroll
    for (int j = 0; j < TOPO_IT; ++j) {
        const size_t i = size_t(blockIdx.x) * TOPO_TILE + size_t(w * TOPO_IT + j) * 32 + lane;
        const bool valid = i < n;
        lo[j] = valid ? delta[i] + 1 : 64;
        hi[j] = valid ? lvl_leaf[i] : -1;
        lc[j] = valid ? lvl_crit[i] : -1;
        critb[j] = valid && (lo[j] <= lc[j]);
    }
#pragma unroll
    for (int j = 0; j < TOPO_IT; ++j) {
        const int row = w * TOPO_IT + j;
        const int wlo = __reduce_min_sync(0xffffffffu, lo[j]), whi = __reduce_max_sync(0xffffffffu, hi[j]);
        if (lane <= NLEVELS) {
            wc[lane][row] = 0;
        }
        __syncwarp();
        for (int l = wlo; l <= whi; ++l) {
            const u32 b = __ballot_sync(0xffffffffu, lo[j] <= l && l <= hi[j]);
            if (lane == 0) {
                wc[l][row] = __popc(b);
            }
        }
        const u32 b = __ballot_sync(0xffffffffu, critb[j]);
        if (lane == 0) {
            wc[NLEVELS][row] = __popc(b);
        }
    }
    __syncthreads();
    // per level: exclusive scan over the rows of the tile, offset by the tile's base and the level's base
    if (threadIdx.x < NLEVELS + 1) {
        const int l = threadIdx.x;
        u32 run = tilecnt[size_t(l) * ntiles + blockIdx.x] + (l < NLEVELS ? lt.base[l] : 0u);
        for (int k = 0; k < NR; ++k) {
            const u32 c = wc[l][k];
            wc[l][k] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < TOPO_IT; ++j) {
        const int row = w * TOPO_IT + j;
        const size_t i = size_t(blockIdx.x) * TOPO_TILE + size_t(row) * 32 + lane;
        const bool valid = i < n;
        const int wlo = __reduce_min_sync(0xffffffffu, lo[j]), whi = __reduce_max_sync(0xffffffffu, hi[j]);
        // DFS base of the row = sum over levels of (BFS base of the row at that level - base of the level)
        {
            u32 v = (lane < NLEVELS) ? wc[lane][row] - lt.base[lane] : 0u;
            v = __reduce_add_sync(0xffffffffu, v);
            if (lane == 0) {
                wprefix[row] = v;
            }
        }
        __syncwarp();
        // DFS base of particle i = nodes beginning at particles < i
        u32 dfsb = wprefix[row];
        for (int l = wlo; l <= whi; ++l) {
            const u32 b = __ballot_sync(0xffffffffu, lo[j] <= l && l <= hi[j]);
            dfsb += __popc(b & ltm);
        }
        if (valid) {
            dfsbase[i] = dfsb;
            if (i == n - 1) {
                dfsbase[n] = n_nodes;
            }
        }
        u32 crit_rank = 0;
        {
            const u32 b = __ballot_sync(0xffffffffu, critb[j]);
            crit_rank = wc[NLEVELS][row] + __popc(b & ltm);
        }
        // Emit. The rank at level l+1 is the first child of the node at level l.
        u32 prev_rank = 0;
        bool prev_pred = false;
        for (int l = wlo; l <= whi + 1; ++l) {
            bool pred = false;
            u32 r = 0;
            if (l <= whi) {
                pred = lo[j] <= l && l <= hi[j];
                const u32 b = __ballot_sync(0xffffffffu, pred);
                r = wc[l][row] + __popc(b & ltm);
            }
            if (prev_pred) {
                nodeB[prev_rank] = make_uint4(static_cast<u32>(i), 0u, pred ? r : 0u, static_cast<u32>(l - 1) << 8);
                node_dfs[prev_rank] = dfsb + static_cast<u32>(l - 1 - lo[j]);
                if (critb[j] && lc[j] == l - 1) {
                    crit_node[crit_rank] = prev_rank;
                    crit_begin[crit_rank] = static_cast<u32>(i);
                }
            }
            prev_pred = pred;
            prev_rank = r;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        crit_begin[n_crit] = static_cast<u32>(n);
    }
}

// Per node: end of its particle range (galloping + binary search on the sorted codes), number of children
// (children are contiguous in BFS order), number of descendants (the reference's n_children).
__global__ void __launch_bounds__(256)
    topo_finalize_kernel(const u64 *__restrict__ codes, size_t n, uint4 *__restrict__ nodeB,
                         const u32 *__restrict__ node_dfs, const u32 *__restrict__ dfsbase, u32 *__restrict__ node_ndesc,
                         level_table lt, u32 n_nodes)
{
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_nodes) {
        return;
    }
    uint4 nb = nodeB[k];
    const u32 b = nb.x, level = nb.w >> 8;
    u32 e;
    if (level == 0) {
        e = static_cast<u32>(n);
    } else {
        const int sh = 3 * (CBITS - static_cast<int>(level));
        const u64 pre = codes[b] >> sh;
        // gallop: largest known index inside the cell is lo; hi is the first index known outside (or n).
        size_t lo = b, step = 1, hi = n;
        while (lo + step < n) {
            if ((codes[lo + step] >> sh) == pre) {
                lo += step;
                step <<= 1;
            } else {
                hi = lo + step;
                break;
            }
        }
        // invariant: cell contains lo, does not contain hi (or hi == n)
        while (hi - lo > 1) {
            const size_t mid = lo + (hi - lo) / 2;
            if ((codes[mid] >> sh) == pre) {
                lo = mid;
            } else {
                hi = mid;
            }
        }
        e = static_cast<u32>(hi);
    }
    nb.y = e;
    nodeB[k] = nb;
    node_ndesc[k] = dfsbase[e] - node_dfs[k] - 1u;
}

// Second finalize pass (needs every node's end): child count.
__global__ void __launch_bounds__(256)
    topo_children_kernel(uint4 *__restrict__ nodeB, level_table lt, u32 n_nodes)
{
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_nodes) {
        return;
    }
    uint4 nb = nodeB[k];
    const u32 level = nb.w >> 8;
    u32 nch = 0;
    if (nb.z) {
        const u32 fc = nb.z, lim = lt.base[level + 2]; // end of level+1 segment
#pragma unroll
        for (u32 t = 0; t < 8; ++t) {
            if (fc + t < lim && nodeB[fc + t].x < nb.y) {
                ++nch;
            }
        }
    }
    nb.w = (level << 8) | nch;
    nodeB[k] = nb;
}

__global__ void __launch_bounds__(256) max_group_kernel(const u32 *__restrict__ crit_begin, u32 n_crit, u32 *__restrict__ out)
{
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    u32 v = 0;
    if (j < n_crit) {
        v = crit_begin[j + 1] - crit_begin[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const u32 t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    if ((threadIdx.x & 31) == 0 && v) {
        atomicMax(out, v);
    }
}

// ---------------------------------------------------------------------------------------------------
// node properties
// ---------------------------------------------------------------------------------------------------
struct dsum4 {
    double m, x, y, z;
};
__device__ __forceinline__ void dsum_add_particle(dsum4 &s, double px, double py, double pz, double pm)
{
    s.m += pm;
    s.x = fma(pm, px, s.x);
    s.y = fma(pm, py, s.y);
    s.z = fma(pm, pz, s.z);
}
__device__ __forceinline__ void dsum_warp_reduce(dsum4 &s)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s.m += __shfl_xor_sync(0xffffffffu, s.m, o);
        s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
        s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
        s.z += __shfl_xor_sync(0xffffffffu, s.z, o);
    }
}

// One warp per chunk of PROPS_CHUNK particles: (sum m, sum m*x, sum m*y, sum m*z) in double.
template <typename F>
__global__ void __launch_bounds__(256)
    chunk_sums_kernel(const vec4<F> *__restrict__ p, size_t n, u32 nchunks, double *__restrict__ out)
{
    const u32 c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= nchunks) {
        return;
    }
    const int lane = threadIdx.x & 31;
    dsum4 s{0, 0, 0, 0};
    const size_t b = size_t(c) * PROPS_CHUNK;
#pragma unroll
    for (int j = 0; j < PROPS_CHUNK / 32; ++j) {
        const size_t i = b + size_t(j) * 32 + lane;
        if (i < n) {
            const vec4<F> v = p[i];
            dsum_add_particle(s, v.x, v.y, v.z, v.w);
        }
    }
    dsum_warp_reduce(s);
    if (lane == 0) {
        out[size_t(c) * 4 + 0] = s.m;
        out[size_t(c) * 4 + 1] = s.x;
        out[size_t(c) * 4 + 2] = s.y;
        out[size_t(c) * 4 + 3] = s.z;
    }
}

// Second level: one warp per super-chunk of PROPS_CHUNK chunks (65536 particles), so that the few huge nodes near
// the root need O(N / 65536 / 32) iterations instead of O(N / 256 / 32) (9 ms at 128 M particles before).
__global__ void __launch_bounds__(256) chunk_sums2_kernel(const double *__restrict__ c1, u32 nchunks, u32 nsuper,
                                                           double *__restrict__ out)
{
    const u32 sc = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (sc >= nsuper) {
        return;
    }
    const int lane = threadIdx.x & 31;
    dsum4 s{0, 0, 0, 0};
    const u32 b = sc * PROPS_CHUNK;
#pragma unroll
    for (int j = 0; j < PROPS_CHUNK / 32; ++j) {
        const u32 c = b + u32(j) * 32 + lane;
        if (c < nchunks) {
            const double4 cs = *reinterpret_cast<const double4 *>(c1 + size_t(c) * 4);
            s.m += cs.x;
            s.x += cs.y;
            s.y += cs.z;
            s.z += cs.w;
        }
    }
    dsum_warp_reduce(s);
    if (lane == 0) {
        out[size_t(sc) * 4 + 0] = s.m;
        out[size_t(sc) * 4 + 1] = s.x;
        out[size_t(sc) * 4 + 2] = s.y;
        out[size_t(sc) * 4 + 3] = s.z;
    }
}

template <typename F>
struct level_dims {
    F dim[NLEVELS];  // box / 2^level  (get_node_dim, tree.hpp:443-448)
    F cell;          // box * (1 / 2^21)
    F half_box;      // box * (1/2)
};

// get_node_centre, tree.hpp:450-482, with explicitly rounded operations.
template <typename F>
__device__ __forceinline__ void node_centre_dev(F out[3], u64 first_code, u32 level, const level_dims<F> &ld)
{
    const int sh = 3 * (CBITS - static_cast<int>(level));
    const u64 cell_code = (level == 0) ? 0ull : ((first_code >> sh) << sh);
    const F half_dim = rn_mul(ld.dim[level], F(0.5));
    const F off = rn_sub(half_dim, ld.half_box);
    const u64 d[3] = {compact3(cell_code), compact3(cell_code >> 1), compact3(cell_code >> 2)};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        out[j] = rn_fma(static_cast<F>(d[j]), ld.cell, off);
    }
}

// Node sums with G lanes per node. The reduction shape (lane-strided partial sums over head particles, whole
// chunks and tail particles, then a fixed shuffle tree) depends only on the node's range, never on the mass
// values, so scaling all masses by a power of two scales every partial exactly (reference test
// update_masses.cpp:56-68).
template <typename F, int G>
__device__ __forceinline__ dsum4 node_sum(const vec4<F> *__restrict__ p, const double *__restrict__ chunks,
                                          const double *__restrict__ chunks2, u32 b, u32 e, int gl, bool active)
{
    dsum4 s{0, 0, 0, 0};
    auto add_particles = [&](u32 i0, u32 i1) {
        for (u32 i = i0 + gl; i < i1; i += G) {
            const vec4<F> v = p[i];
            dsum_add_particle(s, v.x, v.y, v.z, v.w);
        }
    };
    auto add_sums = [&](const double *__restrict__ arr, u32 c0, u32 c1) {
        for (u32 c = c0 + gl; c < c1; c += G) {
            const double4 cs = *reinterpret_cast<const double4 *>(arr + size_t(c) * 4);
            s.m += cs.x;
            s.x += cs.y;
            s.y += cs.z;
            s.z += cs.w;
        }
    };
    if (active) {
        const u32 cb = (b + PROPS_CHUNK - 1) / PROPS_CHUNK, ce = e / PROPS_CHUNK;
        if (cb >= ce) {
            add_particles(b, e);
        } else {
            add_particles(b, cb * PROPS_CHUNK);
            const u32 sb = (cb + PROPS_CHUNK - 1) / PROPS_CHUNK, se = ce / PROPS_CHUNK;
            if (sb >= se) {
                add_sums(chunks, cb, ce);
            } else {
                add_sums(chunks, cb, sb * PROPS_CHUNK);
                add_sums(chunks2, sb, se);
                add_sums(chunks, se * PROPS_CHUNK, ce);
            }
            add_particles(ce * PROPS_CHUNK, e);
        }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        s.m += __shfl_xor_sync(0xffffffffu, s.m, o);
        s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
        s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
        s.z += __shfl_xor_sync(0xffffffffu, s.z, o);
    }
    return s;
}

// Mass, centre of mass (geometric centre for a massless node, tree.hpp:1176-1185), delta for bh_geom.
template <typename F>
__device__ __forceinline__ void node_finalize(const dsum4 &s, u32 k, u32 b, u32 level, const u64 *__restrict__ codes,
                                              vec4<F> *__restrict__ nodeA, F *__restrict__ node_delta, int mac,
                                              const level_dims<F> &ld, u64 *__restrict__ err)
{
    const F tot = static_cast<F>(s.m);
    F com[3], geo[3] = {F(0), F(0), F(0)};
    const u64 c0 = codes[b];
    if (mac == 1) {
        node_centre_dev(geo, c0, level, ld);
    }
    if (tot == F(0)) {
        if (mac == 1) {
            com[0] = geo[0];
            com[1] = geo[1];
            com[2] = geo[2];
        } else {
            node_centre_dev(com, c0, level, ld);
        }
    } else {
        const double inv = 1.0 / s.m;
        com[0] = static_cast<F>(s.x * inv);
        com[1] = static_cast<F>(s.y * inv);
        com[2] = static_cast<F>(s.z * inv);
    }
    u32 ecode = 0;
    if (!isfinite(com[0]) || !isfinite(com[1]) || !isfinite(com[2])) {
        ecode = 5;
    } else if (!isfinite(tot)) {
        ecode = 6;
    }
    nodeA[k] = make_vec4<F>(com[0], com[1], com[2], tot);
    if (mac == 1) {
        const F d0 = rn_sub(com[0], geo[0]), d1 = rn_sub(com[1], geo[1]), d2 = rn_sub(com[2], geo[2]);
        F dd = rn_mul(d0, d0);
        dd = rn_fma(d1, d1, dd);
        dd = rn_fma(d2, d2, dd);
        const F dl = rn_sqrt(dd);
        node_delta[k] = dl;
        if (!ecode && !isfinite(dl)) {
            ecode = 7;
        }
    }
    if (ecode) {
        atomicMin(err, (static_cast<u64>(k) << 8) | ecode);
    }
}

#ifndef RK_PROPS_LANES
#define RK_PROPS_LANES 1
#endif
constexpr int PROPS_LANES = RK_PROPS_LANES; // lanes per small node (86 % are leaves of ~5 particles: the kernel is bound by its
                                             // dependent loads, so nodes in flight count - measured 16 / 8 / 4 / 2 / 1 lanes: 0.255 / 0.167 / 0.120 / 0.097 / 0.085 ms at 4 M)
constexpr u32 PROPS_SMALL = 64; // nodes up to this size are reduced by PROPS_LANES lanes, larger ones by a full warp

// PROPS_LANES lanes per node (86 % of the nodes are leaves with <= 16 particles); larger nodes are queued for the
// warp-per-node kernel.
template <typename F>
__global__ void __launch_bounds__(256)
    node_props_small_kernel(const vec4<F> *__restrict__ p, const u64 *__restrict__ codes,
                            const double *__restrict__ chunks, const uint4 *__restrict__ nodeB,
                            vec4<F> *__restrict__ nodeA, F *__restrict__ node_delta, u32 n_nodes, int mac,
                            level_dims<F> ld, u64 *__restrict__ err, u32 *__restrict__ big_list,
                            u32 *__restrict__ big_count)
{
    const u32 k = (blockIdx.x * blockDim.x + threadIdx.x) / PROPS_LANES;
    const int gl = threadIdx.x & (PROPS_LANES - 1);
    const bool in_range = k < n_nodes;
    uint4 nb = make_uint4(0, 0, 0, 0);
    if (in_range) {
        nb = nodeB[k];
    }
    const bool small = in_range && (nb.y - nb.x) <= PROPS_SMALL;
    const dsum4 s = node_sum<F, PROPS_LANES>(p, chunks, nullptr, nb.x, nb.y, gl, small);
    if (gl == 0 && in_range) {
        if (small) {
            node_finalize<F>(s, k, nb.x, nb.w >> 8, codes, nodeA, node_delta, mac, ld, err);
        } else {
            big_list[atomicAdd(big_count, 1u)] = k;
        }
    }
}

// One warp per queued node, persistent grid.
template <typename F>
__global__ void __launch_bounds__(256)
    node_props_big_kernel(const vec4<F> *__restrict__ p, const u64 *__restrict__ codes, const double *__restrict__ chunks,
                          const double *__restrict__ chunks2, const uint4 *__restrict__ nodeB,
                          vec4<F> *__restrict__ nodeA, F *__restrict__ node_delta, int mac, level_dims<F> ld,
                          u64 *__restrict__ err, const u32 *__restrict__ big_list, const u32 *__restrict__ big_count)
{
    const u32 nbig = *big_count;
    const int lane = threadIdx.x & 31;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nbig; q += nwarps) {
        const u32 k = big_list[q];
        const uint4 nb = nodeB[k];
        const dsum4 s = node_sum<F, 32>(p, chunks, chunks2, nb.x, nb.y, lane, true);
        if (lane == 0) {
            node_finalize<F>(s, k, nb.x, nb.w >> 8, codes, nodeA, node_delta, mac, ld, err);
        }
    }
}

// ---- bottom-up variant (large trees: from PROPS_BOTTOMUP_MIN particles) ------------------------------------------
// The kernels above sum every node of <= 64 particles from its particles, so a particle is re-read once per
// ancestor below that size. Bottom-up, each particle is read once: leaves from their particles, internal nodes from
// their children's fp64 sums (children are contiguous in the level-major array), one launch per level from the
// deepest up. The summation shape still depends only on the tree. Measured (B200, fp32 Plummer): 4 M particles
// 0.218 ms vs 0.165 ms top-down (22 small launches), 32 M particles 0.59 ms vs 1.25 ms: selected by size.
template <typename F>
__global__ void __launch_bounds__(256)
    props_leaf_kernel(const vec4<F> *__restrict__ p, const u64 *__restrict__ codes, const uint4 *__restrict__ nodeB,
                      vec4<F> *__restrict__ nodeA, F *__restrict__ node_delta, double *__restrict__ sums, u32 n_nodes,
                      int mac, level_dims<F> ld, u64 *__restrict__ err, u32 *__restrict__ big_list,
                      u32 *__restrict__ big_count)
{
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_nodes) {
        return;
    }
    const uint4 nb = nodeB[k];
    if ((nb.w & 0xffu) != 0u) {
        return; // internal node: summed from its children by props_level_kernel
    }
    if (nb.y - nb.x > PROPS_SMALL) {
        big_list[atomicAdd(big_count, 1u)] = k; // a leaf at the maximum depth can hold any number of particles
        return;
    }
    dsum4 s{0, 0, 0, 0};
    for (u32 i = nb.x; i < nb.y; ++i) {
        const vec4<F> v = p[i];
        dsum_add_particle(s, v.x, v.y, v.z, v.w);
    }
    *reinterpret_cast<double4 *>(sums + size_t(k) * 4) = make_double4(s.m, s.x, s.y, s.z);
    node_finalize<F>(s, k, nb.x, nb.w >> 8, codes, nodeA, node_delta, mac, ld, err);
}

template <typename F>
__global__ void __launch_bounds__(256)
    props_bigleaf_kernel(const vec4<F> *__restrict__ p, const u64 *__restrict__ codes, const double *__restrict__ chunks,
                         const double *__restrict__ chunks2, const uint4 *__restrict__ nodeB,
                         vec4<F> *__restrict__ nodeA, F *__restrict__ node_delta, double *__restrict__ sums, int mac,
                         level_dims<F> ld, u64 *__restrict__ err, const u32 *__restrict__ big_list,
                         const u32 *__restrict__ big_count)
{
    const u32 nbig = *big_count;
    const int lane = threadIdx.x & 31;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nbig; q += nwarps) {
        const u32 k = big_list[q];
        const uint4 nb = nodeB[k];
        const dsum4 s = node_sum<F, 32>(p, chunks, chunks2, nb.x, nb.y, lane, true);
        if (lane == 0) {
            *reinterpret_cast<double4 *>(sums + size_t(k) * 4) = make_double4(s.m, s.x, s.y, s.z);
            node_finalize<F>(s, k, nb.x, nb.w >> 8, codes, nodeA, node_delta, mac, ld, err);
        }
    }
}

// Internal nodes of one level, [k0, k1): children first_child .. first_child + nch - 1 live in the next level.
template <typename F>
__global__ void __launch_bounds__(256)
    props_level_kernel(const u64 *__restrict__ codes, const uint4 *__restrict__ nodeB, vec4<F> *__restrict__ nodeA,
                       F *__restrict__ node_delta, double *__restrict__ sums, u32 k0, u32 k1, int mac, level_dims<F> ld,
                       u64 *__restrict__ err)
{
    const u32 k = k0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k1) {
        return;
    }
    const uint4 nb = nodeB[k];
    const u32 nch = nb.w & 0xffu;
    if (nch == 0u) {
        return;
    }
    dsum4 s{0, 0, 0, 0};
    for (u32 j = 0; j < nch; ++j) {
        const double4 cs = *reinterpret_cast<const double4 *>(sums + size_t(nb.z + j) * 4);
        s.m += cs.x;
        s.x += cs.y;
        s.y += cs.z;
        s.z += cs.w;
    }
    *reinterpret_cast<double4 *>(sums + size_t(k) * 4) = make_double4(s.m, s.x, s.y, s.z);
    node_finalize<F>(s, k, nb.x, nb.w >> 8, codes, nodeA, node_delta, mac, ld, err);
}

// ---------------------------------------------------------------------------------------------------
// export to the reference's host layout
// ---------------------------------------------------------------------------------------------------
template <typename F>
struct host_node {
    u64 begin, end, n_children, code, level;
    F props[4], dim, delta;
};

template <typename F>
__global__ void __launch_bounds__(256)
    export_nodes_kernel(const u64 *__restrict__ codes, const uint4 *__restrict__ nodeB, const vec4<F> *__restrict__ nodeA,
                        const F *__restrict__ node_delta, const u32 *__restrict__ node_dfs,
                        const u32 *__restrict__ node_ndesc, u32 n_nodes, int mac, level_dims<F> ld,
                        host_node<F> *__restrict__ out)
{
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_nodes) {
        return;
    }
    const uint4 nb = nodeB[k];
    const vec4<F> a = nodeA[k];
    const u32 level = nb.w >> 8;
    host_node<F> h;
    h.begin = nb.x;
    h.end = nb.y;
    h.n_children = node_ndesc[k];
    h.level = level;
    h.code = (1ull << (3 * level)) | (level ? (codes[nb.x] >> (3 * (CBITS - static_cast<int>(level)))) : 0ull);
    h.props[0] = a.x;
    h.props[1] = a.y;
    h.props[2] = a.z;
    h.props[3] = a.w;
    if (mac == 0) {
        h.dim = rn_mul(ld.dim[level], ld.dim[level]); // dim2 = node_dim * node_dim, tree.hpp:1212
        h.delta = F(0);
    } else {
        h.dim = ld.dim[level];
        h.delta = node_delta[k];
    }
    out[node_dfs[k]] = h;
}

__global__ void __launch_bounds__(256)
    export_crit_kernel(const u64 *__restrict__ codes, const uint4 *__restrict__ nodeB, const u32 *__restrict__ crit_node,
                       const u32 *__restrict__ crit_begin, u32 n_crit, u64 *__restrict__ out)
{
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_crit) {
        return;
    }
    const uint4 nb = nodeB[crit_node[j]];
    const u32 level = nb.w >> 8;
    out[3 * size_t(j) + 0]
        = (1ull << (3 * level)) | (level ? (codes[nb.x] >> (3 * (CBITS - static_cast<int>(level)))) : 0ull);
    out[3 * size_t(j) + 1] = crit_begin[j];
    out[3 * size_t(j) + 2] = crit_begin[j + 1];
}

template <typename F>
level_dims<F> make_level_dims(F box)
{
    level_dims<F> ld;
    for (int l = 0; l < NLEVELS; ++l) {
        ld.dim[l] = box / static_cast<F>(u64(1) << l);
    }
    ld.cell = box * (F(1) / static_cast<F>(u64(1) << CBITS));
    ld.half_box = box * (F(1) / F(2));
    return ld;
}

// Sliding-window level: lvl[i] = min(21, 1 + max_{j in [i, i+w]} P'(j)); `a` holds P' on entry.
void window_levels(i8 *a, i8 *bbuf, i8 *lvl, size_t n, size_t w, cudaStream_t st)
{
    const unsigned g = div_up(n, 256);
    if (w >= n) {
        fill_i8_kernel<<<g, 256, 0, st>>>(lvl, n, i8(0)); count_launch();
        return;
    }
    size_t k = 1;
    i8 *in = a, *out = bbuf;
    while (k * 2 <= w + 1) {
        window_double_kernel<<<g, 256, 0, st>>>(in, out, n, k); count_launch();
        i8 *t = in;
        in = out;
        out = t;
        k *= 2;
    }
    window_final_kernel<<<g, 256, 0, st>>>(in, lvl, n, w + 1 - k); count_launch();
}

// Bucket of a sample sort: the number of splitters <= code (splitters ascending, at most 255 of them).
__global__ void __launch_bounds__(256) bucket_ids_kernel(const u64 *__restrict__ codes, size_t n,
                                                         const u64 *__restrict__ splitters, unsigned nsplit,
                                                         u64 *__restrict__ ids)
{
    __shared__ u64 sp[256];
    if (threadIdx.x < nsplit) {
        sp[threadIdx.x] = splitters[threadIdx.x];
    }
    __syncthreads();
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i < n) {
        const u64 c = codes[i];
        unsigned lo = 0, hi = nsplit; // first splitter > c
        while (lo < hi) {
            const unsigned mid = (lo + hi) / 2;
            if (sp[mid] <= c) {
                lo = mid + 1;
            } else {
                hi = mid;
            }
        }
        ids[i] = lo;
    }
}
__global__ void __launch_bounds__(256) gather_u64_kernel(const u64 *__restrict__ in, const u32 *__restrict__ idx,
                                                         u64 *__restrict__ out, size_t n)
{
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i < n) {
        out[i] = in[idx[i]];
    }
}

// One read, up to 8 writes: the same 8-byte words go to every destination (peer memory of the other ranks, reached
// with plain stores over NVLink). Used where the SMs have nothing else to do - the codes of the sorted buckets must be
// everywhere before the topology can start - and the copy engines, which serve 7 peers with fewer engines than
// peers, measured 2.6 ms for what the links carry in 1.
struct bcast_dst {
    u64 *p[8];
};
__device__ __forceinline__ void multimem_store_u64(u64 *mc, u64 v)
{
    asm volatile("multimem.st.relaxed.sys.global.u64 [%0], %1;" ::"l"(mc), "l"(v) : "memory");
}
// d.p[0] is an NVSwitch multicast address: one store per word leaves the GPU and the switch updates every rank's copy
__global__ void __launch_bounds__(256) mcast_copy_kernel(const u64 *__restrict__ src, u64 *mc, size_t n8)
{
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    for (; i + 3 * stride < n8; i += 4 * stride) {
        const u64 v0 = __ldcs(src + i), v1 = __ldcs(src + i + stride), v2 = __ldcs(src + i + 2 * stride),
                  v3 = __ldcs(src + i + 3 * stride);
        multimem_store_u64(mc + i, v0);
        multimem_store_u64(mc + i + stride, v1);
        multimem_store_u64(mc + i + 2 * stride, v2);
        multimem_store_u64(mc + i + 3 * stride, v3);
    }
    for (; i < n8; i += stride) {
        multimem_store_u64(mc + i, __ldcs(src + i));
    }
}
__global__ void __launch_bounds__(256) bcast_copy_kernel(const u64 *__restrict__ src, bcast_dst d, int nd, size_t n8,
                                                         const unsigned char *src_tail, size_t tail, size_t tail_off)
{
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    for (; i + 3 * stride < n8; i += 4 * stride) { // four loads in flight per thread, then 4 x nd stores
        const u64 v0 = __ldcs(src + i), v1 = __ldcs(src + i + stride), v2 = __ldcs(src + i + 2 * stride),
                  v3 = __ldcs(src + i + 3 * stride);
#pragma unroll 1
        for (int k = 0; k < nd; ++k) {
            u64 *q = d.p[k] + i;
            __stcs(q, v0);
            __stcs(q + stride, v1);
            __stcs(q + 2 * stride, v2);
            __stcs(q + 3 * stride, v3);
        }
    }
    for (; i < n8; i += stride) {
        const u64 v = __ldcs(src + i);
#pragma unroll 1
        for (int k = 0; k < nd; ++k) {
            __stcs(d.p[k] + i, v);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < tail) { // the last bytes % 8
        for (int k = 0; k < nd; ++k) {
            reinterpret_cast<unsigned char *>(d.p[k])[tail_off + threadIdx.x] = src_tail[threadIdx.x];
        }
    }
}

// out[perm[i]] = in[i]: results from Morton order to the original particle order (tree.hpp:3320-3330).
template <typename F>
__global__ void __launch_bounds__(256)
    scatter_perm_kernel(const F *__restrict__ in, const u32 *__restrict__ perm, F *__restrict__ out, size_t n)
{
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i < n) {
        out[perm[i]] = in[i];
    }
}

__device__ __forceinline__ u64 mix64(u64 x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}
__global__ void __launch_bounds__(256) digest_kernel(const u32 *__restrict__ w, size_t n, u64 *__restrict__ out)
{
    u64 h = 0;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        h += mix64((static_cast<u64>(i) << 32) ^ mix64(w[i]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        h += __shfl_xor_sync(0xffffffffu, h, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, h);
    }
}
__global__ void lower_bound_kernel(const u32 *__restrict__ arr, size_t n, const u64 *__restrict__ x, size_t k,
                                   u64 *__restrict__ out)
{
    const size_t j = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (j >= k) {
        return;
    }
    size_t lo = 0, hi = n; // first i with arr[i] >= x[j]
    while (lo < hi) {
        const size_t mid = lo + (hi - lo) / 2;
        if (arr[mid] < x[j]) {
            lo = mid + 1;
        } else {
            hi = mid;
        }
    }
    out[j] = lo;
}

} // namespace

// ---------------------------------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------------------------------
void launch_bucket_ids(const u64 *codes, size_t n, const u64 *splitters, unsigned nsplit, u64 *ids, cudaStream_t st)
{
    if (n) {
        bucket_ids_kernel<<<div_up(n, 256), 256, 0, st>>>(codes, n, splitters, nsplit, ids); count_launch();
    }
}
void launch_bcast_copy(void *const *dst, int nd, const void *src, size_t bytes, int sm_count, cudaStream_t st, bool multicast)
{
    if (!bytes || nd <= 0) {
        return;
    }
    if (multicast) { // (bytes % 8 == 0 checked by the caller)
        const size_t n8 = bytes / 8;
        const unsigned grid = static_cast<unsigned>(std::min<size_t>(size_t(sm_count) * 4, (n8 + 255) / 256));
        mcast_copy_kernel<<<grid, 256, 0, st>>>(static_cast<const u64 *>(src), static_cast<u64 *>(dst[0]), n8);
        count_launch();
        return;
    }
    bcast_dst d{};
    for (int k = 0; k < nd; ++k) {
        d.p[k] = static_cast<u64 *>(dst[k]);
    }
    const size_t n8 = bytes / 8, tail = bytes % 8;
    const unsigned grid = static_cast<unsigned>(std::min<size_t>(size_t(sm_count) * 8, (n8 + 255) / 256 + 1));
    bcast_copy_kernel<<<grid, 256, 0, st>>>(static_cast<const u64 *>(src), d, nd, n8,
                                            static_cast<const unsigned char *>(src) + n8 * 8, tail, n8 * 8);
    count_launch();
}
void launch_gather_u64(const u64 *in, const u32 *idx, u64 *out, size_t n, cudaStream_t st)
{
    if (n) {
        gather_u64_kernel<<<div_up(n, 256), 256, 0, st>>>(in, idx, out, n); count_launch();
    }
}
template <typename F>
void launch_scatter_perm(const F *in, const u32 *perm, F *out, size_t n, cudaStream_t st)
{
    if (n) {
        scatter_perm_kernel<F><<<div_up(n, 256), 256, 0, st>>>(in, perm, out, n); count_launch();
    }
}
template void launch_scatter_perm<float>(const float *, const u32 *, float *, size_t, cudaStream_t);
template void launch_scatter_perm<double>(const double *, const u32 *, double *, size_t, cudaStream_t);
void launch_digest(const void *a, size_t bytes, u64 *out, cudaStream_t st)
{
    if (bytes >= 4) {
        digest_kernel<<<148 * 8, 256, 0, st>>>(static_cast<const u32 *>(a), bytes / 4, out); count_launch();
    }
}
void launch_lower_bound(const u32 *arr, size_t n, const u64 *x, size_t k, u64 *out, cudaStream_t st)
{
    if (k) {
        lower_bound_kernel<<<div_up(k, 64), 64, 0, st>>>(arr, n, x, k, out); count_launch();
    }
}
constexpr unsigned STREAM_GRID = 148 * 8; // grid-stride kernels: a multiple of the SM count

template <typename F>
void launch_pack_absmax(const F *x, const F *y, const F *z, const F *m, vec4<F> *out, size_t n, u64 *absmax_bits,
                        cudaStream_t st)
{
    if (n) {
        pack_absmax_kernel<F><<<STREAM_GRID, 256, 0, st>>>(x, y, z, m, out, n, absmax_bits); count_launch();
    }
}
template <typename F>
void launch_set_coords(vec4<F> *p, const F *x, const F *y, const F *z, const F *m, size_t n, u64 *absmax_bits,
                       cudaStream_t st)
{
    if (n) {
        set_coords_kernel<F><<<STREAM_GRID, 256, 0, st>>>(p, x, y, z, m, n, absmax_bits); count_launch();
    }
}
template <typename F>
void launch_unpack(const vec4<F> *in, F *x, F *y, F *z, F *m, size_t n, cudaStream_t st)
{
    if (n) {
        unpack_kernel<F><<<div_up(n, 256), 256, 0, st>>>(in, x, y, z, m, n); count_launch();
    }
}
template <typename F>
void launch_encode(const vec4<F> *p, u64 *codes, size_t n, F inv_box, dev_error *err, cudaStream_t st)
{
    if (n) {
        encode_kernel<F><<<div_up(n, 256), 256, 0, st>>>(p, codes, n, inv_box, reinterpret_cast<u64 *>(err)); count_launch();
    }
}
template <typename F>
void launch_gather(const vec4<F> *pin, const u32 *idx, vec4<F> *pout, size_t n, cudaStream_t st, const F *late_m)
{
    if (n) {
        gather_kernel<F><<<div_up(n, 256), 256, 0, st>>>(pin, idx, pout, n, late_m); count_launch();
    }
}
void launch_perm_invert(const u32 *perm, u32 *inv_perm, size_t n, cudaStream_t st)
{
    if (n) {
        perm_invert_kernel<<<div_up(n, 256), 256, 0, st>>>(perm, inv_perm, n); count_launch();
    }
}
void launch_perm_first(const u32 *last_perm, u32 *perm, u32 *inv_perm, size_t n, cudaStream_t st)
{
    if (n) {
        perm_first_kernel<<<div_up(n, 256), 256, 0, st>>>(last_perm, perm, inv_perm, n); count_launch();
    }
}
void launch_perm_compose(const u32 *old_perm, const u32 *last_perm, u32 *new_perm, u32 *inv_perm, size_t n,
                         cudaStream_t st)
{
    if (n) {
        perm_compose_kernel<<<div_up(n, 256), 256, 0, st>>>(old_perm, last_perm, new_perm, inv_perm, n); count_launch();
    }
}

template <typename F>
void topology_count(build_arrays<F> &b, size_t max_leaf_n, size_t ncrit, cudaStream_t st)
{
    const size_t n = b.n;
    const u32 ntiles = div_up(n, TOPO_TILE);
    b.delta.reserve(n, 1.05);
    b.lvl_leaf.reserve(n, 1.05);
    b.lvl_crit.reserve(n, 1.05);
    b.win_a.reserve(2 * n, 1.05); // two P' arrays
    b.win_b.reserve(n, 1.05);
    b.dfsbase.reserve(n + 1, 1.05);
    b.tilecnt.reserve(size_t(NLEVELS + 1) * ntiles, 1.25);
    b.rowtot.reserve(NLEVELS + 1);
    const size_t w1 = max_leaf_n, w2 = ncrit > max_leaf_n ? ncrit : max_leaf_n;
    if (w2 <= size_t(WIN_MAXW) && w2 >= 1 && w1 < n) {
        // common case: everything in one shared-memory tiled kernel
        const size_t smem = size_t(WIN_TILE + 2 * w2) * 8 + 2 * (size_t(WIN_TILE + w2 + 3) / 4 + 1) * 4;
        RK_CUDA_CHECK(cudaFuncSetAttribute(window_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
        // windows wider than the particle count give level 0 everywhere: clamp (same result, bounded halo)
        const int cw2 = static_cast<int>(w2 < n ? w2 : n), cw1 = static_cast<int>(w1);
        window_fused_kernel<<<div_up(n, WIN_TILE), 256, smem, st>>>(b.codes, n, cw1, cw2 < cw1 ? cw1 : cw2, b.delta.p,
                                                                  b.lvl_leaf.p, b.lvl_crit.p); count_launch();
    } else {
        i8 *p1 = b.win_a.p, *p2 = (w2 != w1) ? b.win_a.p + n : nullptr;
        delta_window_kernel<<<div_up(n, 256), 256, 0, st>>>(b.codes, n, w1, w2, b.delta.p, p1, p2); count_launch();
        window_levels(p1, b.win_b.p, b.lvl_leaf.p, n, w1, st);
        if (p2) {
            window_levels(p2, b.win_b.p, b.lvl_crit.p, n, w2, st);
        } else {
            RK_CUDA_CHECK(cudaMemcpyAsync(b.lvl_crit.p, b.lvl_leaf.p, n, cudaMemcpyDeviceToDevice, st));
        }
    }
    topo_count_kernel<<<ntiles, TOPO_THREADS, 0, st>>>(b.delta.p, b.lvl_leaf.p, b.lvl_crit.p, n, ntiles, b.tilecnt.p); count_launch();
    row_scan_wide_kernel<<<NLEVELS + 1, 1024, 0, st>>>(b.tilecnt.p, ntiles, b.rowtot.p); count_launch();
    RK_CUDA_CHECK(cudaGetLastError());
}

template <typename F>
void topology_emit(build_arrays<F> &b, cudaStream_t st)
{
    const size_t n = b.n;
    const u32 ntiles = div_up(n, TOPO_TILE);
    const u32 M = static_cast<u32>(b.n_nodes), C = static_cast<u32>(b.n_crit);
    topo_emit_kernel<<<ntiles, TOPO_THREADS, 0, st>>>(b.delta.p, b.lvl_leaf.p, b.lvl_crit.p, n, ntiles, b.tilecnt.p,
                                                   b.levels, b.nodeB.p, b.node_dfs.p, b.dfsbase.p, b.crit_node.p,
                                                   b.crit_begin.p, M, C); count_launch();
    topo_finalize_kernel<<<div_up(M, 256), 256, 0, st>>>(b.codes, n, b.nodeB.p, b.node_dfs.p, b.dfsbase.p,
                                                         b.node_ndesc.p, b.levels, M); count_launch();
    topo_children_kernel<<<div_up(M, 256), 256, 0, st>>>(b.nodeB.p, b.levels, M); count_launch();
    RK_CUDA_CHECK(cudaMemsetAsync(b.d_misc.p + 2, 0, sizeof(u32), st));
    max_group_kernel<<<div_up(C, 256), 256, 0, st>>>(b.crit_begin.p, C, b.d_misc.p + 2); count_launch();
    RK_CUDA_CHECK(cudaGetLastError());
}

template <typename F>
void node_properties(build_arrays<F> &b, int mac, F box_size, cudaStream_t st)
{
    const size_t n = b.n;
    const u32 M = static_cast<u32>(b.n_nodes);
    if (!M) {
        return;
    }
    const u32 nchunks = div_up(n, PROPS_CHUNK);
    const u32 nsuper = div_up(nchunks, PROPS_CHUNK);
    b.chunksum.reserve((size_t(nchunks) + nsuper) * 4, 1.05);
    double *chunks2 = b.chunksum.p + size_t(nchunks) * 4;
    chunk_sums_kernel<F><<<div_up(size_t(nchunks) * 32, 256), 256, 0, st>>>(b.psorted.p, n, nchunks, b.chunksum.p); count_launch();
    chunk_sums2_kernel<<<div_up(size_t(nsuper) * 32, 256), 256, 0, st>>>(b.chunksum.p, nchunks, nsuper, chunks2); count_launch();
    // big-node queue lives in the (now free) window scratch: M u32 entries + the counter in d_misc[3]
    b.win_b.reserve(size_t(M) * 4 + 16, 1.05);
    u32 *big_list = reinterpret_cast<u32 *>(b.win_b.p);
    u32 *big_count = b.d_misc.p + 3;
    RK_CUDA_CHECK(cudaMemsetAsync(big_count, 0, sizeof(u32), st));
    const level_dims<F> ld = make_level_dims<F>(box_size);
    u64 *err = reinterpret_cast<u64 *>(b.d_err.p) + 1;
    const bool bottom_up = b.props_bottom_up < 0 ? n >= PROPS_BOTTOMUP_MIN : b.props_bottom_up != 0;
    if (bottom_up) {
        b.nodesum.reserve(size_t(M) * 4, 1.1);
        props_leaf_kernel<F><<<div_up(M, 256), 256, 0, st>>>(b.psorted.p, b.codes, b.nodeB.p, b.nodeA.p, b.node_delta.p,
                                                             b.nodesum.p, M, mac, ld, err, big_list, big_count); count_launch();
        props_bigleaf_kernel<F><<<148 * 4, 256, 0, st>>>(b.psorted.p, b.codes, b.chunksum.p, chunks2, b.nodeB.p, b.nodeA.p,
                                                        b.node_delta.p, b.nodesum.p, mac, ld, err, big_list, big_count); count_launch();
        for (int l = NLEVELS - 1; l >= 0; --l) {
            const u32 k0 = b.levels.base[l], k1 = b.levels.base[l + 1];
            if (k1 > k0) {
                props_level_kernel<F><<<div_up(k1 - k0, 256), 256, 0, st>>>(b.codes, b.nodeB.p, b.nodeA.p, b.node_delta.p,
                                                                            b.nodesum.p, k0, k1, mac, ld, err); count_launch();
            }
        }
        RK_CUDA_CHECK(cudaGetLastError());
        return;
    }
    node_props_small_kernel<F><<<div_up(size_t(M) * PROPS_LANES, 256), 256, 0, st>>>(
        b.psorted.p, b.codes, b.chunksum.p, b.nodeB.p, b.nodeA.p, b.node_delta.p, M, mac, ld, err, big_list, big_count); count_launch();
    node_props_big_kernel<F><<<148 * 4, 256, 0, st>>>(b.psorted.p, b.codes, b.chunksum.p, chunks2, b.nodeB.p, b.nodeA.p,
                                                     b.node_delta.p, mac, ld, err, big_list, big_count); count_launch();
    RK_CUDA_CHECK(cudaGetLastError());
}

template <typename F>
void launch_export_nodes(const build_arrays<F> &b, int mac, F box_size, void *d_out, cudaStream_t st)
{
    const u32 M = static_cast<u32>(b.n_nodes);
    if (M) {
        export_nodes_kernel<F><<<div_up(M, 256), 256, 0, st>>>(b.codes, b.nodeB.p, b.nodeA.p, b.node_delta.p,
                                                               b.node_dfs.p, b.node_ndesc.p, M, mac,
                                                               make_level_dims<F>(box_size),
                                                               static_cast<host_node<F> *>(d_out)); count_launch();
    }
}

void launch_export_crit(const u64 *codes, const uint4 *nodeB, const u32 *crit_node, const u32 *crit_begin, size_t n_crit,
                        u64 *d_out, cudaStream_t st)
{
    if (n_crit) {
        export_crit_kernel<<<div_up(n_crit, 256), 256, 0, st>>>(codes, nodeB, crit_node, crit_begin,
                                                                static_cast<u32>(n_crit), d_out); count_launch();
    }
}

#define RK_INSTANTIATE(F)                                                                                              \
    template void launch_pack_absmax<F>(const F *, const F *, const F *, const F *, vec4<F> *, size_t, u64 *,          \
                                        cudaStream_t);                                                                 \
    template void launch_set_coords<F>(vec4<F> *, const F *, const F *, const F *, const F *, size_t, u64 *,           \
                                       cudaStream_t);                                                                  \
    template void launch_unpack<F>(const vec4<F> *, F *, F *, F *, F *, size_t, cudaStream_t);                         \
    template void launch_encode<F>(const vec4<F> *, u64 *, size_t, F, dev_error *, cudaStream_t);                      \
    template void launch_gather<F>(const vec4<F> *, const u32 *, vec4<F> *, size_t, cudaStream_t, const F *);          \
    template void topology_count<F>(build_arrays<F> &, size_t, size_t, cudaStream_t);                                  \
    template void topology_emit<F>(build_arrays<F> &, cudaStream_t);                                                   \
    template void node_properties<F>(build_arrays<F> &, int, F, cudaStream_t);                                         \
    template void launch_export_nodes<F>(const build_arrays<F> &, int, F, void *, cudaStream_t);

RK_INSTANTIATE(float)
RK_INSTANTIATE(double)

} // namespace rk
