// traverse.cu — ncrit-grouped Barnes-Hut traversal for sm_100a (FP32/FP64-pipe bound; no tensor cores:
// this is not a dense contraction).
//
// Algorithm = the reference's CPU path (include/rakau/tree.hpp): tree_acc_pot 2798-2849,
// tree_acc_pot_mac_check 2597-2793, tree_acc_pot_src_com 2477-2590, tree_acc_pot_leaf 2327-2471,
// tree_self_interactions 2073-2321, G scaling + write-out 2986-3007 — NOT the reference's per-particle
// CUDA kernel (src/rakau_cuda.cu:152-335).
//
// Mapping: one warp owns one critical node (target group, <= ncrit particles). The group's targets live in
// registers (R per lane) and, for the MAC test, in shared memory. The warp walks the level-major node array
// with a shared-memory stack of (first child, count) entries: each step pops 32 nodes, ONE NODE PER LANE;
// the lane evaluates the reference's group MAC (accepted iff mac_lh < dist2 for EVERY target, dist2
// unsoftened, tree.hpp:2666-2672 / 2753) with exactly rounded operations so that decisions — and therefore
// interaction counts — match the oracle. Accepted nodes' (com, mass) and the particles of rejected leaves
// are appended (ballot + popc compaction; leaf particles staged with cp.async) to a shared-memory ring of
// float4/double4 sources. Whenever the ring holds >= 32 sources the warp evaluates them against its
// register-resident targets: broadcast LDS.128 per source, 3 FADD + 3 FFMA + MUFU.RSQ + 3 FMUL + 3 FFMA per
// pair. Accumulation order is fixed by the tree, not by scheduling, so results are run-to-run
// deterministic and G enters as one final multiply (reference test g_constant_acc.cpp:65-88).

#include "common.cuh"
#include "scan.cuh"

namespace rk
{

namespace
{

constexpr int TRAV_WARPS = 8;
constexpr int TRAV_THREADS = TRAV_WARPS * 32;
// targets per lane held in registers: 8 (fp32, 256 targets per pass) or 4 (fp64, register budget)
template <typename F>
struct trav_cfg {
    static constexpr int rmax = sizeof(F) == 8 ? 4 : 8;
};
constexpr int LCAP = 128;        // source ring capacity (power of two)
constexpr int STACK_CAP = 1024;  // (first child, count) entries per warp
constexpr u32 FULL = 0xffffffffu;

__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gsrc)
{
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_vec4(float4 *dst, const float4 *src) { cp_async_16(dst, src); }
__device__ __forceinline__ void cp_async_vec4(double4 *dst, const double4 *src)
{
    cp_async_16(dst, src);
    cp_async_16(reinterpret_cast<char *>(dst) + 16, reinterpret_cast<const char *>(src) + 16);
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ float fast_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double fast_rsqrt(double x) { return rsqrt(x); }

__device__ __forceinline__ u32 warp_incl_scan(u32 v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(FULL, v, o);
        if (lane >= o) {
            v += t;
        }
    }
    return v;
}

// One (target, source) interaction. Accelerations: acc += d * m_src / r^3. Potential: the sum of m_src / r
// is accumulated; the caller multiplies by -m_target (tree.hpp:2026-2041, 2466, 2586).
template <typename F, int Q>
__device__ __forceinline__ void interact(const vec4<F> &s, F tx, F ty, F tz, F eps2, F &ax, F &ay, F &az, F &ap)
{
    const F dx = s.x - tx, dy = s.y - ty, dz = s.z - tz;
    F d2 = fma(dx, dx, eps2);
    d2 = fma(dy, dy, d2);
    d2 = fma(dz, dz, d2);
    const F inv = fast_rsqrt(d2);
    if (Q != 1) {
        const F ms = s.w * (inv * inv * inv);
        ax = fma(dx, ms, ax);
        ay = fma(dy, ms, ay);
        az = fma(dz, ms, az);
    }
    if (Q != 0) {
        ap = fma(s.w, inv, ap);
    }
}

template <typename F, int Q, int RR, int RMAX>
__device__ __forceinline__ void eval_ring(const vec4<F> *__restrict__ ring, u32 head, u32 cnt, F eps2,
                                          const F (&tx)[RMAX], const F (&ty)[RMAX], const F (&tz)[RMAX],
                                          F (&ax)[RMAX], F (&ay)[RMAX], F (&az)[RMAX], F (&ap)[RMAX])
{
#pragma unroll 4
    for (u32 j = 0; j < cnt; ++j) {
        const vec4<F> s = ring[(head + j) & (LCAP - 1)];
#pragma unroll
        for (int k = 0; k < RR; ++k) {
            interact<F, Q>(s, tx[k], ty[k], tz[k], eps2, ax[k], ay[k], az[k], ap[k]);
        }
    }
}

// Self interactions: sources are the group's own particles, the (i, i) pair is masked out.
template <typename F, int Q, int RR, int RMAX>
__device__ __forceinline__ void eval_self(const vec4<F> *__restrict__ src, u32 T, u32 my0 /* t0 + lane */, F eps2,
                                          const F (&tx)[RMAX], const F (&ty)[RMAX], const F (&tz)[RMAX],
                                          F (&ax)[RMAX], F (&ay)[RMAX], F (&az)[RMAX], F (&ap)[RMAX])
{
#pragma unroll 2
    for (u32 j = 0; j < T; ++j) {
        const vec4<F> s = src[j];
#pragma unroll
        for (int k = 0; k < RR; ++k) {
            const F dx = s.x - tx[k], dy = s.y - ty[k], dz = s.z - tz[k];
            F d2 = fma(dx, dx, eps2);
            d2 = fma(dy, dy, d2);
            d2 = fma(dz, dz, d2);
            F inv = fast_rsqrt(d2);
            inv = (j == my0 + 32u * k) ? F(0) : inv;
            if (Q != 1) {
                const F ms = s.w * (inv * inv * inv);
                ax[k] = fma(dx, ms, ax[k]);
                ay[k] = fma(dy, ms, ay[k]);
                az[k] = fma(dz, ms, az[k]);
            }
            if (Q != 0) {
                ap[k] = fma(s.w, inv, ap[k]);
            }
        }
    }
}

#define RK_RR_CASE(N, CALL)                                                                                            \
    case N:                                                                                                            \
        if constexpr (N <= RMAX) {                                                                                     \
            CALL(N);                                                                                                   \
        }                                                                                                              \
        break;
#define RK_RR_SWITCH(rr, CALL)                                                                                         \
    switch (rr) {                                                                                                      \
        RK_RR_CASE(1, CALL)                                                                                            \
        RK_RR_CASE(2, CALL)                                                                                            \
        RK_RR_CASE(3, CALL)                                                                                            \
        RK_RR_CASE(4, CALL)                                                                                            \
        RK_RR_CASE(5, CALL)                                                                                            \
        RK_RR_CASE(6, CALL)                                                                                            \
        RK_RR_CASE(7, CALL)                                                                                            \
        RK_RR_CASE(8, CALL)                                                                                            \
        default: break;                                                                                                \
    }

template <typename F>
__host__ __device__ constexpr size_t warp_smem_bytes(u32 tmax)
{
    return size_t(LCAP) * sizeof(vec4<F>) + size_t(tmax) * sizeof(vec4<F>) + size_t(STACK_CAP) * 4 + 32 * 4 /*nodebuf*/
           + 32 * 4 /*lq_incl*/ + 32 * 4 /*lq_base*/;
}

template <typename F, int Q, int MAC>
__global__ void __launch_bounds__(TRAV_THREADS, 2) traverse_kernel(const trav_params<F> p)
{
    constexpr int RMAX = trav_cfg<F>::rmax;
    extern __shared__ __align__(32) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *base = smem_raw + size_t(warp) * warp_smem_bytes<F>(p.tmax);
    vec4<F> *ring = reinterpret_cast<vec4<F> *>(base);
    vec4<F> *tgt = ring + LCAP;
    u32 *stack = reinterpret_cast<u32 *>(tgt + p.tmax);
    u32 *nodebuf = stack + STACK_CAP;
    u32 *lq_incl = nodebuf + 32;
    u32 *lq_base = lq_incl + 32;
    const u32 ltm = lanemask_lt();
    const F eps2 = p.eps2;

    for (;;) {
        u32 g = 0;
        if (lane == 0) {
            g = atomicAdd(p.work_counter, 1u);
        }
        g = __shfl_sync(FULL, g, 0) + p.c0;
        if (g >= p.c1) {
            break;
        }
        const u32 gnode = p.crit_node[g], gb = p.crit_begin[g], ge = p.crit_begin[g + 1], T = ge - gb;
        const bool staged = T <= p.tmax;
        const vec4<F> *tsrc = staged ? tgt : (p.parts + gb);
        __syncwarp();
        if (staged) {
            for (u32 i = lane; i < T; i += 32) {
                tgt[i] = p.parts[gb + i];
            }
        }
        __syncwarp();

        u64 n_mac = 0, n_acc = 0, n_p2p = 0; // warp-uniform counters

        // Groups larger than 32*RMAX targets are handled in several passes, each repeating the traversal
        // (the MAC always spans the whole group, as in the reference).
        for (u32 t0 = 0; t0 < T; t0 += 32u * RMAX) {
            const u32 tc = (T - t0 < 32u * RMAX) ? (T - t0) : 32u * RMAX;
            const int rr = static_cast<int>((tc + 31u) / 32u);
            F tx[RMAX], ty[RMAX], tz[RMAX], tm[RMAX], ax[RMAX], ay[RMAX], az[RMAX], ap[RMAX];
#pragma unroll
            for (int k = 0; k < RMAX; ++k) {
                const u32 i = t0 + 32u * k + lane;
                vec4<F> v = make_vec4<F>(F(0), F(0), F(0), F(0));
                if (k < rr && i < T) {
                    v = tsrc[i];
                }
                tx[k] = v.x;
                ty[k] = v.y;
                tz[k] = v.z;
                tm[k] = v.w;
                ax[k] = ay[k] = az[k] = ap[k] = F(0);
            }

            u32 sp = 1, lhead = 0, lcount = 0, lq_total = 0, lq_done = 0;
            bool done = false, overflow = false;
            if (lane == 0) {
                stack[0] = 0u; // root: first = 0, count = 1
            }
            __syncwarp();

            for (;;) {
                // ---------------- produce: fill the ring until >= 32 sources or the walk is over -------------
                while (lcount < 32u && !done) {
                    if (lq_done < lq_total) {
                        // copy more particles of the rejected leaves into the ring
                        const u32 room = LCAP - lcount, rem = lq_total - lq_done;
                        const u32 chunk = rem < room ? rem : room;
                        for (u32 f = lq_done + lane; f < lq_done + chunk; f += 32) {
                            int lo = 0;
#pragma unroll
                            for (int s = 16; s > 0; s >>= 1) {
                                if (lq_incl[lo + s - 1] <= f) {
                                    lo += s;
                                }
                            }
                            const u32 pidx = lq_base[lo] + f;
                            cp_async_vec4(&ring[(lhead + lcount + (f - lq_done)) & (LCAP - 1)], p.parts + pidx);
                        }
                        cp_async_wait_all();
                        __syncwarp();
                        lcount += chunk;
                        lq_done += chunk;
                        continue;
                    }
                    if (sp == 0) {
                        done = true;
                        break;
                    }
                    // ---- pop up to 32 nodes ----
                    u32 ecnt = 0, efc = 0;
                    if (static_cast<u32>(lane) < sp) {
                        const u32 e = stack[sp - 1 - lane];
                        ecnt = (e & 7u) + 1u;
                        efc = e >> 3;
                    }
                    const u32 incl = warp_incl_scan(ecnt, lane), excl = incl - ecnt;
                    const bool take = ecnt && excl < 32u;
                    const u32 ntake = __popc(__ballot_sync(FULL, take));
                    bool partial = false;
                    if (take) {
                        const u32 used = (ecnt < 32u - excl) ? ecnt : (32u - excl);
                        for (u32 t = 0; t < used; ++t) {
                            nodebuf[excl + t] = efc + t;
                        }
                        if (used < ecnt) {
                            stack[sp - 1 - lane] = ((efc + used) << 3) | (ecnt - used - 1u);
                            partial = true;
                        }
                    }
                    const bool any_partial = __any_sync(FULL, partial);
                    u32 nnodes = __shfl_sync(FULL, incl, ntake - 1);
                    nnodes = nnodes < 32u ? nnodes : 32u;
                    sp -= ntake - (any_partial ? 1u : 0u);
                    __syncwarp();

                    // ---- one node per lane: classify ----
                    const bool have = static_cast<u32>(lane) < nnodes;
                    u32 k = 0;
                    uint4 nb = make_uint4(0, 0, 0, 0);
                    vec4<F> na = make_vec4<F>(F(0), F(0), F(0), F(0));
                    if (have) {
                        k = nodebuf[lane];
                        nb = p.nodeB[k];
                        na = p.nodeA[k];
                    }
                    const u32 nch = nb.w & 0xffu, level = nb.w >> 8;
                    const bool is_self = have && k == gnode;
                    const bool is_anc = have && !is_self && nb.x <= gb && ge <= nb.y;
                    const bool test = have && !is_self && !is_anc;
                    F mac_lh = F(0);
                    if (test) {
                        if (MAC == 0) {
                            mac_lh = p.mac_tab[level];
                        } else {
                            const F t = rn_fma(p.mac_tab[level], p.mac_value, p.node_delta[k]);
                            mac_lh = rn_mul(t, t);
                        }
                    }
                    // Group MAC, tree.hpp:2741-2759: fails as soon as one target has mac_lh >= dist2.
                    bool fail = !test;
                    for (u32 i = 0; i < T; ++i) {
                        const vec4<F> t = tsrc[i];
                        const F dx = rn_sub(na.x, t.x), dy = rn_sub(na.y, t.y), dz = rn_sub(na.z, t.z);
                        F d2 = rn_mul(dx, dx);
                        d2 = rn_fma(dy, dy, d2);
                        d2 = rn_fma(dz, dz, d2);
                        fail = fail || (mac_lh >= d2);
                        if ((i & 3u) == 3u && __all_sync(FULL, fail)) {
                            break;
                        }
                    }
                    const bool accept = test && !fail;
                    const bool open_leaf = test && fail && nch == 0u;
                    const bool descend = is_anc || (test && fail && nch != 0u);
                    const u32 m_test = __ballot_sync(FULL, test), m_acc = __ballot_sync(FULL, accept),
                              m_leaf = __ballot_sync(FULL, open_leaf), m_desc = __ballot_sync(FULL, descend);
                    n_mac += __popc(m_test);
                    n_acc += __popc(m_acc);
                    // accepted nodes -> ring
                    if (accept) {
                        ring[(lhead + lcount + __popc(m_acc & ltm)) & (LCAP - 1)] = na;
                    }
                    lcount += __popc(m_acc);
                    // rejected internal nodes / ancestors -> stack
                    if (m_desc) {
                        if (sp + 32u > STACK_CAP) {
                            overflow = true;
                        } else if (descend) {
                            stack[sp + __popc(m_desc & ltm)] = (nb.z << 3) | (nch - 1u);
                        }
                        sp += overflow ? 0u : __popc(m_desc);
                    }
                    // rejected leaves -> leaf queue
                    lq_total = 0;
                    lq_done = 0;
                    if (m_leaf) {
                        const u32 c = open_leaf ? (nb.y - nb.x) : 0u;
                        const u32 li = warp_incl_scan(c, lane);
                        lq_incl[lane] = li;
                        lq_base[lane] = nb.x - (li - c);
                        lq_total = __shfl_sync(FULL, li, 31);
                        n_p2p += lq_total;
                    }
                    __syncwarp();
                    if (overflow) {
                        done = true;
                    }
                }
                if (lcount == 0u) {
                    break;
                }
                // ---------------- consume: evaluate up to 32 sources (the only ring call site) ----------------
                const u32 ne = lcount < 32u ? lcount : 32u;
#define RK_CALL_RING(RR) eval_ring<F, Q, RR, RMAX>(ring, lhead, ne, eps2, tx, ty, tz, ax, ay, az, ap)
                RK_RR_SWITCH(rr, RK_CALL_RING)
#undef RK_CALL_RING
                __syncwarp();
                lhead = (lhead + ne) & (LCAP - 1);
                lcount -= ne;
            }
            if (overflow && lane == 0) {
                atomicExch(p.err, 1u);
            }

            // self interactions inside the group, tree.hpp:2073-2321
#define RK_CALL_SELF(RR) eval_self<F, Q, RR, RMAX>(tsrc, T, t0 + lane, eps2, tx, ty, tz, ax, ay, az, ap)
            RK_RR_SWITCH(rr, RK_CALL_SELF)
#undef RK_CALL_SELF

            // G scaling (one final multiply, tree.hpp:2986-3002) and write-out (3004-3007)
#pragma unroll
            for (int kk = 0; kk < RMAX; ++kk) {
                const u32 i = t0 + 32u * kk + lane;
                if (kk < rr && i < T) {
                    u32 dst = gb + i;
                    if (p.perm) {
                        dst = p.perm[dst];
                    }
                    dst -= p.out_offset;
                    if (Q == 0 || Q == 2) {
                        p.out[0][dst] = ax[kk] * p.G;
                        p.out[1][dst] = ay[kk] * p.G;
                        p.out[2][dst] = az[kk] * p.G;
                    }
                    if (Q == 1) {
                        p.out[0][dst] = (-tm[kk] * ap[kk]) * p.G;
                    }
                    if (Q == 2) {
                        p.out[3][dst] = (-tm[kk] * ap[kk]) * p.G;
                    }
                }
            }
            if (t0 == 0 && lane == 0) {
                if (p.group_cost) {
                    p.group_cost[g] = u64(T) * (n_p2p + n_acc + u64(T) - 1u);
                }
                if (p.counters) {
                    atomicAdd(p.counters + 0, n_mac);
                    atomicAdd(p.counters + 1, n_acc);
                    atomicAdd(p.counters + 2, n_p2p * T);
                    atomicAdd(p.counters + 3, u64(T) * (u64(T) - 1u) / 2u);
                    atomicAdd(p.counters + 4, n_acc * T);
                }
            }
            n_mac = n_acc = n_p2p = 0; // count the first pass only
            __syncwarp();
        }
    }
}

// Direct summation for one particle (exact_acc_pot_impl, tree.hpp:3531-3569): one CTA, double accumulation
// of F-precision pair terms, out4 = ax, ay, az, pot (already multiplied by G).
template <typename F>
__global__ void __launch_bounds__(1024) exact_kernel(const vec4<F> *__restrict__ parts, size_t n, size_t idx, F G, F eps2,
                                                     double *__restrict__ out4)
{
    __shared__ double red[4][32];
    const vec4<F> me = parts[idx];
    double a0 = 0, a1 = 0, a2 = 0, pt = 0;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
        if (i == idx) {
            continue;
        }
        const vec4<F> s = parts[i];
        const F dx = s.x - me.x, dy = s.y - me.y, dz = s.z - me.z;
        F d2 = fma(dx, dx, eps2);
        d2 = fma(dy, dy, d2);
        d2 = fma(dz, dz, d2);
        const F inv = F(1) / sqrt(d2), gm = G * s.w * inv, gm3 = inv * inv * gm;
        a0 += double(dx * gm3);
        a1 += double(dy * gm3);
        a2 += double(dz * gm3);
        pt += double(-gm * me.w);
    }
    double v[4] = {a0, a1, a2, pt};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v[q] += __shfl_xor_sync(FULL, v[q], o);
        }
        if ((threadIdx.x & 31) == 0) {
            red[q][threadIdx.x >> 5] = v[q];
        }
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0;
        for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) {
            s += red[threadIdx.x][w];
        }
        out4[threadIdx.x] = s;
    }
}

// FP32-pipe peak probe: 8 independent FFMA chains per thread, full occupancy.
__global__ void __launch_bounds__(256) ffma_kernel(float *out, int iters, float a, float b)
{
    float v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            v0 = fmaf(v0, a, b);
            v1 = fmaf(v1, a, b);
            v2 = fmaf(v2, a, b);
            v3 = fmaf(v3, a, b);
            v4 = fmaf(v4, a, b);
            v5 = fmaf(v5, a, b);
            v6 = fmaf(v6, a, b);
            v7 = fmaf(v7, a, b);
        }
    }
    const float s = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
    if (s == 123.456f) {
        out[0] = s;
    }
}

template <typename F, int Q, int MAC>
void launch_one(const trav_params<F> &p, int sm_count, cudaStream_t st)
{
    const size_t smem = warp_smem_bytes<F>(p.tmax) * TRAV_WARPS;
    RK_CUDA_CHECK(cudaFuncSetAttribute(traverse_kernel<F, Q, MAC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    int per_sm = 0;
    RK_CUDA_CHECK(
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, traverse_kernel<F, Q, MAC>, TRAV_THREADS, smem));
    if (per_sm < 1) {
        per_sm = 1;
    }
    const u32 ngroups = p.c1 - p.c0;
    u32 grid = static_cast<u32>(sm_count) * static_cast<u32>(per_sm); // persistent: a multiple of the SM count
    const u32 need = (ngroups + TRAV_WARPS - 1) / TRAV_WARPS;
    if (grid > need) {
        grid = need;
    }
    if (grid == 0) {
        return;
    }
    traverse_kernel<F, Q, MAC><<<grid, TRAV_THREADS, smem, st>>>(p); count_launch();
    RK_CUDA_CHECK(cudaGetLastError());
}

} // namespace

template <typename F>
void launch_traverse(const trav_params<F> &p, int Q, int mac, int sm_count, cudaStream_t st)
{
#define RK_DISPATCH(QQ, MM)                                                                                            \
    if (Q == QQ && mac == MM) {                                                                                        \
        launch_one<F, QQ, MM>(p, sm_count, st);                                                                        \
        return;                                                                                                        \
    }
    RK_DISPATCH(0, 0)
    RK_DISPATCH(1, 0)
    RK_DISPATCH(2, 0)
    RK_DISPATCH(0, 1)
    RK_DISPATCH(1, 1)
    RK_DISPATCH(2, 1)
#undef RK_DISPATCH
    throw cuda_error(1, "invalid Q / MAC combination");
}

double ffma_microbench(float *ms)
{
    int dev = 0, sms = 0;
    RK_CUDA_CHECK(cudaGetDevice(&dev));
    RK_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float *d = nullptr;
    RK_CUDA_CHECK(cudaMalloc(&d, 64));
    cudaEvent_t e0, e1;
    RK_CUDA_CHECK(cudaEventCreate(&e0));
    RK_CUDA_CHECK(cudaEventCreate(&e1));
    const int iters = 4096, grid = sms * 8;
    ffma_kernel<<<grid, 256>>>(d, 64, 1.0001f, 0.5f); // warm-up
    RK_CUDA_CHECK(cudaEventRecord(e0));
    ffma_kernel<<<grid, 256>>>(d, iters, 1.0001f, 0.5f);
    RK_CUDA_CHECK(cudaEventRecord(e1));
    RK_CUDA_CHECK(cudaEventSynchronize(e1));
    RK_CUDA_CHECK(cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return 2.0 * 8 * 16 * double(iters) * 256.0 * grid;
}

template <typename F>
void launch_exact(const vec4<F> *parts, size_t n, size_t idx, F G, F eps2, double *d_out4, cudaStream_t st)
{
    exact_kernel<F><<<1, 1024, 0, st>>>(parts, n, idx, G, eps2, d_out4); count_launch();
    RK_CUDA_CHECK(cudaGetLastError());
}

template void launch_traverse<float>(const trav_params<float> &, int, int, int, cudaStream_t);
template void launch_traverse<double>(const trav_params<double> &, int, int, int, cudaStream_t);
template void launch_exact<float>(const vec4<float> *, size_t, size_t, float, float, double *, cudaStream_t);
template void launch_exact<double>(const vec4<double> *, size_t, size_t, double, double, double *, cudaStream_t);

} // namespace rk
