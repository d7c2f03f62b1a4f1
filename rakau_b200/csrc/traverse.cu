// traverse.cu — ncrit-grouped Barnes-Hut traversal for sm_100a (FP32/FP64-pipe bound; no tensor cores:
// this is not a dense contraction).
//
// Algorithm = the reference's CPU path (include/rakau/tree.hpp): tree_acc_pot 2798-2849,
// tree_acc_pot_mac_check 2597-2793, tree_acc_pot_src_com 2477-2590, tree_acc_pot_leaf 2327-2471,
// tree_self_interactions 2073-2321, G scaling + write-out 2986-3007 — NOT the reference's per-particle
// CUDA kernel (src/rakau_cuda.cu:152-335).
//
// Mapping: one warp owns one critical node (target group, <= ncrit particles), persistent CTAs of 4 warps take
// groups from an atomic counter. The group's targets are staged in shared memory; their accumulators live in
// shared memory between batches and in registers (tiles of <= 4 targets per lane) inside a batch. The warp walks
// the level-major node array with a shared-memory stack of (first child, count) entries: each step pops up to 4
// entries = up to 32 nodes, ONE NODE PER LANE; the lane evaluates the reference's group MAC (accepted iff
// mac_lh < dist2 for EVERY target, dist2 unsoftened, tree.hpp:2666-2672 / 2753) so that decisions - and therefore
// interaction counts - match the oracle exactly. Accepted nodes' (com, mass) and the particles of rejected leaves
// are appended (ballot + popc compaction; leaf particles staged with cp.async) to a shared-memory ring of
// float4/double4 sources; whenever the ring holds a batch (64 sources) the warp evaluates it against its targets.
// Accumulation order is fixed by the tree, not by scheduling, so results are run-to-run deterministic and G
// enters as one final multiply (reference test g_constant_acc.cpp:65-88).
//
// What the ncu profiles under profiles/ led to (each step measured on the 4M Plummer workload, kernel time
// 12.1 ms for the first version -> 4.84 ms):
//  (1) MAC: the per-target loop was 26 % of all issued instructions. Each lane first lower-bounds dist2 with the
//      group's bounding box (80 % of the tests end there: accepted), then tests ONE target with the reference's
//      exactly rounded arithmetic - the group's support point towards the node, i.e. (almost always) its nearest
//      target: if that one fails the node is rejected exactly as in the reference (18 %). Only the remaining 2 %
//      run the exact loop over all targets, two nodes per pass (one per half-warp).
//  (2) Lane utilisation: groups average 38 targets, so a warp is cut into S = 32/P slices of P lanes (P = 4..32
//      chosen per group): every slice holds all targets (rr = ceil(T/P) per lane) and evaluates a contiguous
//      1/S of each batch; partial sums are combined with a fixed shuffle tree at the end (slot padding 37 % -> 8 %).
//  (3) Instruction cache: eight unrolled register-resident variants made a 136 KB kernel that spent 5.6 issue
//      slots per instruction waiting on fetch; accumulators moved to shared memory, register tiles of 4/2/1.
//  (4) Issue slots: fp32 interactions use the sm_100a packed instructions FFMA2/FADD2/FMUL2 (two targets per
//      register pair, sources through the broadcast operand form), halving the FP32 instructions issued; slices
//      are chosen so that the slot count per lane is even whenever that costs no padding.
//  (5) Batches of 64 sources (was 32) amortise the accumulator round trip; 5 CTAs/SM (96 registers). Round 2: 128
//      sources at 4 CTAs/SM (128 registers), see launch_one().
// Measured dead ends are recorded in DESIGN.md (next-step node prefetch, per-leaf copy loops, 6 CTAs/SM,
// 8-target register tiles).

#include "common.cuh"
#include "scan.cuh"

#include <cstdio>
#include <cstdlib>

namespace rk
{

namespace
{

constexpr int TRAV_WARPS = 4;
constexpr int TRAV_THREADS = TRAV_WARPS * 32;
#ifndef RK_RING_RESET
#define RK_RING_RESET 1
#endif
#ifndef RK_RING_RESET_F64
#define RK_RING_RESET_F64 0 // fp64 (scalar register tiles): the circular 64-source batches measured faster (36.3 vs 39.0 ms, config 3)
#endif
#ifndef RK_RING_GROW
#define RK_RING_GROW 1
#endif
#ifndef RK_RING_ROOM
#define RK_RING_ROOM 32 // 32: the scratch block aliases the step's append area; 64: it lies behind it
#endif
#ifndef RK_UNROLL
#define RK_UNROLL 4
#endif
#ifndef RK_AMB2
#define RK_AMB2 1
#endif
#ifndef RK_PACKED
#define RK_PACKED 1
#endif
#ifndef RK_NP4
#define RK_NP4 0
#endif
#ifndef RK_SINGLE_SCALAR
#define RK_SINGLE_SCALAR 1
#endif
#ifndef RK_ACC15
#define RK_ACC15 0
#endif
#ifndef RK_SKIP_EVAL
#define RK_SKIP_EVAL 0
#endif
#ifndef RK_CTAS
#define RK_CTAS 4
#endif
#ifndef RK_STEAL
#define RK_STEAL 1
#endif
#ifndef RK_F64_CTAS
#define RK_F64_CTAS 3
#endif
#ifndef RK_BATCH_BIG
#define RK_BATCH_BIG 128
#endif
#define RK_PRAGMA_(x) _Pragma(#x)
#define RK_UNROLL_PRAGMA(n) RK_PRAGMA_(unroll n)
// BATCH (template parameter of the kernel) = sources evaluated per consume step; the source ring holds 2 * BATCH
// entries: the batch being filled + room for one step's appends / scratch. Larger batches amortise the per-batch
// accumulator round trip through shared memory and the tile prologues: 64 was 6 % faster than 32, 128 (at 4 CTAs/SM,
// which also frees 128 registers per thread) another 2.6 %; 32 is kept for the large tmax configurations where the
// bigger ring would cost a resident CTA (chosen in launch_one()).
#ifndef RK_STACK
#define RK_STACK 512
#endif
constexpr int STACK_CAP = RK_STACK;   // (first child, count) entries per warp (<= 32 pushes per step, depth <= 21)
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
__device__ __forceinline__ void multimem_store(float *mc, float v)
{
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc), "f"(v) : "memory");
}
__device__ __forceinline__ void multimem_store(double *mc, double v)
{
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc), "d"(v) : "memory");
}

// MUFU.RSQ without the denormal pre/post-scaling of rsqrtf() (a squared distance below 1e-38 is not a
// meaningful input; it flushes to zero exactly like a coincident pair).
__device__ __forceinline__ float fast_rsqrt(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// fp64: the library rsqrt() (MUFU.RSQ64H + refinement). Measured dead ends on B200 (4 M particles, theta = 0.5, 36.3 ms
// with this): seeding two fp64 Newton steps from the fp32 MUFU.RSQ (two F2F conversions per pair: 64.5 ms) or from
// rsqrt.approx.ftz.f64 (60.1 ms).
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

// Evaluate the sources src[jb, je) (this slice's contiguous share of the batch) against a register tile of W targets whose
// positions come from the staged target array and whose accumulators live in shared memory between batches
// (one float4/double4 per (slot, lane): conflict-free LDS.128/STS.128). Keeping only W <= 4 slots in registers
// keeps the kernel small enough for the instruction cache: the first version unrolled 8 register-resident
// variants and spent 5.6 issue slots per instruction waiting on instruction fetch (profiles/r01_traverse_v2_*).
template <typename F, int Q, int W, bool SELF>
__device__ __forceinline__ void eval_tile(const vec4<F> *__restrict__ src, u32 jb, u32 je, F eps2,
                                          const vec4<F> *__restrict__ tpos, u32 T, u32 first_t, u32 P,
                                          vec4<F> *__restrict__ acc)
{
    F tx[W], ty[W], tz[W], ax[W], ay[W], az[W], ap[W];
    u32 self_idx[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        const u32 ti = first_t + P * w;
        const vec4<F> t = tpos[ti < T ? ti : T - 1u];
        const vec4<F> a = acc[32 * w];
        tx[w] = t.x;
        ty[w] = t.y;
        tz[w] = t.z;
        ax[w] = a.x;
        ay[w] = a.y;
        az[w] = a.z;
        ap[w] = a.w;
        self_idx[w] = ti;
    }
RK_UNROLL_PRAGMA(RK_UNROLL)
    for (u32 j = jb; j < je; ++j) {
        const vec4<F> s = src[j];
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const F dx = s.x - tx[w], dy = s.y - ty[w], dz = s.z - tz[w];
            F d2 = fma(dx, dx, eps2);
            d2 = fma(dy, dy, d2);
            d2 = fma(dz, dz, d2);
            F inv = fast_rsqrt(d2);
            if (SELF) {
                inv = (j == self_idx[w]) ? F(0) : inv; // the (i, i) pair
            }
            if (Q != 1) {
                const F ms = s.w * (inv * inv * inv);
                ax[w] = fma(dx, ms, ax[w]);
                ay[w] = fma(dy, ms, ay[w]);
                az[w] = fma(dz, ms, az[w]);
            }
            if (Q != 0) {
                ap[w] = fma(s.w, inv, ap[w]);
            }
        }
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
        acc[32 * w] = make_vec4<F>(ax[w], ay[w], az[w], ap[w]);
    }
}

// ---- packed FP32 (sm_100a FFMA2 / FADD2 / FMUL2, PTX *.f32x2) ------------------------------------------------
// One packed instruction performs two independent IEEE fp32 operations on a 64-bit register pair for ONE issue
// slot (measured, tools/microbench/ffma2.cu: FFMA2 sustains the same 72 TFLOP/s as FFMA with half the issued
// instructions). The traversal kernel is bound by issue slots (57 % FP32, 43 % walk/addressing/control), so the
// fp32 interaction loop pairs two TARGETS of the lane in one register pair; the source components are scalar
// registers which ptxas feeds through the instruction's broadcast operand form (FADD2 Rd, -Rt.F32x2, Rs.F32), so
// the loop body is 1 LDS.128 + 12 packed instructions + 2 MUFU.RSQ per two interactions and no MOV.
// Every lane-level operation is the same rn operation in the same order as the scalar loop: results are
// bit-identical.
__device__ __forceinline__ u64 pk2(float lo, float hi)
{
    u64 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void upk2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b)
{
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b)
{
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// NP pairs of target slots (slots 2k and 2k+1 of the tile; SINGLE: one slot, paired with itself and the upper half
// discarded) against the sources src[jb, je).
template <int Q, int NP, bool SINGLE, bool SELF>
__device__ __forceinline__ void eval_tile_packed(const float4 *__restrict__ src, u32 jb, u32 je, float eps2,
                                                 const float4 *__restrict__ tpos, u32 T, u32 first_t, u32 P,
                                                 float4 *__restrict__ acc)
{
    u64 tx[NP], ty[NP], tz[NP], ax[NP], ay[NP], az[NP], ap[NP];
    u32 self0[NP], self1[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const u32 t0i = first_t + P * (2 * k), t1i = SINGLE ? t0i : first_t + P * (2 * k + 1);
        self0[k] = t0i;
        self1[k] = t1i;
        // component-wise scalar loads straight into the two halves of each register pair (a float4 load would
        // leave the halves in different quads and ptxas re-packs them with MOVs inside the loop)
        const float *p0 = reinterpret_cast<const float *>(tpos + (t0i < T ? t0i : T - 1u)),
                    *p1 = reinterpret_cast<const float *>(tpos + (t1i < T ? t1i : T - 1u));
        const float4 a0 = acc[32 * (2 * k)], a1 = SINGLE ? a0 : acc[32 * (2 * k + 1)];
        tx[k] = pk2(p0[0], p1[0]);
        ty[k] = pk2(p0[1], p1[1]);
        tz[k] = pk2(p0[2], p1[2]);
        ax[k] = pk2(a0.x, a1.x);
        ay[k] = pk2(a0.y, a1.y);
        az[k] = pk2(a0.z, a1.z);
        ap[k] = pk2(a0.w, a1.w);
    }
    const u64 e2 = pk2(eps2, eps2);
RK_UNROLL_PRAGMA(RK_UNROLL)
    for (u32 j = jb; j < je; ++j) {
        const float4 s = src[j];
        const u64 sx = pk2(s.x, s.x), sy = pk2(s.y, s.y), sz = pk2(s.z, s.z), sm = pk2(s.w, s.w);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const u64 dx = sub2(sx, tx[k]), dy = sub2(sy, ty[k]), dz = sub2(sz, tz[k]);
            u64 d2 = fma2(dx, dx, e2);
            d2 = fma2(dy, dy, d2);
            d2 = fma2(dz, dz, d2);
            float d2a, d2b;
            upk2(d2, d2a, d2b);
            float ia = fast_rsqrt(d2a), ib = fast_rsqrt(d2b);
            if (SELF) { // the (i, i) pair
                ia = (j == self0[k]) ? 0.f : ia;
                ib = (j == self1[k]) ? 0.f : ib;
            }
            const u64 inv = pk2(ia, ib);
            if (Q != 1) {
                const u64 ms = mul2(sm, mul2(mul2(inv, inv), inv));
                ax[k] = fma2(dx, ms, ax[k]);
                ay[k] = fma2(dy, ms, ay[k]);
                az[k] = fma2(dz, ms, az[k]);
            }
            if (Q != 0) {
                ap[k] = fma2(sm, inv, ap[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        float4 a0, a1;
        upk2(ax[k], a0.x, a1.x);
        upk2(ay[k], a0.y, a1.y);
        upk2(az[k], a0.z, a1.z);
        upk2(ap[k], a0.w, a1.w);
        acc[32 * (2 * k)] = a0;
        if (!SINGLE) {
            acc[32 * (2 * k + 1)] = a1;
        }
    }
}

// fp32, non-self sources: all rr slots of this lane against the ring entries, two slots per register pair.
template <int Q, bool SELF>
__device__ __forceinline__ void eval_slots_packed(const float4 *__restrict__ src, u32 cnt, u32 sl, u32 ls, float eps2,
                                                  const float4 *__restrict__ tpos, u32 T, u32 t_lane, u32 P, u32 rr,
                                                  float4 *__restrict__ acc_lane)
{
    const u32 per = (cnt + (1u << ls) - 1u) >> ls, jb = sl * per < cnt ? sl * per : cnt, je = jb + per < cnt ? jb + per : cnt;
    u32 k = 0;
#if RK_NP4
#pragma unroll 1
    for (; k + 8u <= rr; k += 8u) {
        eval_tile_packed<Q, 4, false, SELF>(src, jb, je, eps2, tpos, T, t_lane + P * k, P, acc_lane + 32u * k);
    }
    if (k + 4u <= rr) {
        eval_tile_packed<Q, 2, false, SELF>(src, jb, je, eps2, tpos, T, t_lane + P * k, P, acc_lane + 32u * k);
        k += 4u;
    }
#else
#pragma unroll 1
    for (; k + 4u <= rr; k += 4u) {
        eval_tile_packed<Q, 2, false, SELF>(src, jb, je, eps2, tpos, T, t_lane + P * k, P, acc_lane + 32u * k);
    }
#endif
    if (k + 2u <= rr) {
        eval_tile_packed<Q, 1, false, SELF>(src, jb, je, eps2, tpos, T, t_lane + P * k, P, acc_lane + 32u * k);
        k += 2u;
    }
    if (k < rr) {
#if RK_SINGLE_SCALAR
        eval_tile<float, Q, 1, SELF>(src, jb, je, eps2, tpos, T, t_lane + P * k, P, acc_lane + 32u * k);
#else
        eval_tile_packed<Q, 1, true, SELF>(src, jb, je, eps2, tpos, T, t_lane + P * k, P, acc_lane + 32u * k);
#endif
    }
}

// All rr slots of this lane against the same sources: tiles of 4, 2, 1 slots.
template <typename F, int Q, bool SELF>
__device__ __forceinline__ void eval_slots(const vec4<F> *__restrict__ src, u32 cnt, u32 sl, u32 ls, F eps2,
                                           const vec4<F> *__restrict__ tpos, u32 T, u32 t_lane, u32 P, u32 rr,
                                           vec4<F> *__restrict__ acc_lane)
{
    // slice sl takes the contiguous chunk [sl * per, (sl + 1) * per) of the cnt sources
    const u32 per = (cnt + (1u << ls) - 1u) >> ls, jb = sl * per < cnt ? sl * per : cnt, je = jb + per < cnt ? jb + per : cnt;
    u32 k = 0;
#pragma unroll 1
    for (; k + 4u <= rr; k += 4u) {
        eval_tile<F, Q, 4, SELF>(src, jb, je, eps2, tpos, T, t_lane + P * k, P, acc_lane + 32u * k);
    }
    if (k + 2u <= rr) {
        eval_tile<F, Q, 2, SELF>(src, jb, je, eps2, tpos, T, t_lane + P * k, P, acc_lane + 32u * k);
        k += 2u;
    }
    if (k < rr) {
        eval_tile<F, Q, 1, SELF>(src, jb, je, eps2, tpos, T, t_lane + P * k, P, acc_lane + 32u * k);
    }
}

__host__ __device__ constexpr u32 acc_entries(u32 tmax)
{
    // accumulator entries per warp = 32 * (slots per lane). Slicing replicates the accumulators S times, so more
    // entries let more groups use narrow slices without padding: 2 * tmax where it costs no resident CTA.
#if RK_ACC15
    return tmax + tmax / 2u;
#else
    return tmax <= 128u ? 2u * tmax : tmax + tmax / 2u;
#endif
}
template <typename F>
__host__ __device__ constexpr size_t warp_smem_bytes(u32 tmax, u32 LCAP)
{
    // ring + staged targets + accumulators + stack + queues
    return size_t(LCAP) * sizeof(vec4<F>) + size_t(tmax) * sizeof(vec4<F>) + size_t(acc_entries(tmax)) * sizeof(vec4<F>)
           + size_t(STACK_CAP) * 4 + 32 * 4 /*lq_incl*/ + 32 * 4 /*lq_base*/ + 16 * 4 /*run state*/;
}

// First index j in [a, b) with arr[j] >= x, else b; the 32 lanes probe 32 positions per round (monotone array).
__device__ __forceinline__ u32 warp_lower_bound(const u32 *__restrict__ arr, u32 a, u32 b, u32 x, int lane)
{
    u32 lo = a, n = b - a; // the answer lies in [lo, lo + n]
    while (n > 0u) {
        const u32 step = (n + 32u) / 33u, idx = lo + (static_cast<u32>(lane) + 1u) * step - 1u;
        const bool less = idx < lo + n && arr[idx] < x;
        const u32 c = __popc(__ballot_sync(FULL, less)); // probes below x form a prefix of the lanes
        // lane c's probe (if it exists) is the first one known to be >= x
        const u32 nlo = lo + c * step, nhi = lo + (c + 1u) * step - 1u;
        n = (c < 32u && nhi < lo + n ? nhi : lo + n) - nlo;
        lo = nlo;
    }
    return lo;
}

// Work units. p.window == 0: one critical node (group) per unit, walked from the root (the round-1 scheme).
// p.window == W > 0: TWO-PHASE walk of a RUN of sibling groups - the critical nodes whose first particle lies in
// one W-aligned window of the Morton-sorted particles (~7 groups, W .. W + tmax - 1 targets):
//   phase 1, once per run: walk from the root with the bounding box of ALL the run's targets. A node whose MAC
//     holds for the whole box is accepted by every group of the run (tree.hpp:2666-2672 holds for each target), a
//     node whose MAC fails even for the farthest corner of the box is rejected by every group (each of them has a
//     failing target): those two classes - and the ancestors of the whole run - are resolved ONCE, their sources
//     are evaluated against all the run's targets at full lane utilisation, and the partial sums are parked in the
//     output arrays. Every other node (undecided by the box, or overlapping the run) goes to the run's FRONTIER;
//   phase 2, per group: the reference's group walk with the exact group MAC, started from the frontier nodes
//     instead of the root; the group's sums are added to the parked partial sums.
// A group never sees a descendant of a node it accepts, and every node test has the reference's outcome, so the
// per-group decisions and the interaction counts are exactly those of tree_acc_pot (tests T1); only the order of
// the additions differs. On the 4M Plummer tree this removes 47 % of the node visits and 43 % of the interactions
// come from phase 1 (tests/studies/two_phase_walk_study.py).
// (fp64: the double4 rings and accumulators let 3 CTAs fit in shared memory, so the kernel may use 168 registers)
template <typename F, int Q, int MAC, int BATCH_>
__global__ void __launch_bounds__(TRAV_THREADS, sizeof(F) == 8 ? RK_F64_CTAS : RK_CTAS) traverse_kernel(const trav_params<F> p)
{
    constexpr u32 BATCH = BATCH_, LCAP = 2 * BATCH_;
    // Ring discipline. RESET (rings of >= 128 entries): the walk fills the ring up to TH = LCAP - 32 entries (one step
    // appends <= 32 more; the step's scratch block is the same 32 entries, used before the step's appends are written),
    // the consume step evaluates EVERYTHING that is there and the ring starts again at entry 0 - batches of 224+ sources
    // instead of 128 for the same shared memory, i.e. 40 % fewer tile prologues / accumulator round trips. Otherwise
    // (the 64-entry ring of the large-tmax configurations): batches of exactly BATCH sources from a circular ring, the
    // scratch block BATCH entries after its head.
    // In RESET mode nothing needs a power of two, so the ring takes whatever shared memory the resident CTAs leave
    // unused (p.ring entries >= LCAP, chosen in launch_one()).
    constexpr bool RESET = RK_RING_RESET && BATCH_ >= 64 && (sizeof(F) == 4 || RK_RING_RESET_F64);
    const u32 lcap = RESET ? p.ring : LCAP;
    const u32 TH = RESET ? lcap - RK_RING_ROOM : BATCH;
    auto ridx = [&](u32 i) { return RESET ? i : (i & (LCAP - 1)); };
    extern __shared__ __align__(32) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *base = smem_raw + size_t(warp) * warp_smem_bytes<F>(p.tmax, lcap);
    vec4<F> *ring = reinterpret_cast<vec4<F> *>(base);
    vec4<F> *tgt = ring + lcap;
    vec4<F> *acc = tgt + p.tmax;
    u32 *stack = reinterpret_cast<u32 *>(acc + acc_entries(p.tmax));
    const u32 rr_cap = acc_entries(p.tmax) / 32u;              // accumulator slots per lane (one group)
    const u32 rr_cap1 = (p.tmax + acc_entries(p.tmax)) / 32u; // phase 1: the staged-target area holds accumulators too
    // the run's frontier lives at the top of the stack array and grows downwards: front(i) = stack_top[-i]
    u32 *stack_top = stack + (STACK_CAP - 1);
    u32 *lq_incl = stack + STACK_CAP;
    u32 *lq_base = lq_incl + 32;
    // state of the run this warp is attached to: kept in shared memory, it is only needed between stages
    // (registers are what limits the evaluation loop)
    u32 *rs = lq_base + 32;
    enum { RS_J0, RS_J1, RS_E0, RS_NG, RS_SLOT, RS_FCOUNT, RS_MAC1, RS_ACC1, RS_P2P1, RS_PARTIAL, RS_NEED_PH1, RS_GI };
    const u32 ltm = lanemask_lt();
    const F eps2 = p.eps2;
    const u32 W = p.window;
    const u32 win0 = W ? p.crit_begin[p.c0] / W : 0u;
    const u32 n_units = W ? (p.crit_begin[p.c1] - 1u) / W - win0 + 1u : p.c1 - p.c0;

    // Tail of the launch: a run is ~7 groups, a warp gets only ~5 runs of a 4M-particle evaluation, so the last runs
    // would leave most warps idle while a few finish (9 % of the kernel time). The runs of the LAST WAVE (the final
    // steal_k units of the queue) therefore publish their phase-1 state (frontier, counters; the partial sums are
    // already in the output arrays) in global memory and hand out their groups through an atomic counter: a warp that
    // finds the queue empty attaches to a published run and takes groups from it. Who evaluates a group does not
    // change a single operation of its evaluation, so the results stay bit-identical.
    constexpr u32 NONE = 0xffffffffu;
    const u32 ksteal = (RK_STEAL && p.steal && W) ? (p.steal_k < n_units ? p.steal_k : n_units) : 0u;
    bool queue_empty = false;
    // every warp scans the published runs from its own position (no herd on the first run with work left)
    u32 scan_pos = (blockIdx.x * u32(TRAV_WARPS) + u32(warp)) * 97u, backoff = 1000u;

    for (;;) {
        // ---- attach to a run: the next unit of the queue, or a published run of the last wave ----
        // [j0, j1): the run (ALL the groups of the window, so that phase 1 - and with it every result bit - does not
        // depend on how the critical nodes are cut into launches or ranks); e0 .. e0 + ng - 1: its groups in [c0, c1)
        u32 slot = NONE;
        if (!queue_empty) {
            u32 u = 0;
            if (lane == 0) {
                u = atomicAdd(p.work_counter, 1u);
            }
            u = __shfl_sync(FULL, u, 0);
            if (u >= n_units) {
                queue_empty = true;
            } else {
                u32 j0 = p.c0 + u, j1 = j0 + 1u, e0 = j0, ng = 1u;
                if (W) {
                    const u32 lo = (win0 + u) * W;
                    j0 = warp_lower_bound(p.crit_begin, 0u, p.ncrit, lo, lane);
                    const u32 lim = p.ncrit - j0 < W ? p.ncrit : j0 + W; // a window holds at most W groups
                    j1 = warp_lower_bound(p.crit_begin, j0, lim, lo + W, lane);
                    e0 = j0 > p.c0 ? j0 : p.c0;
                    const u32 e1 = j1 < p.c1 ? j1 : p.c1;
                    ng = e1 > e0 ? e1 - e0 : 0u;
                }
                // phase 1 is skipped for a single group: its walk then starts at the root
                const bool need_ph1 = j1 - j0 > 1u && ng != 0u;
                if (n_units - 1u - u < ksteal) {
                    if (need_ph1) {
                        slot = n_units - 1u - u; // published after phase 1
                    } else if (lane == 0) {
                        atomicAdd(p.steal_published, 1u); // nothing to share
                    }
                }
                if (ng == 0u) {
                    continue;
                }
                __syncwarp();
                if (lane == 0) {
                    stack_top[0] = 0u; // frontier = {root} unless phase 1 fills it
                    rs[RS_J0] = j0;
                    rs[RS_J1] = j1;
                    rs[RS_E0] = e0;
                    rs[RS_NG] = ng;
                    rs[RS_SLOT] = slot;
                    rs[RS_FCOUNT] = 1u;
                    rs[RS_MAC1] = 0u; // phase-1 tests / accepted nodes / leaf particles: shared by all the groups
                    rs[RS_ACC1] = 0u;
                    rs[RS_P2P1] = 0u;
                    rs[RS_PARTIAL] = 0u;
                    rs[RS_NEED_PH1] = need_ph1 ? 1u : 0u;
                    rs[RS_GI] = 0u;
                }
                __syncwarp();
            }
        }
        if (queue_empty) {
            if (ksteal == 0u) {
                break;
            }
            for (;;) {
                const u32 pub = *static_cast<volatile u32 *>(p.steal_published); // read BEFORE the scan
                for (u32 sb = 0; sb < ksteal && slot == NONE; sb += 32u) {
                    const u32 si = (scan_pos + sb + static_cast<u32>(lane)) % ksteal;
                    bool ok = false;
                    if (sb + static_cast<u32>(lane) < ksteal) {
                        const volatile u32 *h = p.steal + size_t(si) * 16u;
                        ok = h[2] != 0u && h[0] < h[1];
                    }
                    const u32 m = __ballot_sync(FULL, ok);
                    if (m) {
                        slot = __shfl_sync(FULL, si, __ffs(m) - 1);
                    }
                }
                if (slot != NONE) {
                    scan_pos = slot;
                    backoff = 1000u;
                    break;
                }
                if (pub == ksteal) {
                    break; // every run of the last wave was resolved before the scan and none has work left
                }
                __nanosleep(backoff);
                backoff = backoff < 8000u ? backoff * 2u : backoff;
            }
            if (slot == NONE) {
                break;
            }
            __threadfence();
            const volatile u32 *h = p.steal + size_t(slot) * 16u;
            const u32 fc = h[3];
            __syncwarp();
            if (lane == 0) {
                rs[RS_NG] = h[1];
                rs[RS_FCOUNT] = fc;
                rs[RS_MAC1] = h[4];
                rs[RS_ACC1] = h[5];
                rs[RS_P2P1] = h[6];
                rs[RS_E0] = h[7];
                rs[RS_PARTIAL] = h[8];
                rs[RS_SLOT] = slot;
                rs[RS_NEED_PH1] = 0u;
                rs[RS_GI] = 0u;
            }
            for (u32 i = lane; i < fc; i += 32u) {
                *(stack_top - i) = p.steal_front[size_t(slot) * STACK_CAP + i];
            }
            __syncwarp();
        }
        for (;;) {
        const bool ph1 = rs[RS_NEED_PH1] != 0u;
        u32 g = rs[RS_J0];
        __syncwarp();
        if (ph1) {
            if (lane == 0) {
                rs[RS_NEED_PH1] = 0u;
            }
        } else {
            u32 gi = 0;
            if (lane == 0) {
                const u32 sl_ = rs[RS_SLOT];
                gi = sl_ != NONE ? atomicAdd(p.steal + size_t(sl_) * 16u, 1u) : rs[RS_GI]++;
            }
            gi = __shfl_sync(FULL, gi, 0);
            if (gi >= rs[RS_NG]) {
                break;
            }
            g = rs[RS_E0] + gi;
        }
        const u32 j1 = rs[RS_J1];
        const u32 gnode = ph1 ? 0xffffffffu : p.crit_node[g], gb = p.crit_begin[g], ge = p.crit_begin[ph1 ? j1 : g + 1u],
                  T = ge - gb;
        const bool staged = !ph1 && T <= p.tmax;
        const vec4<F> *gsrc = p.parts + gb;
        __syncwarp();
        // Stage the targets (MAC test + self interactions) and compute the group's bounding box.
        // ... and, for the quick rejection test of the MAC, the group's (approximate) support points along the 8
        // diagonal directions: cand_pos byte q = index of the target maximising (+x, sy*y, sz*z), q = 2*(sy<0)+(sz<0);
        // cand_neg byte q = the same for (-x, sy*y, sz*z).
        F blo[3], bhi[3];
        u32 cand_pos, cand_neg;
        {
            u32 kmax[4] = {0u, 0u, 0u, 0u}, kmin[4] = {0u, 0u, 0u, 0u};
            F lo0 = F(INFINITY), lo1 = F(INFINITY), lo2 = F(INFINITY), hi0 = -F(INFINITY), hi1 = -F(INFINITY),
              hi2 = -F(INFINITY);
            for (u32 i = lane; i < T; i += 32) {
                const vec4<F> v = gsrc[i];
                if (staged) {
                    tgt[i] = v;
                }
                if (i < 256u && !ph1) {
                    // keys = (order-preserving bits of the functional, 8 low bits replaced by the target index)
                    const float fx = static_cast<float>(v.x), fy = static_cast<float>(v.y), fz = static_cast<float>(v.z);
                    const float sk[4] = {(fx + fy) + fz, (fx + fy) - fz, (fx - fy) + fz, (fx - fy) - fz};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const u32 b = __float_as_uint(sk[q]), uu = b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
                        const u32 k1 = (uu & 0xffffff00u) | i, k2 = (~uu & 0xffffff00u) | i;
                        kmax[q] = k1 > kmax[q] ? k1 : kmax[q];
                        kmin[q] = k2 > kmin[q] ? k2 : kmin[q];
                    }
                }
                lo0 = fmin(lo0, v.x);
                hi0 = fmax(hi0, v.x);
                lo1 = fmin(lo1, v.y);
                hi1 = fmax(hi1, v.y);
                lo2 = fmin(lo2, v.z);
                hi2 = fmax(hi2, v.z);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo0 = fmin(lo0, __shfl_xor_sync(FULL, lo0, o));
                lo1 = fmin(lo1, __shfl_xor_sync(FULL, lo1, o));
                lo2 = fmin(lo2, __shfl_xor_sync(FULL, lo2, o));
                hi0 = fmax(hi0, __shfl_xor_sync(FULL, hi0, o));
                hi1 = fmax(hi1, __shfl_xor_sync(FULL, hi1, o));
                hi2 = fmax(hi2, __shfl_xor_sync(FULL, hi2, o));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                kmax[q] = __reduce_max_sync(FULL, kmax[q]) & 0xffu;
                kmin[q] = __reduce_max_sync(FULL, kmin[q]) & 0xffu;
            }
            cand_pos = kmax[0] | (kmax[1] << 8) | (kmax[2] << 16) | (kmax[3] << 24);
            cand_neg = kmin[3] | (kmin[2] << 8) | (kmin[1] << 16) | (kmin[0] << 24);
            blo[0] = lo0;
            blo[1] = lo1;
            blo[2] = lo2;
            bhi[0] = hi0;
            bhi[1] = hi1;
            bhi[2] = hi2;
        }
        __syncwarp();

        u32 n_mac = 0, n_acc = 0, n_p2p = 0; // warp-uniform counters (a group visits far fewer than 2^32 nodes)

        // Groups with more targets than the accumulator array holds are handled in several passes, each
        // repeating the traversal (the MAC always spans the whole group, as in the reference).
        const vec4<F> *tpos = staged ? tgt : gsrc;
        const F bmid[3] = {(blo[0] + bhi[0]) * F(0.5), (blo[1] + bhi[1]) * F(0.5), (blo[2] + bhi[2]) * F(0.5)};
        const u32 cap = ph1 ? rr_cap1 : rr_cap;
        // phase 1 writes partial sums for the targets of this launch only
        const u32 fb = ph1 ? p.crit_begin[rs[RS_E0]] : 0u, fe = ph1 ? p.crit_begin[rs[RS_E0] + rs[RS_NG]] : 0u;
        vec4<F> *acc_lane = (ph1 ? tgt : acc) + lane; // phase 1 reads its targets from global memory (L1)
        for (u32 t0 = 0; t0 < T; t0 += 32u * cap) {
            const u32 tc = (T - t0 < 32u * cap) ? (T - t0) : 32u * cap;
            // Slice the warp: P lanes per slice, S = 32/P slices, rr target slots per lane; minimise rr * P >= tc.
            // Cost of a choice = slots + half a slot per lane when rr is odd: the fp32 loop evaluates two slots per
            // packed instruction, an unpaired last slot runs the scalar loop at twice the issue cost.
            u32 lp = 5u, rr = (tc + 31u) / 32u; // P = 1 << lp
            u32 best = (2u * rr + (rr & 1u)) << 5;
#pragma unroll
            for (u32 l = 4u; l >= 2u; --l) {
                const u32 r = (tc + (1u << l) - 1u) >> l, cost = (2u * r + (r & 1u)) << l;
                if (r <= cap && cost < best) {
                    best = cost;
                    lp = l;
                    rr = r;
                }
            }
            const u32 P = 1u << lp, sl = static_cast<u32>(lane) >> lp, tl = static_cast<u32>(lane) & (P - 1u);
            for (u32 k = 0; k < rr; ++k) {
                acc_lane[32u * k] = make_vec4<F>(F(0), F(0), F(0), F(0));
            }

            // phase 1 starts at the root; a group starts at the run's frontier (the root when there was no phase 1)
            u32 sp = ph1 ? 1u : 0u, fpos = 0u, lhead = 0, lcount = 0, lq_total = 0, lq_done = 0;
            const u32 fend = ph1 ? 0u : rs[RS_FCOUNT];
            u32 fcount = 0u; // phase 1: frontier nodes found so far
            bool done = false, overflow = false, fover = false;
            if (ph1) {
                if (lane == 0) {
                    stack[0] = 0u; // root: first = 0, count = 1
                }
            }
            __syncwarp();

            for (;;) {
                // ---------------- produce: fill the ring until >= 32 sources or the walk is over -------------
                while (lcount < TH && !done) {
                    if (lq_done < lq_total) {
                        // copy more particles of the rejected leaves into the ring
                        const u32 room = lcap - lcount, rem = lq_total - lq_done;
                        const u32 chunk = rem < room ? rem : room;
                        for (u32 f = lq_done + lane; f < lq_done + chunk; f += 32) {
                            int lo = 0;
#pragma unroll
                            for (int st = 16; st > 0; st >>= 1) {
                                if (lq_incl[lo + st - 1] <= f) {
                                    lo += st;
                                }
                            }
                            const u32 pidx = lq_base[lo] + f;
                            cp_async_vec4(&ring[ridx(lhead + lcount + (f - lq_done))], p.parts + pidx);
                        }
                        __syncwarp(); // (the copies stay in flight until the next consume step)
                        lcount += chunk;
                        lq_done += chunk;
                        continue;
                    }
                    bool have = false;
                    u32 k = 0;
                    if (sp != 0u) {
                        // ---- pop up to 4 (first child, count) entries: lane -> entry lane>>3, child lane&7 ----
                        // (an internal node of the 4M Plummer tree has 7.1 children on average, so this keeps
                        // ~7/8 of the lanes busy without the prefix scan / expansion buffer an exact 32-node pop needs)
                        const u32 eidx = static_cast<u32>(lane) >> 3, cidx = static_cast<u32>(lane) & 7u;
                        if (eidx < sp) {
                            const u32 e = stack[sp - 1u - eidx];
                            have = cidx <= (e & 7u);
                            k = (e >> 3) + cidx;
                        }
                        sp -= sp < 4u ? sp : 4u;
                    } else if (fpos < fend) {
                        // ---- the stack is empty: the next 32 frontier nodes ----
                        have = fpos + static_cast<u32>(lane) < fend;
                        k = *(stack_top - (have ? fpos + static_cast<u32>(lane) : 0u));
                        fpos += 32u;
                    } else {
                        done = true;
                        break;
                    }
                    __syncwarp();

                    // ---- one node per lane: classify ----
                    uint4 nb = make_uint4(0, 0, 0, 0);
                    vec4<F> na = make_vec4<F>(F(0), F(0), F(0), F(0));
                    if (have) {
                        nb = p.nodeB[k];
                        na = p.nodeA[k];
                    }
                    const u32 nch = nb.w & 0xffu, level = nb.w >> 8;
                    const bool is_self = have && k == gnode;
                    // (phase 1: [gb, ge) is the whole run, so this is "ancestor of every group")
                    const bool is_anc = have && !is_self && nb.x <= gb && ge <= nb.y;
                    // phase 1: a node overlapping the run is some group's own node, ancestor or descendant
                    bool fr = ph1 && have && !is_anc && nb.x < ge && nb.y > gb;
                    const bool test = have && !is_self && !is_anc && !fr;
                    F mac_lh = F(0);
                    if (test) {
                        if (MAC == 0) {
                            mac_lh = p.mac_tab[level];
                        } else {
                            const F t = rn_fma(p.mac_tab[level], p.mac_value, p.node_delta[k]);
                            mac_lh = rn_mul(t, t);
                        }
                    }
                    // Group MAC, tree.hpp:2741-2759: the node is accepted iff mac_lh < dist2 for EVERY target.
                    // (1) Lower-bound dist2 over the group with its bounding box: if even the nearest point of the
                    // box passes (with a guard band of 2^-20 >> the 8 eps rounding spread of the two evaluations)
                    // every target passes and the decision is the reference's. (2) Otherwise test ONE target with
                    // the reference's exactly rounded arithmetic - the group's support point towards the node,
                    // which is (nearly always) its nearest target: if it fails, the node is rejected, exactly as in
                    // the reference. On the 4M Plummer tree 80 % of the tests end at (1), 18 % at (2); only the
                    // remaining 2 % run the exact loop over all the targets.
                    // Phase 1 (box of a whole run): (2') if the MAC fails even at the farthest corner of the box it
                    // fails for every target, so every group rejects the node; anything else joins the frontier.
                    bool sure_acc = false, sure_rej = false;
                    if (test) {
                        F dmin2 = F(0);
                        const F c[3] = {na.x, na.y, na.z};
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const F gap = fmax(F(0), fmax(blo[j] - c[j], c[j] - bhi[j]));
                            dmin2 = fma(gap, gap, dmin2);
                        }
                        sure_acc = mac_lh < dmin2 * (F(1) - F(9.5367431640625e-07));
                        if (!sure_acc) {
                            if (ph1) {
                                F dmax2 = F(0);
#pragma unroll
                                for (int j = 0; j < 3; ++j) {
                                    const F far = fmax(c[j] - blo[j], bhi[j] - c[j]);
                                    dmax2 = fma(far, far, dmax2);
                                }
                                sure_rej = mac_lh >= dmax2 * (F(1) + F(9.5367431640625e-07));
                                fr = !sure_rej;
                            } else {
                                const u32 q = (c[1] < bmid[1] ? 2u : 0u) + (c[2] < bmid[2] ? 1u : 0u);
                                const u32 ci = ((c[0] < bmid[0] ? cand_neg : cand_pos) >> (8u * q)) & 0xffu;
                                const vec4<F> t = tpos[ci];
                                const F dx = rn_sub(na.x, t.x), dy = rn_sub(na.y, t.y), dz = rn_sub(na.z, t.z);
                                F d2 = rn_mul(dx, dx);
                                d2 = rn_fma(dy, dy, d2);
                                d2 = rn_fma(dz, dz, d2);
                                sure_rej = mac_lh >= d2;
                            }
                        }
                    }
                    const bool need = test && !sure_acc && !sure_rej && !fr;
                    bool fail = !need;
                    if (__any_sync(FULL, need)) {
                        if (staged) {
                            // cooperative: the ambiguous nodes are compacted into shared memory (com, mac_lh) and
                            // the 32 lanes share the targets of ONE node at a time (broadcast LDS.128 per node)
                            const u32 m_need = __ballot_sync(FULL, need);
                            const u32 n_need = __popc(m_need), my_slot = __popc(m_need & ltm);
                            // a free 32-entry block (lcount < TH here)
                            vec4<F> *amb = ring + (RESET ? (RK_RING_ROOM == 32 ? lcount : TH + 32u) : ridx(lhead + BATCH));
                            if (need) {
                                amb[my_slot] = make_vec4<F>(na.x, na.y, na.z, mac_lh);
                            }
                            __syncwarp();
                            u32 fail_bits = 0;
#if RK_AMB2
                            // two nodes per pass: each half-warp shares the targets of one node
                            const u32 half = static_cast<u32>(lane) >> 4, hl = static_cast<u32>(lane) & 15u;
#pragma unroll 1
                            for (u32 a = 0; a < n_need; a += 2u) {
                                const vec4<F> c = amb[a + half < n_need ? a + half : a];
                                bool f = false;
#pragma unroll 1
                                for (u32 i = hl; i < T; i += 16) {
                                    const vec4<F> t = tgt[i];
                                    const F dx = rn_sub(c.x, t.x), dy = rn_sub(c.y, t.y), dz = rn_sub(c.z, t.z);
                                    F d2 = rn_mul(dx, dx);
                                    d2 = rn_fma(dy, dy, d2);
                                    d2 = rn_fma(dz, dz, d2);
                                    f = f || (c.w >= d2);
                                }
                                const u32 fb = __ballot_sync(FULL, f);
                                fail_bits |= ((fb & 0xffffu) ? 1u : 0u) << a;
                                fail_bits |= ((fb >> 16) && a + 1u < n_need ? 1u : 0u) << (a + 1u);
                            }
#else
#pragma unroll 1
                            for (u32 a = 0; a < n_need; ++a) {
                                const vec4<F> c = amb[a];
                                bool f = false;
#pragma unroll 1
                                for (u32 i = lane; i < T; i += 32) {
                                    const vec4<F> t = tgt[i];
                                    const F dx = rn_sub(c.x, t.x), dy = rn_sub(c.y, t.y), dz = rn_sub(c.z, t.z);
                                    F d2 = rn_mul(dx, dx);
                                    d2 = rn_fma(dy, dy, d2);
                                    d2 = rn_fma(dz, dz, d2);
                                    f = f || (c.w >= d2);
                                }
                                fail_bits |= __any_sync(FULL, f) ? (1u << a) : 0u;
                            }
#endif
                            __syncwarp();
                            if (need) {
                                fail = (fail_bits >> my_slot) & 1u;
                            }
                        } else {
                            for (u32 i = 0; i < T; ++i) {
                                const vec4<F> t = gsrc[i];
                                const F dx = rn_sub(na.x, t.x), dy = rn_sub(na.y, t.y), dz = rn_sub(na.z, t.z);
                                F d2 = rn_mul(dx, dx);
                                d2 = rn_fma(dy, dy, d2);
                                d2 = rn_fma(dz, dz, d2);
                                fail = fail || (mac_lh >= d2);
                                if ((i & 7u) == 7u && __all_sync(FULL, fail)) {
                                    break;
                                }
                            }
                        }
                    }
                    const bool accept = sure_acc || (need && !fail);
                    const bool rejected = sure_rej || (need && fail);
                    const bool open_leaf = rejected && nch == 0u;
                    const bool descend = is_anc || (rejected && nch != 0u);
                    const u32 m_test = __ballot_sync(FULL, test && !fr), m_acc = __ballot_sync(FULL, accept),
                              m_leaf = __ballot_sync(FULL, open_leaf), m_desc = __ballot_sync(FULL, descend);
                    n_mac += __popc(m_test);
                    n_acc += __popc(m_acc);
                    if (ph1) {
                        // undecided / overlapping nodes -> frontier of the run
                        const u32 m_fr = __ballot_sync(FULL, fr);
                        if (m_fr) {
                            if (sp + fcount + 64u > STACK_CAP) { // (room for this step's 32 pushes as well)
                                fover = true;
                            } else if (fr) {
                                *(stack_top - (fcount + __popc(m_fr & ltm))) = k;
                            }
                            fcount += fover ? 0u : __popc(m_fr);
                        }
                    }
                    // accepted nodes -> ring
                    if (accept) {
                        ring[ridx(lhead + lcount + __popc(m_acc & ltm))] = na;
                    }
                    lcount += __popc(m_acc);
                    // rejected internal nodes / ancestors -> stack
                    if (m_desc) {
                        if (sp + 32u + (ph1 ? fcount : fend) > STACK_CAP) {
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
                        // queue the leaves (inclusive scan of their sizes); their particles are copied in ring-sized
                        // chunks by the next iterations of the produce loop (a per-leaf copy loop measured no faster)
                        const u32 li = warp_incl_scan(c, lane);
                        lq_incl[lane] = li;
                        lq_base[lane] = nb.x - (li - c);
                        lq_total = __shfl_sync(FULL, li, 31);
                        n_p2p += lq_total;
                    }
                    __syncwarp();
                    if (overflow || fover) {
                        done = true;
                    }
                }
                if (lcount == 0u || fover) {
                    break;
                }
                // ---------------- consume: evaluate up to 32 sources (the only ring call site) ----------------
                const u32 ne = (RESET || lcount < BATCH) ? lcount : BATCH;
                cp_async_wait_all(); // leaf particles still in flight
                __syncwarp();
                if (RK_SKIP_EVAL && p.G != F(-12345)) {
                    // timing experiment: walk only
                } else if constexpr (sizeof(F) == 4 && RK_PACKED) {
                    eval_slots_packed<Q, false>(reinterpret_cast<const float4 *>(ring + lhead), ne, sl, 5u - lp, eps2,
                                         reinterpret_cast<const float4 *>(tpos), T, t0 + tl, P, rr,
                                         reinterpret_cast<float4 *>(acc_lane));
                } else {
                    eval_slots<F, Q, false>(ring + lhead, ne, sl, 5u - lp, eps2, tpos, T, t0 + tl, P, rr, acc_lane);
                }
                __syncwarp();
                lhead = RESET ? 0u : ridx(lhead + ne);
                lcount -= ne;
            }
            if (overflow && lane == 0) {
                atomicExch(p.err, 1u);
            }
            if (ph1) {
                cp_async_wait_all();
                __syncwarp();
                if (fover) {
                    // more frontier nodes than the buffer holds: this run's groups walk from the root instead
                    if (lane == 0) {
                        stack_top[0] = 0u; // (RS_FCOUNT is still 1, the phase-1 counters 0, no partial sums)
                    }
                    __syncwarp();
                    break;
                }
                if (lane == 0) {
                    rs[RS_FCOUNT] = fcount;
                    rs[RS_PARTIAL] = 1u;
                    rs[RS_MAC1] = n_mac;
                    rs[RS_ACC1] = n_acc;
                    rs[RS_P2P1] = n_p2p;
                }
                __syncwarp();
            } else {
                // self interactions inside the group, tree.hpp:2073-2321 (sources = the group's own particles)
                if constexpr (sizeof(F) == 4 && RK_PACKED) {
                    eval_slots_packed<Q, true>(reinterpret_cast<const float4 *>(staged ? tgt : gsrc), T, sl, 5u - lp, eps2,
                                               reinterpret_cast<const float4 *>(tpos), T, t0 + tl, P, rr,
                                               reinterpret_cast<float4 *>(acc_lane));
                } else if (staged) {
                    eval_slots<F, Q, true>(tgt, T, sl, 5u - lp, eps2, tpos, T, t0 + tl, P, rr, acc_lane);
                } else {
                    eval_slots<F, Q, true>(gsrc, T, sl, 5u - lp, eps2, tpos, T, t0 + tl, P, rr, acc_lane);
                }
            }

            // Combine the slices' partial sums (fixed shuffle tree: deterministic) and put the totals back into the
            // accumulator area in TARGET order, so that the write-back below runs over consecutive targets with all 32
            // lanes: full 128-byte stores (which matters most for the mirrors, whose stores cross NVLink) instead of P
            // lanes per slot. Phase 1 parks the run's partial sums in the output arrays; a group adds them to its own,
            // applies G as one final multiply (tree.hpp:2986-3002) and writes out (3004-3007).
            vec4<F> *const accb = ph1 ? tgt : acc;
            __syncwarp(); // the accumulator stores of the evaluation loops above are ordered before the re-layout
            for (u32 k = 0; k < rr; ++k) {
                vec4<F> a = acc_lane[32u * k];
                for (u32 o = P; o < 32u; o <<= 1) {
                    a.x += __shfl_xor_sync(FULL, a.x, o);
                    a.y += __shfl_xor_sync(FULL, a.y, o);
                    a.z += __shfl_xor_sync(FULL, a.z, o);
                    a.w += __shfl_xor_sync(FULL, a.w, o);
                }
                // slot k of every lane has been read; target P * k + tl lies in a slot <= k, i.e. in one nobody reads again
                __syncwarp();
                if (sl == 0u && P * k + tl < tc) {
                    accb[P * k + tl] = a;
                }
            }
            __syncwarp();
            for (u32 i = t0 + static_cast<u32>(lane); i < t0 + tc; i += 32u) {
                if (ph1 && !(gb + i >= fb && gb + i < fe)) { // (phase 1: in-range targets only)
                    continue;
                }
                vec4<F> a = accb[i - t0];
                {
                    u32 dst = gb + i;
                    if (p.perm) {
                        dst = p.perm[dst];
                    }
                    dst -= p.out_offset;
                    if (ph1) {
                        if (Q == 0 || Q == 2) {
                            p.out[0][dst] = a.x;
                            p.out[1][dst] = a.y;
                            p.out[2][dst] = a.z;
                        }
                        if (Q != 0) {
                            p.out[Q == 1 ? 0 : 3][dst] = a.w;
                        }
                        continue;
                    }
                    if (rs[RS_PARTIAL]) {
                        if (Q == 0 || Q == 2) {
                            a.x += p.out[0][dst];
                            a.y += p.out[1][dst];
                            a.z += p.out[2][dst];
                        }
                        if (Q != 0) {
                            a.w += p.out[Q == 1 ? 0 : 3][dst];
                        }
                    }
                    const F tmass = tpos[i].w;
                    constexpr int NRES = Q == 0 ? 3 : (Q == 1 ? 1 : 4);
                    F v[NRES];
                    if (Q == 0 || Q == 2) {
                        v[0] = a.x * p.G;
                        v[1] = a.y * p.G;
                        v[2] = a.z * p.G;
                    }
                    if (Q != 0) {
                        v[NRES - 1] = (-tmass * a.w) * p.G;
                    }
#pragma unroll
                    for (int j = 0; j < NRES; ++j) {
                        __stcs(p.outf[j] + dst, v[j]);
                    }
                    // copies of the output arrays on other devices (peer memory) or in mapped host memory: the exchange
                    // of a multi-GPU evaluation happens here, store by store, underneath the arithmetic of the launch
                    for (u32 r = 0; r < p.n_mirror; ++r) {
                        if ((p.mirror_multicast >> r) & 1u) {
                            // an NVSwitch multicast address: one store leaves the GPU, the switch writes every rank's copy
#pragma unroll
                            for (int j = 0; j < NRES; ++j) {
                                multimem_store(p.mirror[r][j] + dst, v[j]);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < NRES; ++j) {
                                __stcs(p.mirror[r][j] + dst, v[j]);
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (!ph1 && t0 == 0 && lane == 0) {
                const u32 t_mac = n_mac + rs[RS_MAC1], t_acc = n_acc + rs[RS_ACC1], t_p2p = n_p2p + rs[RS_P2P1];
                if (p.group_cost) {
                    p.group_cost[g] = u64(T) * (u64(t_p2p) + t_acc + u64(T) - 1u);
                }
                if (p.counters) {
                    atomicAdd(p.counters + 0, u64(t_mac));
                    atomicAdd(p.counters + 1, u64(t_acc));
                    atomicAdd(p.counters + 2, u64(t_p2p) * T);
                    atomicAdd(p.counters + 3, u64(T) * (u64(T) - 1u) / 2u);
                    atomicAdd(p.counters + 4, u64(t_acc) * T);
                }
            }
            n_mac = n_acc = n_p2p = 0; // count the first pass only
            __syncwarp();
        }
        if (ph1 && rs[RS_SLOT] != NONE) {
            // publish the run: frontier + counters (the partial sums were written by the loop above)
            const u32 slot_ = rs[RS_SLOT], fc = rs[RS_FCOUNT];
            for (u32 i = lane; i < fc; i += 32u) {
                p.steal_front[size_t(slot_) * STACK_CAP + i] = *(stack_top - i);
            }
            volatile u32 *h = p.steal + size_t(slot_) * 16u;
            if (lane == 0) {
                h[1] = rs[RS_NG];
                h[3] = fc;
                h[4] = rs[RS_MAC1];
                h[5] = rs[RS_ACC1];
                h[6] = rs[RS_P2P1];
                h[7] = rs[RS_E0];
                h[8] = rs[RS_PARTIAL];
            }
            __threadfence(); // every lane: its partial sums and frontier entries ...
            __syncwarp();
            if (lane == 0) {
                __threadfence(); // ... are ordered before the flag a thief polls (release)
                h[2] = 1u;       // ready
                __threadfence();
                atomicAdd(p.steal_published, 1u);
            }
            __syncwarp();
        }
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

// FP64-pipe peak probe (same shape as ffma_kernel).
__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b)
{
    double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            v0 = fma(v0, a, b);
            v1 = fma(v1, a, b);
            v2 = fma(v2, a, b);
            v3 = fma(v3, a, b);
            v4 = fma(v4, a, b);
            v5 = fma(v5, a, b);
            v6 = fma(v6, a, b);
            v7 = fma(v7, a, b);
        }
    }
    const double s = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
    if (s == 123.456) {
        out[0] = s;
    }
}

template <typename F, int Q, int MAC, int BATCH>
int trav_occupancy(u32 tmax, size_t &smem, u32 ring = 2 * BATCH)
{
    smem = warp_smem_bytes<F>(tmax, ring) * TRAV_WARPS;
    int per_sm = 0;
    if (cudaFuncSetAttribute(traverse_kernel<F, Q, MAC, BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem))
            != cudaSuccess
        || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, traverse_kernel<F, Q, MAC, BATCH>, TRAV_THREADS, smem)
               != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return per_sm;
}

template <typename F, int Q, int MAC>
void launch_one(const trav_params<F> &p, int sm_count, cudaStream_t st, char *name)
{
    // batches of BIG sources (fp32: 128 at 4 CTAs/SM and 128 registers - measured 2.6 % faster at ncrit 128 and 7 % at
    // ncrit 256 than 64 at 5 CTAs/SM and 96 registers; fp64: 64) unless the larger ring costs a resident CTA
    constexpr int BIG = sizeof(F) == 4 ? RK_BATCH_BIG : 64;
    size_t smem64 = 0, smem32 = 0;
    const int occ64 = trav_occupancy<F, Q, MAC, BIG>(p.tmax, smem64), occ32 = trav_occupancy<F, Q, MAC, 32>(p.tmax, smem32);
    const bool big = occ64 >= occ32 && occ64 > 0;
    int per_sm = big ? occ64 : occ32;
    // the ring of the BIG variant takes the shared memory its resident CTAs leave unused (RESET mode of the kernel)
    u32 ring = 2 * BIG;
    if (big && RK_RING_RESET && RK_RING_GROW && (sizeof(F) == 4 || RK_RING_RESET_F64)) {
        static u32 cached_tmax = 0, cached_ring = 0; // (per instantiation; occupancy depends on tmax only)
        if (cached_tmax != p.tmax) {
            u32 r = ring;
            size_t sm = 0;
            while (r + 32u <= 1024u && trav_occupancy<F, Q, MAC, BIG>(p.tmax, sm, r + 32u) == occ64) {
                r += 32u;
            }
            cached_tmax = p.tmax;
            cached_ring = r;
        }
        ring = cached_ring;
        if (trav_occupancy<F, Q, MAC, BIG>(p.tmax, smem64, ring) != occ64) { // (sets the attribute for this size)
            throw cuda_error(1, "inconsistent occupancy of the traversal kernel");
        }
    }
    trav_params<F> q = p;
    q.ring = ring;
    if (name) {
        std::snprintf(name, 96, "traverse_kernel<%s,Q=%d,MAC=%d,BATCH=%d> window=%u ctas_per_sm=%d ring=%u",
                      sizeof(F) == 4 ? "float" : "double", Q, MAC, big ? BIG : 32, p.window, per_sm, big ? ring : 64u);
    }
    static const bool debug = std::getenv("RK_DEBUG_LAUNCH") != nullptr;
    if (debug) {
        std::fprintf(stderr, "[rk] traverse_kernel<%s,Q=%d,MAC=%d,BATCH=%d> CTAs/SM %d (64: %d with %zu B, 32: %d with %zu B) tmax %u window %u\n",
                     sizeof(F) == 4 ? "float" : "double", Q, MAC, big ? BIG : 32, per_sm, occ64, smem64, occ32, smem32, p.tmax,
                     p.window);
    }
    if (per_sm < 1) {
        throw cuda_error(1, "the traversal kernel does not fit on this device");
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
    if (big) {
        traverse_kernel<F, Q, MAC, BIG><<<grid, TRAV_THREADS, smem64, st>>>(q);
    } else {
        traverse_kernel<F, Q, MAC, 32><<<grid, TRAV_THREADS, smem32, st>>>(q);
    }
    count_launch();
    RK_CUDA_CHECK(cudaGetLastError());
}

} // namespace

#ifndef RK_TWO_PHASE
#define RK_TWO_PHASE 1
#endif
unsigned trav_stack_cap() { return STACK_CAP; }
unsigned trav_steal_k()
{
    static const unsigned k = [] {
        const char *e = std::getenv("RK_STEAL_K");
        const long v = e ? std::atol(e) : 2048;
        return static_cast<unsigned>(v < 0 ? 0 : (v > long(TRAV_STEAL_SLOTS) ? long(TRAV_STEAL_SLOTS) : v));
    }();
    return k;
}

u32 trav_window(u32 tmax, size_t max_group)
{
    // phase 1 keeps one accumulator per target of the run in the staged-target + accumulator areas of the warp:
    // a window of W particles holds groups that end before W + max_group - 1 <= (tmax + acc_entries) targets
    if (!RK_TWO_PHASE || max_group > tmax) {
        return 0u;
    }
    return acc_entries(tmax); // = 32 * rr_cap1 - tmax
}

template <typename F>
void launch_traverse(const trav_params<F> &p_, int Q, int mac, int sm_count, cudaStream_t st, char *name)
{
    trav_params<F> p = p_;
    for (int j = 0; j < 4; ++j) {
        p.outf[j] = p.outf[j] ? p.outf[j] : p.out[j]; // final results go where the partial sums are unless told otherwise
    }
#define RK_DISPATCH(QQ, MM)                                                                                            \
    if (Q == QQ && mac == MM) {                                                                                        \
        launch_one<F, QQ, MM>(p, sm_count, st, name);                                                                  \
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

double dfma_microbench(float *ms)
{
    int dev = 0, sms = 0;
    RK_CUDA_CHECK(cudaGetDevice(&dev));
    RK_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *d = nullptr;
    RK_CUDA_CHECK(cudaMalloc(&d, 64));
    cudaEvent_t e0, e1;
    RK_CUDA_CHECK(cudaEventCreate(&e0));
    RK_CUDA_CHECK(cudaEventCreate(&e1));
    const int iters = 1024, grid = sms * 8;
    dfma_kernel<<<grid, 256>>>(d, 16, 1.0001, 0.5); // warm-up
    RK_CUDA_CHECK(cudaEventRecord(e0));
    dfma_kernel<<<grid, 256>>>(d, iters, 1.0001, 0.5);
    RK_CUDA_CHECK(cudaEventRecord(e1));
    RK_CUDA_CHECK(cudaEventSynchronize(e1));
    RK_CUDA_CHECK(cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return 2.0 * 8 * 16 * double(iters) * 256.0 * grid;
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

template void launch_traverse<float>(const trav_params<float> &, int, int, int, cudaStream_t, char *);
template void launch_traverse<double>(const trav_params<double> &, int, int, int, cudaStream_t, char *);
template void launch_exact<float>(const vec4<float> *, size_t, size_t, float, float, double *, cudaStream_t);
template void launch_exact<double>(const vec4<double> *, size_t, size_t, double, double, double *, cudaStream_t);

} // namespace rk
