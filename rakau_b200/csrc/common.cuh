// common.cuh — shared declarations for the sm_100a kernels of librakau_b200.so.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>

namespace rk
{

using u64 = unsigned long long;
using u32 = unsigned;
using i8 = signed char;

constexpr int CBITS = 21;           // cbits_v<uint64_t,3>, reference detail/tree_fwd.hpp:141-150
constexpr int NLEVELS = CBITS + 1;  // levels 0..21

// Packed particle / node payload: (x, y, z, m). float4 = one LDG.128, double4 = two.
template <typename F>
struct vec4_of;
template <>
struct vec4_of<float> {
    using type = float4;
};
template <>
struct vec4_of<double> {
    using type = double4;
};
template <typename F>
using vec4 = typename vec4_of<F>::type;

template <typename F>
__host__ __device__ inline vec4<F> make_vec4(F x, F y, F z, F w)
{
    vec4<F> v;
    v.x = x;
    v.y = y;
    v.z = z;
    v.w = w;
    return v;
}

// Exactly-rounded primitives: nvcc must not contract or reassociate these (bit-exact discretisation and
// MAC decisions, SURVEY §7 hard part 1).
__device__ __forceinline__ float rn_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double rn_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float rn_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double rn_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float rn_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double rn_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float rn_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double rn_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float rn_sqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double rn_sqrt(double a) { return __dsqrt_rn(a); }

// Number of leading 3-bit digits (of 21) shared by two 63-bit Morton codes.
__host__ __device__ inline int shared_digits(u64 a, u64 b)
{
    const u64 x = a ^ b;
    if (x == 0) {
        return CBITS;
    }
#if defined(__CUDA_ARCH__)
    const int lz = __clzll(static_cast<long long>(x));
#else
    const int lz = __builtin_clzll(x);
#endif
    return (lz - 1) / 3;
}

struct cuda_error : std::runtime_error {
    int status;
    cuda_error(int st, const std::string &s) : std::runtime_error(s), status(st) {}
};

#define RK_CUDA_CHECK(expr)                                                                                            \
    do {                                                                                                               \
        cudaError_t rk_e_ = (expr);                                                                                    \
        if (rk_e_ != cudaSuccess) {                                                                                    \
            throw ::rk::cuda_error(rk_e_ == cudaErrorMemoryAllocation ? 5 : 4,                                         \
                                   std::string("CUDA error in ") + #expr + ": " + cudaGetErrorString(rk_e_));           \
        }                                                                                                              \
    } while (0)

// Number of kernels this library has launched (bench.py reports it as gpu_launches).
extern unsigned long long g_kernel_launches;
inline void count_launch(unsigned n = 1) { __atomic_fetch_add(&g_kernel_launches, n, __ATOMIC_RELAXED); }

inline unsigned div_up(size_t a, size_t b) { return static_cast<unsigned>((a + b - 1) / b); }

// Grow-only device buffer.
template <typename T>
struct dbuf {
    T *p = nullptr;
    size_t cap = 0;
    void reserve(size_t n, double slack = 1.0)
    {
        if (n <= cap) {
            return;
        }
        release();
        const size_t want = static_cast<size_t>(static_cast<double>(n) * slack) + 16;
        RK_CUDA_CHECK(cudaMalloc(reinterpret_cast<void **>(&p), want * sizeof(T)));
        cap = want;
    }
    void release()
    {
        if (p) {
            cudaFree(p);
        }
        p = nullptr;
        cap = 0;
    }
    ~dbuf() { release(); }
    dbuf() = default;
    dbuf(const dbuf &) = delete;
    dbuf &operator=(const dbuf &) = delete;
};

// ---------------------------------------------------------------------------------------------------
// Radix sort (sort.cu)
// ---------------------------------------------------------------------------------------------------
struct sort_scratch {
    dbuf<u32> ghist;    // 8 x 256 global digit histograms
    dbuf<u32> tilehist; // 256 x ntiles (digit-major)
    u32 *h_ghist = nullptr; // pinned, 8 x 256
};
// Stable LSD radix sort of (key, idx) pairs by 63-bit key. keys_a holds the input; on return *keys_out /
// *idx_out point at the buffers (a or b) holding the sorted result. idx input is implicit iota.
// Returns the number of passes run.
int radix_sort_pairs(u64 *keys_a, u64 *keys_b, u32 *idx_a, u32 *idx_b, size_t n, sort_scratch &sc, cudaStream_t st,
                     u64 **keys_out, u32 **idx_out);

void radix_partition_pass(const u64 *keys_in, u64 *keys_out, u32 *idx_out, size_t n, sort_scratch &sc, cudaStream_t st,
                          const u32 **counts_dev);

// ---------------------------------------------------------------------------------------------------
// Build (build.cu)
// ---------------------------------------------------------------------------------------------------
// Error record written by device kernels (first offending particle wins).
struct dev_error {
    u32 code;  // 0 none; 1 non-finite discretisation; 2 out-of-bounds fp; 3 out-of-bounds int;
               // 4 non-finite coordinate (box deduction); 5 com non-finite; 6 mass non-finite; 7 delta non-finite
    u32 index; // particle or node index
    u32 dim;   // coordinate index
    u32 pad;
};

struct level_table {
    u32 base[NLEVELS + 2]; // BFS base of each level; base[NLEVELS] = M (end of last level), base[NLEVELS+1] unused
};

template <typename F>
struct build_arrays {
    size_t n = 0;
    // particles
    dbuf<F> stage[4];         // SoA staging (x,y,z,m) for H2D / getters
    dbuf<vec4<F>> pin;        // packed, pre-sort order
    dbuf<vec4<F>> psorted;    // packed, Morton order
    dbuf<u64> keys_a, keys_b; // sort double buffers
    dbuf<u32> idx_a, idx_b;
    u64 *codes = nullptr;     // sorted codes (one of keys_a/keys_b)
    u32 *last_perm = nullptr; // sorted indices (one of idx_a/idx_b)
    dbuf<u32> perm, perm_tmp, inv_perm;
    // topology scratch (per particle)
    dbuf<i8> delta, lvl_leaf, lvl_crit; // delta(i), D(i), Lc(i)
    dbuf<i8> win_a, win_b;              // sliding-window doubling buffers
    dbuf<u32> dfsbase;                  // n+1
    dbuf<u32> tilecnt;                  // (NLEVELS+1) x ntiles
    dbuf<u32> rowtot;                   // NLEVELS+1
    // nodes, BFS (level-major) order
    size_t n_nodes = 0, n_crit = 0;
    dbuf<vec4<F>> nodeA; // com xyz + mass
    dbuf<uint4> nodeB;   // begin, end, first child, (nch | level << 8)
    dbuf<F> node_delta;  // bh_geom only
    dbuf<u32> node_dfs;  // DFS pre-order index of each BFS node
    dbuf<u32> node_ndesc; // number of descendants (the reference's n_children)
    dbuf<u32> crit_node;  // BFS index of each critical node
    dbuf<u32> crit_begin; // n_crit + 1 (crit_begin[n_crit] = n)
    dbuf<double> chunksum; // 4 doubles per chunk of PROPS_CHUNK particles
    dbuf<double> nodesum;  // 4 doubles per node (bottom-up node properties)
    int props_bottom_up = -1; // node properties: -1 by size (bottom-up from PROPS_BOTTOMUP_MIN particles), 0 / 1 forced
    level_table levels;
    dbuf<dev_error> d_err;
    dbuf<u32> d_misc; // [0..1] abs-max bits (u64), [2] max group size
};

constexpr int PROPS_CHUNK = 256;
constexpr size_t PROPS_BOTTOMUP_MIN = size_t(8) << 20;
#ifndef RK_TOPO_TILE
#define RK_TOPO_TILE 512
#endif
constexpr int TOPO_TILE = RK_TOPO_TILE; // particles per CTA of the node count / emit kernels

// Kernels are wrapped in launch functions so that capi.cu stays free of <<<>>> syntax details.
template <typename F>
void launch_pack_absmax(const F *x, const F *y, const F *z, const F *m, vec4<F> *out, size_t n, u64 *absmax_bits,
                        cudaStream_t st);
template <typename F>
void launch_unpack(const vec4<F> *in, F *x, F *y, F *z, F *m, size_t n, cudaStream_t st);
template <typename F>
void launch_set_coords(vec4<F> *p, const F *x, const F *y, const F *z, const F *m, size_t n, u64 *absmax_bits,
                       cudaStream_t st);
template <typename F>
void launch_encode(const vec4<F> *p, u64 *codes, size_t n, F inv_box, dev_error *err, cudaStream_t st);
template <typename F>
void launch_gather(const vec4<F> *pin, const u32 *idx, vec4<F> *pout, size_t n, cudaStream_t st,
                   const F *late_m = nullptr);
void launch_perm_invert(const u32 *perm, u32 *inv_perm, size_t n, cudaStream_t st);
void launch_perm_first(const u32 *last_perm, u32 *perm, u32 *inv_perm, size_t n, cudaStream_t st);
void launch_perm_compose(const u32 *old_perm, const u32 *last_perm, u32 *new_perm, u32 *inv_perm, size_t n,
                         cudaStream_t st);
void launch_iota(u32 *p, size_t n, cudaStream_t st);
// ids[i] = number of splitters <= codes[i] (the bucket of a sample sort, < 256); out[i] = in[idx[i]] for 64-bit words
void launch_bucket_ids(const u64 *codes, size_t n, const u64 *splitters, unsigned nsplit, u64 *ids, cudaStream_t st);
void launch_gather_u64(const u64 *in, const u32 *idx, u64 *out, size_t n, cudaStream_t st);
// the same bytes to nd <= 8 destinations (8-byte aligned, like src) with one kernel: one read, nd stores per word
// multicast: dst[0] is an NVSwitch multicast address (bytes % 8 == 0), written once with multimem.st
void launch_bcast_copy(void *const *dst, int nd, const void *src, size_t bytes, int sm_count, cudaStream_t st,
                       bool multicast = false);
template <typename F>
void launch_scatter_perm(const F *in, const u32 *perm, F *out, size_t n, cudaStream_t st);
// out += sum over the 32-bit words w_i of the array of mix64(i, w_i): an order-independent fingerprint of a device array
// (bytes must be a multiple of 4).
void launch_digest(const void *a, size_t bytes, u64 *out, cudaStream_t st);
// out[j] = first index i in [0, n] with arr[i] >= x[j] (n if none).
void launch_lower_bound(const u32 *arr, size_t n, const u64 *x, size_t k, u64 *out, cudaStream_t st);

// Topology: fills delta/lvl_leaf/lvl_crit, tilecnt, rowtot. After it the host reads rowtot to size the nodes.
template <typename F>
void topology_count(build_arrays<F> &b, size_t max_leaf_n, size_t ncrit, cudaStream_t st);
// Emits nodes in BFS order + critical nodes; requires b.levels and node buffers sized.
template <typename F>
void topology_emit(build_arrays<F> &b, cudaStream_t st);
// Node properties (mass, com, delta). dim_tab: node dimension per level (host-computed, F precision).
template <typename F>
void node_properties(build_arrays<F> &b, int mac, F box_size, cudaStream_t st);
// Host AoS (DFS order) from the BFS SoA.
template <typename F>
void launch_export_nodes(const build_arrays<F> &b, int mac, F box_size, void *d_out /* rk_node_f32/f64[M] */,
                         cudaStream_t st);
void launch_export_crit(const u64 *codes, const uint4 *nodeB, const u32 *crit_node, const u32 *crit_begin, size_t n_crit,
                        u64 *d_out /* triplets */, cudaStream_t st);
// Import of an external DFS AoS tree (rk_traverse_external_tree): fills BFS arrays + critical nodes for ncrit.
template <typename F>
void import_external_tree(build_arrays<F> &b, const void *d_nodes_aos, size_t tree_size, int mac, size_t ncrit,
                          cudaStream_t st);

// ---------------------------------------------------------------------------------------------------
// Leapfrog (leapfrog.cu): kick / drift / re-index / conserved quantities of benchmark_leapfrog.cpp:286-384
// ---------------------------------------------------------------------------------------------------
template <typename F>
void launch_lf_reorder(const F *const in[3], const u32 *perm, F *const out[3], size_t n, cudaStream_t st);
template <typename F>
void launch_lf_kick_drift(const F *const acc[3], const F *const v[3], const vec4<F> *pos, F half_dt, F dt, F *const kv[3],
                          vec4<F> *pin, size_t n, u64 *absmax, cudaStream_t st);
template <typename F>
void launch_lf_kick_reindex(const F *const acc[3], const F *const kv[3], const u32 *last_perm, F half_dt, F *const v[3],
                            size_t n, cudaStream_t st);
// scratch[0..6] = sum x, y, z, vx, vy, vz, (m v^2 / 2 + pot); scratch needs lf_scratch_doubles() doubles.
template <typename F>
void launch_lf_integrals(const vec4<F> *pos, const F *const v[3], const F *pot, size_t n, double *scratch, cudaStream_t st);
unsigned lf_scratch_doubles();

// ---------------------------------------------------------------------------------------------------
// Traversal (traverse.cu)
// ---------------------------------------------------------------------------------------------------
constexpr unsigned TRAV_MAX_MIRRORS = 8;
template <typename F>
struct trav_params {
    const vec4<F> *parts;
    const vec4<F> *nodeA;
    const uint4 *nodeB;
    const F *node_delta;
    const u32 *crit_node;
    const u32 *crit_begin;
    u32 c0, c1;        // critical-node range
    u32 ncrit;         // all critical nodes of the tree
    u32 *work_counter; // zeroed before launch
    F mac_tab[NLEVELS]; // bh: dim2(level) * theta^-2 ; bh_geom: dim(level)
    F mac_value, eps2, G;
    F *out[4];
    // where the FINAL results are written (streaming stores): NULL = out[j]. With host outputs in pinned memory this is
    // the caller's buffer itself (mapped into the device's address space), so no device-to-host copy follows the launch;
    // out[j] (device memory) still receives the partial sums phase 1 parks for phase 2, which are read back.
    F *outf[4];
    // further copies of the output arrays that receive every final result (same index): the other ranks' peer-mapped
    // buffers of a multi-GPU evaluation, a mapped host buffer ...
    F *mirror[TRAV_MAX_MIRRORS][4];
    u32 n_mirror;
    u32 mirror_multicast; // bit r: mirror r is an NVSwitch multicast address (written with multimem.st)
    const u32 *perm; // non-null: ordered outputs (scatter through perm)
    u64 *group_cost; // per critical node (nullable)
    u64 *counters;   // 5 x u64: mac_tests, accepted, p2p_pairs, self_pairs, sum T*accepted (nullable)
    u32 tmax;        // targets staged in shared memory per warp
    u32 ring;        // entries of the source ring per warp (set by launch_traverse)
    u32 *err;        // stack overflow flag
    u32 out_offset;  // subtracted from the particle index when writing (external-tree drop-in)
    u32 window;      // two-phase walk: particles per run of sibling groups (trav_window()); 0 = one group per unit
    // work stealing at the tail of a launch (NULL = off): steal_k records of 16 words (next group, groups, ready,
    // frontier size, phase-1 counters, first group, partial flag), zeroed before the launch, + their frontiers
    u32 *steal, *steal_front, *steal_published;
    u32 steal_k;
};
constexpr unsigned TRAV_STEAL_SLOTS = 2048; // records allocated
// records used: the runs of the last wave that publish themselves (RK_STEAL_K in the environment overrides it: tuning)
unsigned trav_steal_k();
// words per stolen frontier = the per-warp stack capacity of traverse.cu
unsigned trav_stack_cap();
// Window of the two-phase walk for groups of at most max_group targets (0 when a group exceeds the staging area).
u32 trav_window(u32 tmax, size_t max_group);
// name: NULL or a buffer of >= 96 bytes that receives the launched variant ("traverse_kernel<float,Q=0,...>").
template <typename F>
void launch_traverse(const trav_params<F> &p, int Q, int mac, int sm_count, cudaStream_t st, char *name = nullptr);
// FFMA / DFMA-bound microbenchmarks on the current device: return the flop count, *ms the CUDA-event time.
double ffma_microbench(float *ms);
double dfma_microbench(float *ms);
template <typename F>
void launch_exact(const vec4<F> *parts, size_t n, size_t idx, F G, F eps2, double *d_out4, cudaStream_t st);

} // namespace rk
