/* rakau_b200.h — C ABI of librakau_b200.so, the B200 (sm_100a) implementation of rakau's Barnes-Hut hot path.
 *
 * This is the drop-in boundary. The reference has exactly one accelerator seam,
 *   include/rakau/detail/cuda_fwd.hpp:23-30   cuda_min_size / cuda_device_count / cuda_acc_pot_impl<Q,NDim,F,UInt,MAC>
 * called from include/rakau/tree.hpp:3135,3191,3207,3220 (traversal only; build is CPU-only in the reference).
 * The entry points below replace that seam (rk_min_size, rk_device_count, rk_traverse_external_tree) and
 * add the seam the reference lacks: a device-resident tree (Morton encode, sort, octree build, node
 * properties — reference construct_impl tree.hpp:1329-1487, build_tree 932-1111, sync 3678-3743) whose
 * traversal (acc_pot_impl 2853-3265) never leaves the GPU.
 *
 * Conventions
 *  - plain C types only; floating-point arrays are `const void*`/`void*` holding float (fp_bits=32) or
 *    double (fp_bits=64), matching the tree's precision;
 *  - every function returning int returns an rk_status; on error rk_last_error(tree) holds the message the
 *    C++ wrapper (include/rakau/tree.hpp in this repo) rethrows as the reference's exception type;
 *  - a tree handle owns all of its device state; const operations on one handle are serialised internally;
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails with RK_ERR_RUNTIME.
 */
#ifndef RAKAU_B200_H
#define RAKAU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RK_API __attribute__((visibility("default")))
#else
#define RK_API
#endif

typedef enum rk_status {
    RK_OK = 0,
    RK_ERR_INVALID_ARGUMENT = 1, /* std::invalid_argument in the reference */
    RK_ERR_DOMAIN = 2,           /* std::domain_error (theta / eps / G checks, tree.hpp:3268-3317) */
    RK_ERR_OVERFLOW = 3,         /* std::overflow_error */
    RK_ERR_RUNTIME = 4,          /* std::runtime_error (any CUDA failure, rakau_cuda.cu:83-127) */
    RK_ERR_BAD_ALLOC = 5         /* std::bad_alloc (rakau_cuda.cu:47-55) */
} rk_status;

enum { RK_MAC_BH = 0, RK_MAC_BH_GEOM = 1 };       /* rakau::mac, detail/tree_fwd.hpp:46 */
enum { RK_Q_ACCS = 0, RK_Q_POTS = 1, RK_Q_ACCS_POTS = 2 }; /* Q, detail/tree_fwd.hpp:129-137 */
enum { RK_PERM = 0, RK_LAST_PERM = 1, RK_INV_PERM = 2 };
enum { RK_HOST = 0, RK_DEVICE = 1 };

typedef struct rk_tree rk_tree;

/* Node as returned to the host: field order of base_tree_node_t + tree_node_t (detail/tree_fwd.hpp:76-116).
 * `dim` is dim2 for RK_MAC_BH and dim for RK_MAC_BH_GEOM; `delta` is 0 for RK_MAC_BH. */
typedef struct rk_node_f32 {
    uint64_t begin, end, n_children, code, level;
    float props[4], dim, delta;
} rk_node_f32;
typedef struct rk_node_f64 {
    uint64_t begin, end, n_children, code, level;
    double props[4], dim, delta;
} rk_node_f64;
/* Critical node, tree_cnode_t (detail/tree_fwd.hpp:119-125). */
typedef struct rk_cnode {
    uint64_t code, begin, end;
} rk_cnode;

typedef struct rk_build_info {
    double box_size;      /* deduced or given */
    uint64_t n_nodes;     /* tree size */
    uint64_t n_crit;      /* critical nodes */
    uint64_t max_group;   /* largest critical node */
    uint32_t sort_passes; /* radix passes actually run */
    float ms_total;       /* CUDA-event time of the whole build on the tree's stream */
    float ms_encode, ms_sort, ms_permute, ms_topology, ms_props;
} rk_build_info;

typedef struct rk_eval_info {
    uint64_t mac_tests;    /* node MAC evaluations (group level) */
    uint64_t accepted;     /* accepted nodes */
    uint64_t p2p_pairs;    /* target x leaf-particle pairs */
    uint64_t self_pairs;   /* unordered pairs inside the groups */
    uint64_t interactions; /* sum over groups of tgt*(leaf sources + accepted + tgt-1), SURVEY §8(d) */
    uint64_t n_groups;     /* groups evaluated by this call */
    uint32_t kernel_launches;
    float ms_kernel;       /* CUDA-event time of the traversal kernel(s) only */
    float ms_total;        /* kernel + output scatter/copies issued by the call */
} rk_eval_info;

typedef struct rk_leapfrog_info {
    float ms_step; /* CUDA-event time of the whole step */
    float ms_integrals, ms_kick_drift, ms_rebuild, ms_traverse, ms_reindex;
    uint64_t interactions, n_nodes;
    double com[3], com_v[3], energy; /* track_integrals: values at the BEGINNING of the step (benchmark_leapfrog.cpp:292-347) */
} rk_leapfrog_info;

/* ---- the reference's seam ---------------------------------------------------------------------------- */
/* cuda_device_count(), src/rakau_cuda.cu:32 — 0 on any CUDA error. */
RK_API unsigned rk_device_count(void);
/* cuda_min_size(), src/rakau_cuda.cu:26 — minimum number of particles worth sending to a GPU. */
RK_API unsigned rk_min_size(void);

/* ---- tree lifetime ----------------------------------------------------------------------------------- */
RK_API rk_tree *rk_tree_create(int fp_bits, int mac, int device);
RK_API void rk_tree_destroy(rk_tree *t);
RK_API const char *rk_last_error(const rk_tree *t);
/* Message of the last failed rk_tree_create (no handle exists to ask). */
RK_API const char *rk_create_error(void);
/* Use an externally owned cudaStream_t (e.g. torch's current stream) for all work of this tree. The handle is
 * taken literally: 0 is the legacy default stream (torch's default stream); (void*)-1 reverts to the tree's own
 * non-blocking stream. */
RK_API int rk_tree_set_stream(rk_tree *t, void *cuda_stream);
RK_API int rk_tree_synchronize(rk_tree *t);
/* Tuning switches that never change results beyond the documented tolerances. "props_bottom_up": node properties
 * bottom-up (1: every particle read once, one launch per level), top-down (0) or chosen by size (-1, the default).
 * "zero_copy_out" (default 1): unordered results for host buffers in PINNED memory (cudaHostAlloc / cudaHostRegister)
 * are written by the traversal kernel straight into those buffers, with no device-to-host copy behind the launch;
 * 0 = always copy. Pageable host buffers and ordered outputs are always copied. */
RK_API int rk_tree_set_option(rk_tree *t, const char *name, long long value);
/* Output mirrors: n <= 8 further copies of the output arrays (ptrs[4 * r + j] = array j of mirror r, device-accessible
 * memory: another GPU's buffer mapped into this process, mapped pinned host memory) that every following
 * rk_tree_acc_pot / rk_tree_acc_pot_range with DEVICE outputs writes as well, at the same indices, from inside the
 * traversal kernel - the output exchange of a multi-GPU evaluation (each rank mirrors its range into its peers'
 * buffers over NVLink while it computes; src/rakau_cuda.cu:492-527 copies each device's slice back afterwards instead).
 * n = 0 switches it off. The caller orders the peers' reads after this call (a barrier across the ranks).
 * multicast_mask: bit r set = mirror r is an NVSwitch MULTICAST address (cuMulticast / symmetric memory): it is written
 * with multimem.st, one store leaving the GPU and the switch updating every rank's copy. */
RK_API int rk_tree_set_output_mirrors(rk_tree *t, unsigned n, void *const *ptrs, unsigned multicast_mask);

/* ---- construction: construct_impl, tree.hpp:1329-1487 ------------------------------------------------ */
/* Copies n particles (SoA x, y, z, m; host or device pointers) into the tree, deduces the box if
 * `deduce_box` (determine_box_size, tree.hpp:1278-1319), Morton-encodes (disc_single_coord 381-429 +
 * morton_encoder 222-242), stable-sorts (indirect_code_sort 1266-1274), permutes (apply_isort 484-507),
 * builds nodes + critical nodes (build_tree 932-1111) and node properties (compute_node_properties
 * 1116-1237). */
RK_API int rk_tree_build(rk_tree *t, const void *x, const void *y, const void *z, const void *m, size_t n, int where,
                         double box_size, int deduce_box, size_t max_leaf_n, size_t ncrit, rk_build_info *info);
/* sync(), tree.hpp:3678-3743: new coordinates and masses in the CURRENT Morton order (NULL = unchanged), as left
 * behind by the update_particles_u functor (tree.hpp:3746-3765). */
RK_API int rk_tree_update_positions(rk_tree *t, const void *x, const void *y, const void *z, const void *m, int where,
                                    rk_build_info *info);
/* update_masses_dispatch, tree.hpp:3782-3805: new masses in the current Morton order; topology untouched. */
RK_API int rk_tree_update_masses(rk_tree *t, const void *m, int where);
/* ---- multi-GPU building blocks (one process per GPU; no reference counterpart, the reference is single-process) --
 * Distributed sample sort: every rank sorts its shard with the GLOBAL box (rk_tree_sort_shard; codes == NULL:
 * encode first), exchanges splitter buckets (NCCL all-to-all, done by the caller), sorts its bucket again with
 * the received codes, all-gathers the sorted buckets and builds the replicated tree from the globally sorted
 * arrays (rk_tree_build_presorted: topology + node properties only). All pointers are DEVICE pointers.
 * parts_ready_event: NULL, or a cudaEvent_t after which x, y, z, m and perm are valid - the topology is built from
 * the codes alone, so the caller can gather the particle arrays on another stream underneath it. */
RK_API int rk_tree_sort_shard(rk_tree *t, const void *x, const void *y, const void *z, const void *m,
                              const uint64_t *codes, size_t n, double box_size);
/* The same exchange without the local pre-sort: rk_tree_encode_shard packs and Morton-encodes the shard (codes in input
 * order, readable with rk_tree_get_codes_device, from which the ranks sample their splitters); rk_tree_partition_shard
 * then groups the particles by splitter bucket with ONE stable radix pass on the bucket id (bucket = number of
 * splitters <= code; nsplit <= 255 device-resident ascending splitters) and returns the nsplit + 1 bucket sizes in
 * counts (host). Afterwards codes / particles / last_perm of the shard are in bucket order. */
RK_API int rk_tree_encode_shard(rk_tree *t, const void *x, const void *y, const void *z, const void *m, size_t n,
                                double box_size);
RK_API int rk_tree_partition_shard(rk_tree *t, const uint64_t *splitters, unsigned nsplit, uint64_t *counts);
RK_API int rk_tree_get_codes_device(rk_tree *t, uint64_t *out);
RK_API int rk_tree_build_presorted(rk_tree *t, const void *x, const void *y, const void *z, const void *m,
                                   const uint64_t *codes, const uint32_t *perm, size_t n, double box_size,
                                   size_t max_leaf_n, size_t ncrit, void *parts_ready_event, rk_build_info *info);
/* Copy-engine device-to-device copy on `stream` (cudaMemcpyAsync). dst may be PEER memory mapped into this process
 * (CUDA IPC / symmetric memory): the multi-GPU output exchange pushes finished result slices into every peer's
 * buffer with it while the next traversal launch occupies the SMs (an NCCL kernel would wait for them). */
RK_API int rk_device_copy_async(void *dst, const void *src, size_t bytes, void *stream);
/* The same bytes to ndst <= 8 destinations (8-byte aligned like src; typically the same slice of every peer's buffer)
 * with ONE kernel on `stream`: every 8-byte word is read once and stored ndst times, the remote stores travelling over
 * NVLink. For the moments of the multi-GPU build when the SMs have nothing else to do (the all-gather of the bucket
 * codes, which the topology needs before it can start). multicast != 0: dst[0] (ndst == 1, bytes % 8 == 0) is an NVSwitch
 * multicast address; every word leaves the GPU once (multimem.st) and the switch writes all the copies. */
RK_API int rk_device_bcast_copy(void *const *dst, unsigned ndst, const void *src, size_t bytes, void *stream, int multicast);
/* determine_box_size's final arithmetic (tree.hpp:1309-1312) for a global max |coordinate|. */
RK_API double rk_deduce_box(int fp_bits, double absmax);
/* Copy constructor / assignment (tree.hpp:1735-1743, 1785-1822): deep device-to-device copy of src into dst
 * (same fp_bits and mac). */
RK_API int rk_tree_clone(rk_tree *dst, const rk_tree *src);
/* clear(), tree.hpp:1882-1904. */
RK_API int rk_tree_clear(rk_tree *t);

/* ---- getters (lazy D2H) ------------------------------------------------------------------------------ */
RK_API size_t rk_tree_nparts(const rk_tree *t);
RK_API size_t rk_tree_nnodes(const rk_tree *t);
RK_API size_t rk_tree_ncrit(const rk_tree *t);
RK_API double rk_tree_box_size(const rk_tree *t);
/* p_its_u(): Morton-ordered SoA (any pointer may be NULL). */
RK_API int rk_tree_get_parts(rk_tree *t, void *x, void *y, void *z, void *m);
/* Same, written to DEVICE pointers (device-resident integrators: the leapfrog of benchmark_leapfrog.cpp
 * without a host round trip). */
RK_API int rk_tree_get_parts_device(rk_tree *t, void *x, void *y, void *z, void *m);
/* perm / last_perm / inv_perm as uint32_t[nparts] on the DEVICE (re-indexing velocities after an update,
 * benchmark_leapfrog.cpp:375-383). */
RK_API int rk_tree_get_perm_device(rk_tree *t, int which, uint32_t *out);
/* c_it_u(): sorted Morton codes. */
RK_API int rk_tree_get_codes(rk_tree *t, uint64_t *codes);
/* perm()/last_perm()/inv_perm(), widened to 64 bit. */
RK_API int rk_tree_get_perm(rk_tree *t, int which, uint64_t *out);
/* nodes(): rk_node_f32 / rk_node_f64 array in the reference's DFS pre-order. */
RK_API int rk_tree_get_nodes(rk_tree *t, void *nodes);
RK_API int rk_tree_get_crit(rk_tree *t, rk_cnode *crit);

/* ---- traversal: acc_pot_dispatch / acc_pot_impl, tree.hpp:3293-3334, 2853-3265 ----------------------- */
/* out[0..2] accelerations (Q=0), out[0] potentials (Q=1), out[0..3] accs+pots (Q=2); each of nparts
 * elements, host or device. ordered=0: Morton order (accs_u ...); ordered=1: original order (accs_o ...).
 * split/nsplit: the reference's `split` kwarg (validated as tree.hpp:2857-2868); when [crit_begin,crit_end)
 * is given through rk_tree_acc_pot_range the call evaluates only those critical nodes. */
RK_API int rk_tree_acc_pot(rk_tree *t, int Q, int ordered, double theta, double G, double eps, const double *split,
                           size_t nsplit, void *const out[4], int where, rk_eval_info *info);
/* Evaluate critical nodes [crit_begin, crit_end) only (multi-GPU sharding by Morton range). Outputs are
 * written for the particles those nodes cover, at their global positions in `out`. */
RK_API int rk_tree_acc_pot_range(rk_tree *t, int Q, int ordered, double theta, double G, double eps, size_t crit_begin,
                                 size_t crit_end, void *const out[4], int where, rk_eval_info *info);
/* First particle of the critical nodes idx[0..k) (idx = ncrit gives nparts): the particle range a Morton range of
 * critical nodes covers, without fetching the whole list. */
RK_API int rk_tree_crit_begin_at(rk_tree *t, const size_t *idx, size_t k, uint64_t *out);
/* Per-critical-node interaction counts of the last full evaluation (cost weights for sharding). */
RK_API int rk_tree_get_group_costs(rk_tree *t, uint64_t *costs);
/* Device pointer to the same per-critical-node costs (uint64_t[ncrit]); NULL before the first evaluation. */
RK_API const void *rk_tree_group_costs_device(rk_tree *t);
/* First critical node whose first particle is >= particle_idx[j] (ncrit if none), k <= 16: re-snaps cuts kept as
 * particle indices to the critical nodes of a rebuilt tree (the reference snaps its split the same way,
 * tree.hpp:3168-3178). */
RK_API int rk_tree_crit_lower_bound(rk_tree *t, const uint64_t *particle_idx, size_t k, uint64_t *out);
/* out[j][perm[i]] = in[j][i] for nres DEVICE arrays of nparts elements: results from the tree's Morton order to the
 * original particle order (what the `_o` functions return, tree.hpp:3320-3330), asynchronously on the tree's stream.
 * For callers that assembled Morton-order results themselves (the one-process-per-GPU evaluation). */
RK_API int rk_tree_to_original_order(rk_tree *t, int nres, const void *const in[4], void *const out[4]);
/* Order-independent 64-bit fingerprints of the device-resident arrays: codes, perm, particles (Morton order), node
 * topology (begin, end, first child, level | children), node properties (com, mass), critical node indices, critical
 * node first particles, (n_nodes << 32) ^ n_crit. Two trees with equal fingerprints are equal bit for bit (up to
 * hash collisions): the multi-GPU path checks its tree against the single-GPU build with it. */
RK_API int rk_tree_digest(rk_tree *t, uint64_t out[8]);
/* Name and launch configuration of the traversal kernel variant the last evaluation of this tree launched. */
RK_API const char *rk_tree_last_kernel(const rk_tree *t);
/* exact_acc_pot_impl, tree.hpp:3531-3569: direct sum for one particle. idx in Morton order (ordered=0) or
 * original order (ordered=1). out4 = ax, ay, az, pot. */
RK_API int rk_tree_exact(rk_tree *t, size_t idx, int ordered, double G, double eps, double out4[4]);

/* ---- literal drop-in for cuda_acc_pot_impl (cuda_fwd.hpp:27-30, rakau_cuda.cu:348-528) ---------------- */
/* Stateless: host AoS tree in DFS order (rk_node_f32/f64), host SoA particles in Morton order, host codes.
 * Evaluates particles [split_indices[0], nparts) like the reference (the GPU share), writing to out[j] at
 * offset 0 (offset_output=0) or at split_indices[0] (offset_output=1). mac_value is theta^-2 (bh) or
 * theta^-1 (bh_geom), as passed by tree.hpp:3207. ncrit (not in the reference signature) selects the
 * target grouping of the CPU path; pass 0 for the reference default (128). */
RK_API int rk_traverse_external_tree(int fp_bits, int mac, int Q, void *const out[4], const uint64_t *split_indices,
                                     size_t nsplit, const void *tree, size_t tree_size, const void *const parts[4],
                                     const uint64_t *codes, size_t nparts, double mac_value, double G, double eps2,
                                     int offset_output, size_t ncrit, rk_eval_info *info, char *errbuf,
                                     size_t errbuf_len);

/* ---- device-resident kick-drift-kick integrator: the time loop of benchmark/benchmark_leapfrog.cpp:252-384 ------- */
/* Velocities (ORIGINAL particle order, host or device) are re-ordered into the tree's order (`reorder`, 252-267) and
 * kept on the device; the initial accelerations (and potentials when track_integrals) are computed (282). The tree must
 * have been built; theta / G / eps are used by every later evaluation. */
RK_API int rk_tree_leapfrog_init(rk_tree *t, const void *vx, const void *vy, const void *vz, int where, double theta,
                                 double G, double eps, int track_integrals);
/* One step: [conserved quantities 292-347] kick 349-356, drift = update_particles_u functor 359-370 + sync 3678-3743
 * (rebuild), accelerations 372, velocity update through last_perm 375-383. Nothing crosses PCIe. */
RK_API int rk_tree_leapfrog_step(rk_tree *t, double dt, rk_leapfrog_info *info);
/* what: 0 velocities, 1 accelerations of the last evaluation, 2 kicked velocities of the last step (in the order before
 * that step's rebuild), 3 potentials in `a` (track_integrals); tree order; NULL pointers are skipped. */
RK_API int rk_tree_leapfrog_get(rk_tree *t, int what, void *a, void *b, void *c, int where);

/* ---- synthetic inputs of the reference's benchmarks (host only) ------------------------------------------ */
/* Plummer sphere of benchmark/common.hpp:39-126. mode 0: sequential branch (first = 0, count = n_total);
 * mode 1: chunked deterministic form of the parallel branch (chunk seeded with its first index); `first` must be
 * a multiple of `chunk`, so ranks can generate disjoint shards [first, first + count) of one global stream. */
RK_API int rk_plummer(int fp_bits, size_t n_total, size_t first, size_t count, double a, double size, int mode,
                      size_t chunk, int nthreads, void *m, void *x, void *y, void *z);

/* Plummer positions and velocities of benchmark/benchmark_leapfrog.cpp:49-102, clipped at 10 core radii as at
 * 191-215 (G = M = 1). Arrays hold n elements; *kept receives the number of particles that survive the clip. */
RK_API int rk_plummer_leapfrog(int fp_bits, size_t n, double a, void *x, void *y, void *z, void *vx, void *vy,
                               void *vz, size_t *kept);

/* ---- measurement helpers (bench.py) ------------------------------------------------------------------- */
/* Kernels launched by this library since it was loaded. */
RK_API unsigned long long rk_kernel_launch_count(void);
/* FFMA microbenchmark on `device`: measured FP32-pipe peak in TFLOP/s (2 flop per FFMA), the denominator of the
 * traversal roofline (do not assume the nominal clock). */
RK_API int rk_measure_fp32_peak(int device, double *tflops, double *ms);
/* The same with DFMA: the FP64-pipe peak (config 3). */
RK_API int rk_measure_fp64_peak(int device, double *tflops, double *ms);

#ifdef __cplusplus
}
#endif
#endif
