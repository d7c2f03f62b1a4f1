// capi.cu — host side of librakau_b200.so: the device-resident tree object and the extern "C" boundary
// declared in include/rakau_b200.h. Host logic mirrors the reference's construct_impl (tree.hpp:1329-1487),
// sync (3678-3743), update_masses_dispatch (3782-3805), acc_pot_dispatch (3293-3334) and acc_pot_impl's
// argument validation (2857-2868, 3134-3141), including the exception messages its tests look for.
// There is no CPU fallback: every compute entry point needs a CUDA device.

#include "../../include/rakau_b200.h"
#include "common.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

namespace rk
{

unsigned long long g_kernel_launches = 0;

struct api_error : std::runtime_error {
    int status;
    api_error(int st, const std::string &s) : std::runtime_error(s), status(st) {}
};

template <typename F>
struct host_node_t {
    u64 begin, end, n_children, code, level;
    F props[4], dim, delta;
};
static_assert(sizeof(host_node_t<float>) == sizeof(rk_node_f32), "node layout");
static_assert(sizeof(host_node_t<double>) == sizeof(rk_node_f64), "node layout");

struct timer_events {
    cudaEvent_t ev[8] = {};
    void init()
    {
        for (auto &e : ev) {
            RK_CUDA_CHECK(cudaEventCreate(&e));
        }
    }
    void destroy()
    {
        for (auto &e : ev) {
            if (e) {
                cudaEventDestroy(e);
            }
        }
    }
};

template <typename F>
class tree
{
public:
    tree(int mac, int device) : m_mac(mac), m_device(device)
    {
        RK_CUDA_CHECK(cudaSetDevice(device));
        cudaDeviceProp prop;
        RK_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        m_sm_count = prop.multiProcessorCount;
        RK_CUDA_CHECK(cudaStreamCreateWithFlags(&m_own_stream, cudaStreamNonBlocking));
        m_stream = m_own_stream;
        m_ev.init();
        RK_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void **>(&m_hpin), 192 * sizeof(u64))); // [40..46]: leapfrog integrals, [48..175]: bucket counts
        m_b.d_err.reserve(2);
        m_b.d_misc.reserve(8);
        m_counters.reserve(40);
        m_work.reserve(8);
    }
    ~tree()
    {
        cudaSetDevice(m_device);
        cudaStreamSynchronize(m_stream);
        m_ev.destroy();
        if (m_hpin) {
            cudaFreeHost(m_hpin);
        }
        if (m_sc.h_ghist) {
            cudaFreeHost(m_sc.h_ghist);
        }
        for (auto &e : m_lf_ev) {
            if (e) {
                cudaEventDestroy(e);
            }
        }
        if (m_own_stream) {
            cudaStreamDestroy(m_own_stream);
        }
        if (m_copy_stream) {
            cudaStreamDestroy(m_copy_stream);
            for (auto &s : m_aux_stream) {
                cudaStreamDestroy(s);
            }
            for (auto &e : m_chunk_ev) {
                cudaEventDestroy(e);
            }
            cudaEventDestroy(m_copy_ev);
        }
    }
    void use() const { RK_CUDA_CHECK(cudaSetDevice(m_device)); }
    // Any cudaStream_t is accepted, INCLUDING 0 (the legacy default stream, which is what torch's default stream
    // is); (void*)-1 reverts to the tree's own non-blocking stream.
    void set_stream(void *s)
    {
        m_stream = (s == reinterpret_cast<void *>(-1)) ? m_own_stream : static_cast<cudaStream_t>(s);
    }
    void synchronize()
    {
        use();
        RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
    }

    size_t nparts() const { return m_b.n; }
    size_t nnodes() const { return m_b.n_nodes; }
    size_t ncrit_nodes() const { return m_b.n_crit; }
    double box_size() const { return static_cast<double>(m_box); }

    void clear()
    {
        m_b.n = 0;
        m_b.n_nodes = 0;
        m_b.n_crit = 0;
        m_b.codes = nullptr;
        m_b.last_perm = nullptr;
        m_box = F(0);
        m_box_deduced = false;
        m_max_group = 0;
        m_costs_valid = false;
        m_cuts_valid = false;
        m_have_inv = false; // sort_shard / traverse_external write a new permutation
        ++m_epoch;
        m_lf_ready = false;
        m_h_crit_begin.clear();
    }

    // ---- construct_impl, tree.hpp:1329-1487 --------------------------------------------------------------
    void build(const void *x, const void *y, const void *z, const void *m, size_t n, int where, double box_size,
               bool deduce, size_t max_leaf_n, size_t ncrit, rk_build_info *info)
    {
        use();
        clear();
        const F bs = static_cast<F>(box_size);
        // parameter checks, tree.hpp:1350-1362
        if (!std::isfinite(bs) || bs < F(0)) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "The box size must be a finite non-negative value, but it is "
                                                         + std::to_string(bs) + " instead");
        }
        if (!max_leaf_n) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "The maximum number of particles per leaf must be nonzero");
        }
        if (!ncrit) {
            throw api_error(RK_ERR_INVALID_ARGUMENT,
                            "The critical number of particles for the vectorised computation of the "
                            "potentials/accelerations must be nonzero");
        }
        if (n > 0xfffffff0ull) {
            throw api_error(RK_ERR_OVERFLOW, "The number of particles (" + std::to_string(n)
                                                 + ") is too large, and it results in an overflow condition");
        }
        m_box = bs;
        m_box_deduced = deduce;
        m_max_leaf_n = max_leaf_n;
        m_ncrit = ncrit;
        m_b.n = n;
        try {
            reserve_particles(n);
            RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[0], m_stream));
            const F *dx, *dy, *dz, *dm;
            m_late_m = nullptr;
            if (where == RK_HOST && m && n >= (size_t(1) << 20)) {
                // Host input: the coordinates are needed at once (box, codes), the masses only when the particles
                // are permuted, so their upload runs on the copy stream underneath the encode and the sort.
                upload4(x, y, z, nullptr, n, where, dx, dy, dz, dm);
                ensure_copy_stream();
                m_b.stage[3].reserve(n, 1.05);
                RK_CUDA_CHECK(cudaEventRecord(m_chunk_ev[0], m_stream)); // behind x, y, z on the link, not beside them
                RK_CUDA_CHECK(cudaStreamWaitEvent(m_copy_stream, m_chunk_ev[0], 0));
                RK_CUDA_CHECK(cudaMemcpyAsync(m_b.stage[3].p, m, n * sizeof(F), cudaMemcpyHostToDevice, m_copy_stream));
                RK_CUDA_CHECK(cudaEventRecord(m_copy_ev, m_copy_stream));
                m_late_m = m_b.stage[3].p;
            } else {
                upload4(x, y, z, m, n, where, dx, dy, dz, dm);
            }
            reset_flags();
            launch_pack_absmax<F>(dx, dy, dz, dm, m_b.pin.p, n, reinterpret_cast<u64 *>(m_b.d_misc.p), m_stream);
            rebuild(true, info);
        } catch (...) {
            if (m_late_m) { // do not leave an upload from the caller's buffer in flight
                cudaStreamSynchronize(m_copy_stream);
                m_late_m = nullptr;
            }
            clear();
            throw;
        }
    }

    // ---- copy constructor, tree.hpp:1735-1743: deep device-to-device copy -----------------------------------
    void clone_from(const tree &o)
    {
        use();
        clear();
        const size_t n = o.m_b.n, M = o.m_b.n_nodes, C = o.m_b.n_crit;
        m_box = o.m_box;
        m_box_deduced = o.m_box_deduced;
        m_max_leaf_n = o.m_max_leaf_n;
        m_ncrit = o.m_ncrit;
        m_max_group = o.m_max_group;
        m_b.n = n;
        m_b.n_nodes = M;
        m_b.n_crit = C;
        m_b.levels = o.m_b.levels;
        if (!n) {
            return;
        }
        RK_CUDA_CHECK(cudaStreamSynchronize(o.m_stream));
        reserve_particles(n);
        auto cp = [this](void *dst, const void *src, size_t bytes) {
            if (bytes) {
                RK_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, m_stream));
            }
        };
        cp(m_b.psorted.p, o.m_b.psorted.p, n * sizeof(vec4<F>));
        cp(m_b.keys_a.p, o.m_b.codes, n * sizeof(u64));
        cp(m_b.idx_a.p, o.m_b.last_perm, n * sizeof(u32));
        m_b.codes = m_b.keys_a.p;
        m_b.last_perm = m_b.idx_a.p;
        cp(m_b.perm.p, o.m_b.perm.p, n * sizeof(u32));
        m_have_inv = false;
        m_b.nodeA.reserve(M, 1.1);
        m_b.nodeB.reserve(M, 1.1);
        m_b.node_dfs.reserve(M, 1.1);
        m_b.node_ndesc.reserve(M, 1.1);
        m_b.crit_node.reserve(C, 1.1);
        m_b.crit_begin.reserve(C + 1, 1.1);
        cp(m_b.nodeA.p, o.m_b.nodeA.p, M * sizeof(vec4<F>));
        cp(m_b.nodeB.p, o.m_b.nodeB.p, M * sizeof(uint4));
        cp(m_b.node_dfs.p, o.m_b.node_dfs.p, M * sizeof(u32));
        cp(m_b.node_ndesc.p, o.m_b.node_ndesc.p, M * sizeof(u32));
        if (m_mac == RK_MAC_BH_GEOM) {
            m_b.node_delta.reserve(M, 1.1);
            cp(m_b.node_delta.p, o.m_b.node_delta.p, M * sizeof(F));
        }
        cp(m_b.crit_node.p, o.m_b.crit_node.p, C * sizeof(u32));
        cp(m_b.crit_begin.p, o.m_b.crit_begin.p, (C + 1) * sizeof(u32));
        RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
    }


    // ---- multi-GPU building blocks (one process per GPU) --------------------------------------
    // Sort a shard: pack, Morton-encode with the GLOBAL box (or take the given codes), stable radix sort, gather.
    // Afterwards codes / particles / last_perm of the shard are available on the device; no tree is built.
    void sort_shard(const void *x, const void *y, const void *z, const void *m, const uint64_t *codes_dev, size_t n,
                    double box_size)
    {
        use();
        clear();
        const F bs = static_cast<F>(box_size);
        if (!std::isfinite(bs) || bs <= F(0)) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "rk_tree_sort_shard needs the global box size");
        }
        m_box = bs;
        m_b.n = n;
        if (!n) {
            return;
        }
        try {
            reserve_particles(n);
            reset_flags();
            launch_pack_absmax<F>(static_cast<const F *>(x), static_cast<const F *>(y), static_cast<const F *>(z),
                                  static_cast<const F *>(m), m_b.pin.p, n, reinterpret_cast<u64 *>(m_b.d_misc.p), m_stream);
            const F inv_box = F(1) / m_box;
            if (codes_dev) {
                RK_CUDA_CHECK(cudaMemcpyAsync(m_b.keys_a.p, codes_dev, n * sizeof(u64), cudaMemcpyDeviceToDevice, m_stream));
            } else {
                launch_encode<F>(m_b.pin.p, m_b.keys_a.p, n, inv_box, m_b.d_err.p, m_stream);
            }
            u64 *codes;
            u32 *lperm;
            radix_sort_pairs(m_b.keys_a.p, m_b.keys_b.p, m_b.idx_a.p, m_b.idx_b.p, n, m_sc, m_stream, &codes, &lperm);
            m_b.codes = codes;
            m_b.last_perm = lperm;
            launch_gather<F>(m_b.pin.p, lperm, m_b.psorted.p, n, m_stream);
            launch_perm_first(lperm, m_b.perm.p, nullptr, n, m_stream);
            RK_CUDA_CHECK(cudaMemcpyAsync(m_hpin, m_b.d_err.p, 2 * sizeof(u64), cudaMemcpyDeviceToHost, m_stream));
            RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
            if (!codes_dev) {
                check_encode_error(m_hpin[0], inv_box);
            }
        } catch (...) {
            clear();
            throw;
        }
    }
    // Sample-sort without the local pre-sort: (1) encode_shard packs and Morton-encodes the shard (codes readable
    // with get_codes_device, in input order); the ranks agree on the splitters from samples of those codes;
    // (2) partition_shard moves the particles of each splitter bucket together with ONE stable radix pass on the
    // bucket id (instead of the eight passes of a full sort): bucket r is then a contiguous slice whose particles keep
    // their input order, so the receiving rank's stable sort of (runs in rank order) reproduces the single-GPU order.
    void encode_shard(const void *x, const void *y, const void *z, const void *m, size_t n, double box_size)
    {
        use();
        clear();
        const F bs = static_cast<F>(box_size);
        if (!std::isfinite(bs) || bs <= F(0)) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "rk_tree_encode_shard needs the global box size");
        }
        m_box = bs;
        m_b.n = n;
        if (!n) {
            return;
        }
        try {
            reserve_particles(n);
            reset_flags();
            launch_pack_absmax<F>(static_cast<const F *>(x), static_cast<const F *>(y), static_cast<const F *>(z),
                                  static_cast<const F *>(m), m_b.pin.p, n, reinterpret_cast<u64 *>(m_b.d_misc.p), m_stream);
            launch_encode<F>(m_b.pin.p, m_b.keys_a.p, n, F(1) / m_box, m_b.d_err.p, m_stream);
            m_b.codes = m_b.keys_a.p;
            RK_CUDA_CHECK(cudaMemcpyAsync(m_hpin, m_b.d_err.p, 2 * sizeof(u64), cudaMemcpyDeviceToHost, m_stream));
            RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
            check_encode_error(m_hpin[0], F(1) / m_box);
        } catch (...) {
            clear();
            throw;
        }
    }
    void partition_shard(const uint64_t *splitters_dev, unsigned nsplit, uint64_t *counts_out)
    {
        use();
        const size_t n = m_b.n;
        if (nsplit > 255u || !counts_out) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "rk_tree_partition_shard: at most 255 splitters");
        }
        for (unsigned r = 0; r <= nsplit; ++r) {
            counts_out[r] = 0;
        }
        if (!n) {
            return;
        }
        if (m_b.codes != m_b.keys_a.p) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "rk_tree_partition_shard: call rk_tree_encode_shard first");
        }
        try {
            // bucket ids -> keys_b; one stable pass: ids (unused afterwards) -> perm_tmp-sized scratch, permutation -> idx_a
            launch_bucket_ids(m_b.keys_a.p, n, reinterpret_cast<const u64 *>(splitters_dev), nsplit, m_b.keys_b.p, m_stream);
            m_ids_sorted.reserve(n, 1.05);
            const u32 *counts_dev = nullptr;
            radix_partition_pass(m_b.keys_b.p, m_ids_sorted.p, m_b.idx_a.p, n, m_sc, m_stream, &counts_dev);
            // codes and particles in bucket order
            launch_gather_u64(m_b.keys_a.p, m_b.idx_a.p, m_b.keys_b.p, n, m_stream);
            m_b.codes = m_b.keys_b.p;
            m_b.last_perm = m_b.idx_a.p;
            launch_gather<F>(m_b.pin.p, m_b.idx_a.p, m_b.psorted.p, n, m_stream);
            launch_perm_first(m_b.idx_a.p, m_b.perm.p, nullptr, n, m_stream);
            u32 *h = reinterpret_cast<u32 *>(m_hpin + 48);
            RK_CUDA_CHECK(cudaMemcpyAsync(h, counts_dev, 256 * sizeof(u32), cudaMemcpyDeviceToHost, m_stream));
            RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
            for (unsigned r = 0; r <= nsplit; ++r) {
                counts_out[r] = h[r];
            }
        } catch (...) {
            clear();
            throw;
        }
    }
    void get_codes_device(uint64_t *out)
    {
        use();
        if (m_b.n) {
            RK_CUDA_CHECK(cudaMemcpyAsync(out, m_b.codes, m_b.n * sizeof(u64), cudaMemcpyDeviceToDevice, m_stream));
        }
    }
    // Build the tree from globally sorted device arrays (SoA particles, codes, perm = original index of each
    // sorted particle): topology + node properties only.
    // parts_ready: optional cudaEvent_t after which x, y, z, m and perm are valid. The topology needs only the codes,
    // so a caller can all-gather the particle arrays on another stream while it is being built.
    void build_presorted(const void *x, const void *y, const void *z, const void *m, const uint64_t *codes_dev,
                         const uint32_t *perm_dev, size_t n, double box_size, size_t max_leaf_n, size_t ncrit,
                         void *parts_ready, rk_build_info *info)
    {
        use();
        clear();
        const F bs = static_cast<F>(box_size);
        if (!std::isfinite(bs) || bs <= F(0) || !max_leaf_n || !ncrit) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "rk_tree_build_presorted: invalid arguments");
        }
        m_box = bs;
        m_box_deduced = false;
        m_max_leaf_n = max_leaf_n;
        m_ncrit = ncrit;
        m_b.n = n;
        try {
            reserve_particles(n);
            RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[0], m_stream));
            reset_flags();
            RK_CUDA_CHECK(cudaMemcpyAsync(m_b.keys_a.p, codes_dev, n * sizeof(u64), cudaMemcpyDeviceToDevice, m_stream));
            m_b.codes = m_b.keys_a.p;
            m_b.last_perm = m_b.idx_a.p;
            for (int k = 1; k <= 3; ++k) {
                RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[k], m_stream));
            }
            finish_build(0, info, [&] {
                if (parts_ready) {
                    RK_CUDA_CHECK(cudaStreamWaitEvent(m_stream, static_cast<cudaEvent_t>(parts_ready), 0));
                }
                launch_pack_absmax<F>(static_cast<const F *>(x), static_cast<const F *>(y), static_cast<const F *>(z),
                                      static_cast<const F *>(m), m_b.psorted.p, n,
                                      reinterpret_cast<u64 *>(m_b.d_misc.p), m_stream);
                RK_CUDA_CHECK(
                    cudaMemcpyAsync(m_b.perm.p, perm_dev, n * sizeof(u32), cudaMemcpyDeviceToDevice, m_stream));
                RK_CUDA_CHECK(
                    cudaMemcpyAsync(m_b.idx_a.p, perm_dev, n * sizeof(u32), cudaMemcpyDeviceToDevice, m_stream));
            });
        } catch (...) {
            clear();
            throw;
        }
    }

    // ---- sync(), tree.hpp:3678-3743 ------------------------------------------------------------------------
    void update_positions(const void *x, const void *y, const void *z, const void *m, int where, rk_build_info *info)
    {
        use();
        const size_t n = m_b.n;
        try {
            RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[0], m_stream));
            const F *dx, *dy, *dz, *dm;
            upload4(x, y, z, m, n, where, dx, dy, dz, dm);
            reset_flags();
            // the current Morton order becomes the pre-sort order of the new build
            std::swap(m_b.pin.p, m_b.psorted.p);
            std::swap(m_b.pin.cap, m_b.psorted.cap);
            launch_set_coords<F>(m_b.pin.p, dx, dy, dz, dm, n, reinterpret_cast<u64 *>(m_b.d_misc.p), m_stream);
            rebuild(false, info);
        } catch (...) {
            clear();
            throw;
        }
    }

    // ---- update_masses_dispatch, tree.hpp:3782-3805 --------------------------------------------------------
    void update_masses(const void *m, int where)
    {
        use();
        const size_t n = m_b.n;
        try {
            const F *dx, *dy, *dz, *dm;
            upload4(nullptr, nullptr, nullptr, m, n, where, dx, dy, dz, dm);
            reset_flags();
            launch_set_coords<F>(m_b.psorted.p, nullptr, nullptr, nullptr, dm, n, nullptr, m_stream);
            node_properties<F>(m_b, m_mac, m_box, m_stream);
            RK_CUDA_CHECK(cudaMemcpyAsync(m_hpin, m_b.d_err.p, 2 * sizeof(u64), cudaMemcpyDeviceToHost, m_stream));
            RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
            check_props_error(m_hpin[1]);
            m_costs_valid = false;
            ++m_epoch;
        } catch (...) {
            clear();
            throw;
        }
    }

    // ---- getters ---------------------------------------------------------------------------------------------
    void get_parts(void *x, void *y, void *z, void *m)
    {
        use();
        const size_t n = m_b.n;
        if (!n) {
            return;
        }
        void *outs[4] = {x, y, z, m};
        F *d[4];
        for (int j = 0; j < 4; ++j) {
            m_b.stage[j].reserve(n, 1.05);
            d[j] = outs[j] ? m_b.stage[j].p : nullptr;
        }
        launch_unpack<F>(m_b.psorted.p, d[0], d[1], d[2], d[3], n, m_stream);
        for (int j = 0; j < 4; ++j) {
            if (outs[j]) {
                RK_CUDA_CHECK(cudaMemcpyAsync(outs[j], d[j], n * sizeof(F), cudaMemcpyDeviceToHost, m_stream));
            }
        }
        RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
    }
    void get_parts_device(void *x, void *y, void *z, void *m)
    {
        use();
        launch_unpack<F>(m_b.psorted.p, static_cast<F *>(x), static_cast<F *>(y), static_cast<F *>(z), static_cast<F *>(m),
                         m_b.n, m_stream);
        RK_CUDA_CHECK(cudaGetLastError());
    }
    void get_perm_device(int which, uint32_t *out)
    {
        use();
        const u32 *src = which == RK_PERM ? m_b.perm.p : (which == RK_LAST_PERM ? m_b.last_perm : inv_perm_dev());
        if (m_b.n) {
            RK_CUDA_CHECK(cudaMemcpyAsync(out, src, m_b.n * sizeof(u32), cudaMemcpyDeviceToDevice, m_stream));
        }
    }
    void get_codes(uint64_t *codes)
    {
        use();
        if (m_b.n) {
            RK_CUDA_CHECK(cudaMemcpyAsync(codes, m_b.codes, m_b.n * sizeof(u64), cudaMemcpyDeviceToHost, m_stream));
            RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
        }
    }
    // inv_perm is materialised on first use
    const u32 *inv_perm_dev()
    {
        if (!m_have_inv && m_b.n) {
            launch_perm_invert(m_b.perm.p, m_b.inv_perm.p, m_b.n, m_stream);
            m_have_inv = true;
        }
        return m_b.inv_perm.p;
    }
    void get_perm(int which, uint64_t *out)
    {
        use();
        const size_t n = m_b.n;
        if (!n) {
            return;
        }
        const u32 *src = which == RK_PERM ? m_b.perm.p : (which == RK_LAST_PERM ? m_b.last_perm : inv_perm_dev());
        std::vector<u32> tmp(n);
        RK_CUDA_CHECK(cudaMemcpyAsync(tmp.data(), src, n * sizeof(u32), cudaMemcpyDeviceToHost, m_stream));
        RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
        for (size_t i = 0; i < n; ++i) {
            out[i] = tmp[i];
        }
    }
    void get_nodes(void *nodes)
    {
        use();
        const size_t M = m_b.n_nodes;
        if (!M) {
            return;
        }
        dbuf<host_node_t<F>> tmp;
        tmp.reserve(M);
        launch_export_nodes<F>(m_b, m_mac, m_box, tmp.p, m_stream);
        RK_CUDA_CHECK(cudaMemcpyAsync(nodes, tmp.p, M * sizeof(host_node_t<F>), cudaMemcpyDeviceToHost, m_stream));
        RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
    }
    void get_crit(rk_cnode *crit)
    {
        use();
        const size_t C = m_b.n_crit;
        if (!C) {
            return;
        }
        dbuf<u64> tmp;
        tmp.reserve(3 * C);
        launch_export_crit(m_b.codes, m_b.nodeB.p, m_b.crit_node.p, m_b.crit_begin.p, C, tmp.p, m_stream);
        RK_CUDA_CHECK(cudaMemcpyAsync(crit, tmp.p, 3 * C * sizeof(u64), cudaMemcpyDeviceToHost, m_stream));
        RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
    }
    // First particle of the given critical nodes (index n_crit gives nparts).
    void crit_begin_at(const size_t *idx, size_t k, uint64_t *out)
    {
        use();
        for (size_t i = 0; i < k; ++i) {
            if (idx[i] > m_b.n_crit) {
                throw api_error(RK_ERR_INVALID_ARGUMENT, "critical node index out of range");
            }
        }
        if (k <= 64 && m_h_crit_begin.size() != m_b.n_crit + 1) {
            // a handful of cut points (multi-GPU range cuts, every evaluation): fetch just those entries instead of
            // mirroring the whole array (14 MB through pageable memory at 128 M particles)
            u32 *h = reinterpret_cast<u32 *>(m_hpin + 16);
            for (size_t i = 0; i < k; ++i) {
                RK_CUDA_CHECK(cudaMemcpyAsync(h + i, m_b.crit_begin.p + idx[i], sizeof(u32), cudaMemcpyDeviceToHost, m_stream));
            }
            RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
            for (size_t i = 0; i < k; ++i) {
                out[i] = h[i];
            }
            return;
        }
        host_crit_begin();
        for (size_t i = 0; i < k; ++i) {
            out[i] = m_h_crit_begin[idx[i]];
        }
    }
    const void *group_costs_device() const { return m_costs_valid ? m_group_cost.p : nullptr; }
    void get_group_costs(uint64_t *costs)
    {
        use();
        if (!m_costs_valid) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "No evaluation has been run on this tree yet");
        }
        if (m_b.n_crit) {
            RK_CUDA_CHECK(
                cudaMemcpyAsync(costs, m_group_cost.p, m_b.n_crit * sizeof(u64), cudaMemcpyDeviceToHost, m_stream));
            RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
        }
    }

    // ---- acc_pot_dispatch (tree.hpp:3293-3334) + acc_pot_impl (2853-3265) ----------------------------------
    void acc_pot(int Q, bool ordered, double theta_d, double G_d, double eps_d, const double *split, size_t nsplit,
                 bool ranged, size_t c0, size_t c1, void *const out[4], int where, rk_eval_info *info)
    {
        use();
        if (Q < 0 || Q > 2) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "Invalid value for Q");
        }
        const F theta = static_cast<F>(theta_d), G = static_cast<F>(G_d), eps = static_cast<F>(eps_d);
        if (!std::isfinite(theta) || theta <= F(0)) {
            throw api_error(RK_ERR_DOMAIN, "The MAC value must be finite and positive, but it is "
                                               + std::to_string(theta) + " instead");
        }
        const F mac_value = (m_mac == RK_MAC_BH) ? F(1) / (theta * theta) : F(1) / theta;
        if (!std::isfinite(mac_value) || mac_value <= F(0)) {
            throw api_error(RK_ERR_DOMAIN, "The transformed MAC value must be finite and positive, but it is "
                                               + std::to_string(mac_value) + " instead");
        }
        // compute_eps2 / check_G_const, tree.hpp:3268-3289
        if (!std::isfinite(eps) || eps < F(0)) {
            throw api_error(RK_ERR_DOMAIN, "The softening length must be finite and non-negative, but it is "
                                               + std::to_string(eps) + " instead");
        }
        const F eps2 = eps * eps;
        if (!std::isfinite(eps2) || eps2 < F(0)) {
            throw api_error(RK_ERR_DOMAIN,
                            "The square of the softening length must be finite and non-negative, but it is "
                                + std::to_string(eps2) + " instead");
        }
        if (!std::isfinite(G)) {
            throw api_error(RK_ERR_DOMAIN, "The value of the gravitational constant G must be finite, but it is "
                                               + std::to_string(G) + " instead");
        }
        // split validation, tree.hpp:2857-2868 and 3134-3141. There is no CPU path here: every share runs on
        // this tree's GPU (documented deviation).
        for (size_t i = 0; i < nsplit; ++i) {
            if (!std::isfinite(split[i])) {
                throw api_error(RK_ERR_INVALID_ARGUMENT, "The 'split' parameter cannot contain non-finite values");
            }
        }
        for (size_t i = 0; i < nsplit; ++i) {
            if (split[i] < 0.) {
                throw api_error(RK_ERR_INVALID_ARGUMENT,
                                "The 'split' parameter must contain only non-negative values");
            }
        }
        if (nsplit && std::all_of(split, split + nsplit, [](double v) { return v == 0.; })) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "The values in the 'split' parameter cannot all be zero");
        }
        if (nsplit) {
            const unsigned ndev = rk_device_count();
            if (nsplit - 1u > ndev) {
                throw api_error(RK_ERR_INVALID_ARGUMENT,
                                "Cannot split the computation of accelerations/potentials: the split vector refers to "
                                    + std::to_string(nsplit - 1u) + " accelerators, but only " + std::to_string(ndev)
                                    + " were detected");
            }
        }
        const size_t n = m_b.n, C = m_b.n_crit;
        if (!ranged) {
            c0 = 0;
            c1 = C;
        }
        if (c0 > c1 || c1 > C) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "Invalid range of critical nodes");
        }
        const int nres = Q == 0 ? 3 : (Q == 1 ? 1 : 4);
        if (info) {
            std::memset(info, 0, sizeof(*info));
        }
        if (!n || c0 == c1) {
            return;
        }
        for (int j = 0; j < nres; ++j) {
            if (!out[j]) {
                throw api_error(RK_ERR_INVALID_ARGUMENT, "Null output pointer");
            }
        }
        // split = {cpu, gpu0, gpu1, ...} with two or more accelerator shares: one process drives several devices, as
        // the reference does (tree.hpp:3147-3198, src/rakau_cuda.cu:492-527)
        if (!ranged && nsplit >= 3 && std::count_if(split + 2, split + nsplit, [](double v) { return v > 0.; }) > 0) {
            acc_pot_multi(Q, ordered, mac_value, G, eps2, split, nsplit, out, where, info);
            return;
        }

        trav_params<F> p{};
        p.parts = m_b.psorted.p;
        p.nodeA = m_b.nodeA.p;
        p.nodeB = m_b.nodeB.p;
        p.node_delta = m_b.node_delta.p;
        p.crit_node = m_b.crit_node.p;
        p.crit_begin = m_b.crit_begin.p;
        p.c0 = static_cast<u32>(c0);
        p.c1 = static_cast<u32>(c1);
        p.ncrit = static_cast<u32>(C);
        p.work_counter = m_work.p;
        for (int l = 0; l < NLEVELS; ++l) {
            const F nd = m_box / static_cast<F>(u64(1) << l); // get_node_dim, tree.hpp:443-448
            p.mac_tab[l] = (m_mac == RK_MAC_BH) ? (nd * nd) * mac_value : nd;
        }
        p.mac_value = mac_value;
        p.eps2 = eps2;
        p.G = G;
        p.perm = ordered ? m_b.perm.p : nullptr;
        if (!m_costs_valid) {
            m_group_cost.reserve(C, 1.1);
            RK_CUDA_CHECK(cudaMemsetAsync(m_group_cost.p, 0, C * sizeof(u64), m_stream));
        }
        p.group_cost = m_group_cost.p;
        p.counters = reinterpret_cast<u64 *>(m_counters.p);
        u32 tmax = static_cast<u32>((std::min<size_t>(m_max_group, 256) + 31) / 32 * 32);
        p.tmax = tmax ? tmax : 32;
        p.err = m_work.p + 1;
        p.out_offset = 0;
        p.window = trav_window(p.tmax, m_max_group);
        if (p.window) {
            // tail work stealing (traverse.cu): records + frontiers of the last wave of runs
            m_steal.reserve(size_t(TRAV_STEAL_SLOTS) * 16 + 16 + size_t(TRAV_STEAL_SLOTS) * trav_stack_cap());
            p.steal = m_steal.p;
            p.steal_published = m_steal.p + size_t(TRAV_STEAL_SLOTS) * 16;
            p.steal_front = p.steal_published + 16;
            p.steal_k = trav_steal_k();
            RK_CUDA_CHECK(cudaMemsetAsync(m_steal.p, 0, (size_t(TRAV_STEAL_SLOTS) * 16 + 16) * sizeof(u32), m_stream));
        }
        for (int j = 0; j < nres; ++j) {
            if (where == RK_DEVICE) {
                p.out[j] = static_cast<F *>(out[j]);
            } else {
                m_out[j].reserve(n, 1.05);
                p.out[j] = m_out[j].p;
            }
        }
        if (where == RK_DEVICE && m_n_mirror) {
            for (unsigned r = 0; r < m_n_mirror; ++r) {
                for (int j = 0; j < nres; ++j) {
                    if (!m_out_mirror[r][j]) {
                        throw api_error(RK_ERR_INVALID_ARGUMENT, "Null output mirror pointer");
                    }
                    p.mirror[r][j] = m_out_mirror[r][j];
                }
            }
            p.n_mirror = m_n_mirror;
            p.mirror_multicast = m_mirror_multicast;
        }
        RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[4], m_stream));
        RK_CUDA_CHECK(cudaMemsetAsync(m_work.p, 0, 8 * sizeof(u32), m_stream));
        RK_CUDA_CHECK(cudaMemsetAsync(m_counters.p, 0, 8 * sizeof(u64), m_stream));
        const bool partial = (c0 != 0 || c1 != C);
        if (partial && ordered && where == RK_HOST) {
            for (int j = 0; j < nres; ++j) {
                RK_CUDA_CHECK(cudaMemsetAsync(p.out[j], 0, n * sizeof(F), m_stream));
            }
        }
        RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[5], m_stream));
        // particle range covered by [c0, c1)
        size_t pb = 0, pe = n;
        if (partial && where == RK_HOST) {
            // only host outputs need the particle range; two entries through the pinned scratch, not the whole list
            const size_t idx[2] = {c0, c1};
            uint64_t be[2];
            crit_begin_at(idx, 2, be);
            pb = be[0];
            pe = be[1];
        }
        unsigned launches = 1;
        // Host outputs in Morton order that live in pinned (mapped) memory: the kernel writes the final results straight
        // into the caller's buffers - ONE launch with its work-stealing tail and no device-to-host copy behind it.
        bool zero_copy = where == RK_HOST && !ordered && m_zero_copy_out != 0;
        for (int j = 0; zero_copy && j < nres; ++j) {
            cudaPointerAttributes at{};
            if (cudaPointerGetAttributes(&at, out[j]) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
                cudaGetLastError();
                zero_copy = false;
            } else {
                p.outf[j] = static_cast<F *>(at.devicePointer);
            }
        }
        if (!zero_copy) {
            for (int j = 0; j < 4; ++j) {
                p.outf[j] = nullptr;
            }
        }
        const bool pipelined = !zero_copy && where == RK_HOST && !ordered && (pe - pb) >= (size_t(1) << 20);
        if (!pipelined) {
            launch_traverse<F>(p, Q, m_mac, m_sm_count, m_stream, m_kernel_name);
            RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[6], m_stream));
            if (where == RK_HOST && !zero_copy) {
                for (int j = 0; j < nres; ++j) {
                    if (ordered) {
                        RK_CUDA_CHECK(
                            cudaMemcpyAsync(out[j], p.out[j], n * sizeof(F), cudaMemcpyDeviceToHost, m_stream));
                    } else {
                        RK_CUDA_CHECK(cudaMemcpyAsync(static_cast<F *>(out[j]) + pb, p.out[j] + pb,
                                                      (pe - pb) * sizeof(F), cudaMemcpyDeviceToHost, m_stream));
                    }
                }
            }
        } else {
            // Host outputs in Morton order: cut the critical nodes into NCHUNK launches (40/30/20/10 %) on separate
            // streams - the CTAs of launch k+1 fill the SMs as those of launch k drain, so there is one tail, not
            // NCHUNK - and copy each launch's contiguous output range back on the copy stream as soon as that launch
            // has finished. Only the last tenth of the device-to-host traffic is exposed.
            ensure_copy_stream();
            const size_t nc = c1 - c0;
            size_t cuts[NCHUNK + 1], pcut[NCHUNK + 1];
            cuts[0] = c0, cuts[NCHUNK] = c1, pcut[0] = pb, pcut[NCHUNK] = pe;
            if (!partial && m_cuts_valid) {
                for (int k = 1; k < NCHUNK; ++k) {
                    cuts[k] = m_cut_crit[k - 1];
                    pcut[k] = m_cut_begin[k - 1];
                }
            } else {
                host_crit_begin();
                for (int k = 1; k < NCHUNK; ++k) {
                    cuts[k] = c0 + nc * chunk_mark(k) / 10;
                    pcut[k] = m_h_crit_begin[cuts[k]];
                }
            }
            launches = 0;
            for (int k = 0; k < NCHUNK; ++k) {
                if (cuts[k + 1] == cuts[k]) {
                    continue;
                }
                cudaStream_t st = k ? m_aux_stream[k - 1] : m_stream;
                if (k) {
                    RK_CUDA_CHECK(cudaStreamWaitEvent(st, m_ev.ev[5], 0));
                }
                trav_params<F> q = p;
                q.c0 = static_cast<u32>(cuts[k]);
                q.c1 = static_cast<u32>(cuts[k + 1]);
                q.work_counter = m_work.p + 4 + k; // one work counter per launch
                if (k != NCHUNK - 1) {
                    q.steal = nullptr; // the next launch fills the SMs this one leaves: only the last one has a tail
                }
                launch_traverse<F>(q, Q, m_mac, m_sm_count, st, m_kernel_name);
                RK_CUDA_CHECK(cudaEventRecord(m_chunk_ev[k], st));
                ++launches;
            }
            for (int k = 0; k < NCHUNK; ++k) {
                if (cuts[k + 1] == cuts[k]) {
                    continue;
                }
                if (k) {
                    RK_CUDA_CHECK(cudaStreamWaitEvent(m_stream, m_chunk_ev[k], 0));
                }
                const size_t b = pcut[k], e = pcut[k + 1];
                RK_CUDA_CHECK(cudaStreamWaitEvent(m_copy_stream, m_chunk_ev[k], 0));
                for (int j = 0; j < nres; ++j) {
                    RK_CUDA_CHECK(cudaMemcpyAsync(static_cast<F *>(out[j]) + b, p.out[j] + b, (e - b) * sizeof(F),
                                                  cudaMemcpyDeviceToHost, m_copy_stream));
                }
            }
            RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[6], m_stream)); // all launches done
            RK_CUDA_CHECK(cudaEventRecord(m_copy_ev, m_copy_stream));
            RK_CUDA_CHECK(cudaStreamWaitEvent(m_stream, m_copy_ev, 0)); // the tree's stream sees the copies
        }
        RK_CUDA_CHECK(cudaMemcpyAsync(m_hpin, m_counters.p, 8 * sizeof(u64), cudaMemcpyDeviceToHost, m_stream));
        RK_CUDA_CHECK(cudaMemcpyAsync(m_hpin + 8, m_work.p, 4 * sizeof(u32), cudaMemcpyDeviceToHost, m_stream));
        RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[7], m_stream));
        RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
        const u32 *hw = reinterpret_cast<const u32 *>(m_hpin + 8);
        if (hw[1]) {
            throw api_error(RK_ERR_RUNTIME, "Traversal stack overflow in the CUDA kernel");
        }
        m_costs_valid = true; // groups outside the evaluated ranges hold 0
        if (info) {
            info->mac_tests = m_hpin[0];
            info->accepted = m_hpin[1];
            info->p2p_pairs = m_hpin[2];
            info->self_pairs = m_hpin[3];
            info->n_groups = c1 - c0;
            info->kernel_launches = launches;
            RK_CUDA_CHECK(cudaEventElapsedTime(&info->ms_kernel, m_ev.ev[5], m_ev.ev[6]));
            RK_CUDA_CHECK(cudaEventElapsedTime(&info->ms_total, m_ev.ev[4], m_ev.ev[7]));
            // interactions = sum over groups of T*(leaf sources + accepted) + T*(T-1)
            info->interactions = info->p2p_pairs + 2 * info->self_pairs + m_hpin[4]; // [4] = sum T * accepted
        }
    }


    // ---- one process, several devices: the reference's `split` kwarg (tree.hpp:3147-3198) ---------------------------
    // split[0] is the CPU share of the reference; there is no CPU path here, so it is evaluated by the first
    // accelerator together with split[1]. Accelerator j >= 1 is device (this tree's device + j) mod device count: it
    // holds a mirror of the traversal arrays (copied over NVLink when the tree has changed), evaluates its contiguous
    // Morton range of critical nodes - the cuts are the reference's: cumulative weights projected onto the particle
    // indices and snapped to the next critical node - and its results are copied back into this device's output
    // arrays. Runs are cut-independent (traverse.cu), so the result equals the one-device result bit for bit.
    void acc_pot_multi(int Q, bool ordered, F mac_value, F G, F eps2, const double *split, size_t nsplit,
                       void *const out[4], int where, rk_eval_info *info)
    {
        const size_t n = m_b.n, C = m_b.n_crit, nacc = nsplit - 1;
        const int nres = Q == 0 ? 3 : (Q == 1 ? 1 : 4);
        int ndev = 0;
        RK_CUDA_CHECK(cudaGetDeviceCount(&ndev));
        // particle cuts -> critical-node cuts
        std::vector<double> share(nacc);
        share[0] = split[0] + split[1];
        for (size_t j = 1; j < nacc; ++j) {
            share[j] = split[j + 1];
        }
        const double total = std::accumulate(share.begin(), share.end(), 0.);
        std::vector<uint64_t> pidx(nacc - 1), ccut(nacc + 1);
        double run = 0;
        for (size_t j = 0; j + 1 < nacc; ++j) {
            run += share[j];
            pidx[j] = static_cast<uint64_t>(run / total * static_cast<double>(n));
        }
        ccut[0] = 0;
        ccut[nacc] = C;
        for (size_t j = 0; j + 1 < nacc; j += 16) {
            const size_t k = std::min<size_t>(16, nacc - 1 - j);
            crit_lower_bound(pidx.data() + j, k, ccut.data() + 1 + j);
        }
        for (size_t j = 1; j <= nacc; ++j) {
            ccut[j] = std::max(ccut[j], ccut[j - 1]);
        }
        std::vector<size_t> cidx(ccut.begin(), ccut.end());
        std::vector<uint64_t> pcut(nacc + 1);
        for (size_t j = 0; j <= nacc; j += 64) {
            crit_begin_at(cidx.data() + j, std::min<size_t>(64, nacc + 1 - j), pcut.data() + j);
        }
        use();
        for (int k = 0; k < nres; ++k) {
            m_out[k].reserve(n, 1.05); // the peers copy their ranges in, also when this device's own share is empty
        }
        // mirrors
        if (m_mirror.size() < nacc - 1) {
            m_mirror.resize(nacc - 1);
        }
        for (size_t j = 1; j < nacc; ++j) {
            if (ccut[j + 1] == ccut[j]) {
                continue;
            }
            auto &mp = m_mirror[j - 1];
            const int dev = (m_device + static_cast<int>(j)) % ndev;
            if (!mp || mp->m_device != dev) {
                mp.reset(new tree<F>(m_mac, dev));
            }
            mp->mirror_from(*this);
        }
        use();
        // launch every share (asynchronously), own share last so that the peers start first
        std::vector<rk_eval_info> infos(nacc);
        for (size_t j = nacc; j-- > 0;) {
            if (ccut[j + 1] == ccut[j]) {
                continue;
            }
            tree<F> &T = j ? *m_mirror[j - 1] : *this;
            T.launch_range(Q, mac_value, G, eps2, ccut[j], ccut[j + 1]);
        }
        // collect: peers copy their Morton range of the results into this device's arrays
        for (size_t j = 1; j < nacc; ++j) {
            if (ccut[j + 1] == ccut[j]) {
                continue;
            }
            tree<F> &T = *m_mirror[j - 1];
            T.use();
            const size_t b = pcut[j], e = pcut[j + 1];
            for (int k = 0; k < nres; ++k) {
                RK_CUDA_CHECK(cudaMemcpyPeerAsync(m_out[k].p + b, m_device, T.m_out[k].p + b, T.m_device,
                                                  (e - b) * sizeof(F), T.m_stream));
            }
            T.finish_range(&infos[j]);
        }
        use();
        if (ccut[1] != ccut[0]) {
            finish_range(&infos[0]);
        }
        // outputs: Morton order as computed, or scattered to the original order (tree.hpp:3320-3330)
        for (int k = 0; k < nres; ++k) {
            F *src = m_out[k].p;
            if (ordered) {
                m_out_ord[k].reserve(n, 1.05);
                launch_scatter_perm<F>(m_out[k].p, m_b.perm.p, m_out_ord[k].p, n, m_stream);
                src = m_out_ord[k].p;
            }
            RK_CUDA_CHECK(cudaMemcpyAsync(out[k], src, n * sizeof(F),
                                          where == RK_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, m_stream));
        }
        RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
        if (info) {
            for (const auto &ei : infos) {
                info->mac_tests += ei.mac_tests;
                info->accepted += ei.accepted;
                info->p2p_pairs += ei.p2p_pairs;
                info->self_pairs += ei.self_pairs;
                info->interactions += ei.interactions;
                info->n_groups += ei.n_groups;
                info->kernel_launches += ei.kernel_launches;
                info->ms_kernel = std::max(info->ms_kernel, ei.ms_kernel);
            }
            info->ms_total = info->ms_kernel;
        }
    }
    // Copy the traversal arrays of `src` (another device) into this tree when they have changed.
    void mirror_from(const tree &src)
    {
        if (m_mirror_of == &src && m_mirror_epoch == src.m_epoch) {
            return;
        }
        use();
        const size_t n = src.m_b.n, M = src.m_b.n_nodes, C = src.m_b.n_crit;
        m_b.n = n;
        m_b.n_nodes = M;
        m_b.n_crit = C;
        m_box = src.m_box;
        m_max_group = src.m_max_group;
        m_ncrit = src.m_ncrit;
        m_max_leaf_n = src.m_max_leaf_n;
        m_b.psorted.reserve(n, 1.05);
        m_b.nodeA.reserve(M, 1.1);
        m_b.nodeB.reserve(M, 1.1);
        m_b.crit_node.reserve(C, 1.1);
        m_b.crit_begin.reserve(C + 1, 1.1);
        RK_CUDA_CHECK(cudaStreamSynchronize(src.m_stream));
        auto cp = [&](void *dst, const void *from, size_t bytes) {
            if (bytes) {
                RK_CUDA_CHECK(cudaMemcpyPeerAsync(dst, m_device, from, src.m_device, bytes, m_stream));
            }
        };
        cp(m_b.psorted.p, src.m_b.psorted.p, n * sizeof(vec4<F>));
        cp(m_b.nodeA.p, src.m_b.nodeA.p, M * sizeof(vec4<F>));
        cp(m_b.nodeB.p, src.m_b.nodeB.p, M * sizeof(uint4));
        if (m_mac == RK_MAC_BH_GEOM) {
            m_b.node_delta.reserve(M, 1.1);
            cp(m_b.node_delta.p, src.m_b.node_delta.p, M * sizeof(F));
        }
        cp(m_b.crit_node.p, src.m_b.crit_node.p, C * sizeof(u32));
        cp(m_b.crit_begin.p, src.m_b.crit_begin.p, (C + 1) * sizeof(u32));
        m_costs_valid = false;
        m_mirror_of = &src;
        m_mirror_epoch = src.m_epoch;
    }
    // Asynchronous evaluation of the critical nodes [c0, c1) into this tree's own result arrays (Morton order);
    // finish_range() waits for it and returns the counters.
    void launch_range(int Q, F mac_value, F G, F eps2, size_t c0, size_t c1)
    {
        use();
        const size_t n = m_b.n, C = m_b.n_crit;
        trav_params<F> p{};
        p.parts = m_b.psorted.p;
        p.nodeA = m_b.nodeA.p;
        p.nodeB = m_b.nodeB.p;
        p.node_delta = m_b.node_delta.p;
        p.crit_node = m_b.crit_node.p;
        p.crit_begin = m_b.crit_begin.p;
        p.c0 = static_cast<u32>(c0);
        p.c1 = static_cast<u32>(c1);
        p.ncrit = static_cast<u32>(C);
        p.work_counter = m_work.p;
        for (int l = 0; l < NLEVELS; ++l) {
            const F nd = m_box / static_cast<F>(u64(1) << l);
            p.mac_tab[l] = (m_mac == RK_MAC_BH) ? (nd * nd) * mac_value : nd;
        }
        p.mac_value = mac_value;
        p.eps2 = eps2;
        p.G = G;
        p.perm = nullptr;
        if (!m_costs_valid) {
            m_group_cost.reserve(C, 1.1);
            RK_CUDA_CHECK(cudaMemsetAsync(m_group_cost.p, 0, C * sizeof(u64), m_stream));
        }
        p.group_cost = m_group_cost.p;
        p.counters = reinterpret_cast<u64 *>(m_counters.p);
        const u32 tmax = static_cast<u32>((std::min<size_t>(m_max_group, 256) + 31) / 32 * 32);
        p.tmax = tmax ? tmax : 32;
        p.err = m_work.p + 1;
        p.window = trav_window(p.tmax, m_max_group);
        if (p.window) {
            m_steal.reserve(size_t(TRAV_STEAL_SLOTS) * 16 + 16 + size_t(TRAV_STEAL_SLOTS) * trav_stack_cap());
            p.steal = m_steal.p;
            p.steal_published = m_steal.p + size_t(TRAV_STEAL_SLOTS) * 16;
            p.steal_front = p.steal_published + 16;
            p.steal_k = trav_steal_k();
            RK_CUDA_CHECK(cudaMemsetAsync(m_steal.p, 0, (size_t(TRAV_STEAL_SLOTS) * 16 + 16) * sizeof(u32), m_stream));
        }
        const int nres = Q == 0 ? 3 : (Q == 1 ? 1 : 4);
        for (int j = 0; j < nres; ++j) {
            m_out[j].reserve(n, 1.05);
            p.out[j] = m_out[j].p;
        }
        RK_CUDA_CHECK(cudaMemsetAsync(m_work.p, 0, 8 * sizeof(u32), m_stream));
        RK_CUDA_CHECK(cudaMemsetAsync(m_counters.p, 0, 8 * sizeof(u64), m_stream));
        RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[5], m_stream));
        launch_traverse<F>(p, Q, m_mac, m_sm_count, m_stream, m_kernel_name);
        RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[6], m_stream));
        m_range_groups = c1 - c0;
    }
    void finish_range(rk_eval_info *info)
    {
        use();
        RK_CUDA_CHECK(cudaMemcpyAsync(m_hpin, m_counters.p, 8 * sizeof(u64), cudaMemcpyDeviceToHost, m_stream));
        RK_CUDA_CHECK(cudaMemcpyAsync(m_hpin + 8, m_work.p, 4 * sizeof(u32), cudaMemcpyDeviceToHost, m_stream));
        RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
        if (reinterpret_cast<const u32 *>(m_hpin + 8)[1]) {
            throw api_error(RK_ERR_RUNTIME, "Traversal stack overflow in the CUDA kernel");
        }
        m_costs_valid = true;
        std::memset(info, 0, sizeof(*info));
        info->mac_tests = m_hpin[0];
        info->accepted = m_hpin[1];
        info->p2p_pairs = m_hpin[2];
        info->self_pairs = m_hpin[3];
        info->interactions = info->p2p_pairs + 2 * info->self_pairs + m_hpin[4];
        info->n_groups = m_range_groups;
        info->kernel_launches = 1;
        RK_CUDA_CHECK(cudaEventElapsedTime(&info->ms_kernel, m_ev.ev[5], m_ev.ev[6]));
        info->ms_total = info->ms_kernel;
    }

    // ---- literal drop-in for cuda_acc_pot_impl (rakau_cuda.cu:348-528): traverse a host-built tree ----------
    // The DFS AoS node array of the reference is re-laid level-major on the host (O(M) loops), uploaded together
    // with the particles, and walked by the same kernel as the device-built trees.
    void traverse_external(int Q, void *const out[4], const uint64_t *split_indices, size_t nsplit,
                           const host_node_t<F> *nodes, size_t M, const void *const parts[4], size_t n, F mac_value,
                           F G, F eps2, bool offset_output, size_t ncrit, rk_eval_info *info)
    {
        use();
        clear();
        if (info) {
            std::memset(info, 0, sizeof(*info));
        }
        if (Q < 0 || Q > 2 || !nsplit || !nodes || !M || !n) {
            if (!n || !M) {
                return;
            }
            throw api_error(RK_ERR_INVALID_ARGUMENT, "rk_traverse_external_tree: invalid arguments");
        }
        if (!ncrit) {
            ncrit = 128; // the reference's default_ncrit (tree.hpp:589-595)
        }
        if (n > 0xfffffff0ull || M > 0xfffffff0ull / 8) {
            throw api_error(RK_ERR_OVERFLOW, "rk_traverse_external_tree: the tree is too large");
        }
        const size_t first = static_cast<size_t>(split_indices[0]);
        if (first >= n) {
            return; // everything stays on the caller's CPU share
        }
        // level-major order
        u32 cnt[NLEVELS] = {};
        for (size_t k = 0; k < M; ++k) {
            if (nodes[k].level >= u64(NLEVELS)) {
                throw api_error(RK_ERR_INVALID_ARGUMENT, "rk_traverse_external_tree: node level out of range");
            }
            ++cnt[nodes[k].level];
        }
        u32 run[NLEVELS];
        u32 acc = 0;
        for (int l = 0; l < NLEVELS; ++l) {
            m_b.levels.base[l] = acc;
            run[l] = acc;
            acc += cnt[l];
        }
        m_b.levels.base[NLEVELS] = acc;
        m_b.levels.base[NLEVELS + 1] = acc;
        std::vector<u32> bfs(M);
        for (size_t k = 0; k < M; ++k) {
            bfs[k] = run[nodes[k].level]++;
        }
        std::vector<vec4<F>> hA(M);
        std::vector<uint4> hB(M);
        std::vector<F> hD(m_mac == RK_MAC_BH_GEOM ? M : 0);
        F tab[NLEVELS] = {};
        for (size_t k = 0; k < M; ++k) {
            const auto &nd = nodes[k];
            u32 nch = 0;
            for (size_t c = k + 1, e = k + nd.n_children; c <= e; c += 1 + nodes[c].n_children) {
                ++nch;
            }
            const u32 b = bfs[k];
            hA[b] = make_vec4<F>(nd.props[0], nd.props[1], nd.props[2], nd.props[3]);
            hB[b] = make_uint4(static_cast<u32>(nd.begin), static_cast<u32>(nd.end), nd.n_children ? bfs[k + 1] : 0u,
                               (static_cast<u32>(nd.level) << 8) | nch);
            if (m_mac == RK_MAC_BH_GEOM) {
                hD[b] = nd.delta;
            }
            tab[nd.level] = nd.dim; // dim2 (bh) or dim (bh_geom): a function of the level only
        }
        // critical nodes: first node on each root path with <= ncrit particles, or a leaf (tree.hpp:801-803)
        std::vector<u32> hcn, hcb;
        size_t max_group = 0;
        for (size_t k = 0; k < M;) {
            const auto &nd = nodes[k];
            const size_t np = nd.end - nd.begin;
            if (np <= ncrit || !nd.n_children) {
                hcn.push_back(bfs[k]);
                hcb.push_back(static_cast<u32>(nd.begin));
                max_group = std::max(max_group, np);
                k += 1 + nd.n_children;
            } else {
                ++k;
            }
        }
        const size_t C = hcn.size();
        hcb.push_back(static_cast<u32>(n));
        // first critical node of the GPU share: split_indices[0] was snapped to a node boundary by the caller
        const size_t c0 = std::lower_bound(hcb.begin(), hcb.begin() + C, static_cast<u32>(first)) - hcb.begin();
        if (c0 == C || hcb[c0] != first) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "rk_traverse_external_tree: split_indices[0] is not the first "
                                                     "particle of a critical node for the given ncrit");
        }
        // upload
        m_b.n = n;
        m_b.n_nodes = M;
        m_b.n_crit = C;
        m_max_group = max_group;
        reserve_particles(n);
        const F *dx, *dy, *dz, *dm;
        upload4(parts[0], parts[1], parts[2], parts[3], n, RK_HOST, dx, dy, dz, dm);
        reset_flags();
        launch_pack_absmax<F>(dx, dy, dz, dm, m_b.psorted.p, n, reinterpret_cast<u64 *>(m_b.d_misc.p), m_stream);
        m_b.nodeA.reserve(M, 1.0);
        m_b.nodeB.reserve(M, 1.0);
        m_b.crit_node.reserve(C, 1.0);
        m_b.crit_begin.reserve(C + 1, 1.0);
        auto up = [this](void *dst, const void *src, size_t bytes) {
            RK_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, m_stream));
        };
        up(m_b.nodeA.p, hA.data(), M * sizeof(vec4<F>));
        up(m_b.nodeB.p, hB.data(), M * sizeof(uint4));
        if (m_mac == RK_MAC_BH_GEOM) {
            m_b.node_delta.reserve(M, 1.0);
            up(m_b.node_delta.p, hD.data(), M * sizeof(F));
        }
        up(m_b.crit_node.p, hcn.data(), C * sizeof(u32));
        up(m_b.crit_begin.p, hcb.data(), (C + 1) * sizeof(u32));

        trav_params<F> p{};
        p.parts = m_b.psorted.p;
        p.nodeA = m_b.nodeA.p;
        p.nodeB = m_b.nodeB.p;
        p.node_delta = m_b.node_delta.p;
        p.crit_node = m_b.crit_node.p;
        p.crit_begin = m_b.crit_begin.p;
        p.c0 = static_cast<u32>(c0);
        p.c1 = static_cast<u32>(C);
        p.ncrit = static_cast<u32>(C);
        p.work_counter = m_work.p;
        for (int l = 0; l < NLEVELS; ++l) {
            p.mac_tab[l] = (m_mac == RK_MAC_BH) ? tab[l] * mac_value : tab[l];
        }
        p.mac_value = mac_value;
        p.eps2 = eps2;
        p.G = G;
        p.perm = nullptr;
        m_group_cost.reserve(C, 1.0);
        RK_CUDA_CHECK(cudaMemsetAsync(m_group_cost.p, 0, C * sizeof(u64), m_stream));
        p.group_cost = m_group_cost.p;
        p.counters = reinterpret_cast<u64 *>(m_counters.p);
        const u32 tmax = static_cast<u32>((std::min<size_t>(max_group, 256) + 31) / 32 * 32);
        p.tmax = tmax ? tmax : 32;
        p.err = m_work.p + 1;
        p.out_offset = static_cast<u32>(first);
        p.window = trav_window(p.tmax, max_group);
        const int nres = Q == 0 ? 3 : (Q == 1 ? 1 : 4);
        const size_t cnt_out = n - first;
        for (int j = 0; j < nres; ++j) {
            m_out[j].reserve(cnt_out, 1.0);
            p.out[j] = m_out[j].p;
        }
        RK_CUDA_CHECK(cudaMemsetAsync(m_work.p, 0, 4 * sizeof(u32), m_stream));
        RK_CUDA_CHECK(cudaMemsetAsync(m_counters.p, 0, 8 * sizeof(u64), m_stream));
        RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[5], m_stream));
        launch_traverse<F>(p, Q, m_mac, m_sm_count, m_stream, m_kernel_name);
        RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[6], m_stream));
        for (int j = 0; j < nres; ++j) {
            F *dst = static_cast<F *>(out[j]) + (offset_output ? first : 0);
            RK_CUDA_CHECK(cudaMemcpyAsync(dst, p.out[j], cnt_out * sizeof(F), cudaMemcpyDeviceToHost, m_stream));
        }
        RK_CUDA_CHECK(cudaMemcpyAsync(m_hpin, m_counters.p, 8 * sizeof(u64), cudaMemcpyDeviceToHost, m_stream));
        RK_CUDA_CHECK(cudaMemcpyAsync(m_hpin + 8, m_work.p, 4 * sizeof(u32), cudaMemcpyDeviceToHost, m_stream));
        RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
        if (reinterpret_cast<const u32 *>(m_hpin + 8)[1]) {
            throw api_error(RK_ERR_RUNTIME, "Traversal stack overflow in the CUDA kernel");
        }
        if (info) {
            info->mac_tests = m_hpin[0];
            info->accepted = m_hpin[1];
            info->p2p_pairs = m_hpin[2];
            info->self_pairs = m_hpin[3];
            info->interactions = info->p2p_pairs + 2 * info->self_pairs + m_hpin[4];
            info->n_groups = C - c0;
            info->kernel_launches = 1;
            RK_CUDA_CHECK(cudaEventElapsedTime(&info->ms_kernel, m_ev.ev[5], m_ev.ev[6]));
            info->ms_total = info->ms_kernel;
        }
    }

    void exact(size_t idx, bool ordered, double G, double eps_d, double out4[4])
    {
        use();
        const F eps = static_cast<F>(eps_d);
        if (!std::isfinite(eps) || eps < F(0)) {
            throw api_error(RK_ERR_DOMAIN, "The softening length must be finite and non-negative, but it is "
                                               + std::to_string(eps) + " instead");
        }
        const F eps2 = eps * eps;
        if (!std::isfinite(eps2) || eps2 < F(0)) {
            throw api_error(RK_ERR_DOMAIN,
                            "The square of the softening length must be finite and non-negative, but it is "
                                + std::to_string(eps2) + " instead");
        }
        if (!std::isfinite(static_cast<F>(G))) {
            throw api_error(RK_ERR_DOMAIN, "The value of the gravitational constant G must be finite, but it is "
                                               + std::to_string(static_cast<F>(G)) + " instead");
        }
        if (idx >= m_b.n) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "Particle index out of range");
        }
        if (ordered) {
            u32 v;
            RK_CUDA_CHECK(cudaMemcpyAsync(&v, inv_perm_dev() + idx, sizeof(u32), cudaMemcpyDeviceToHost, m_stream));
            RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
            idx = v;
        }
        double *d = reinterpret_cast<double *>(m_counters.p) + 4;
        launch_exact<F>(m_b.psorted.p, m_b.n, idx, static_cast<F>(G), eps2, d, m_stream);
        RK_CUDA_CHECK(cudaMemcpyAsync(m_hpin + 16, d, 4 * sizeof(double), cudaMemcpyDeviceToHost, m_stream));
        RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
        std::memcpy(out4, m_hpin + 16, 4 * sizeof(double));
    }

    // ---- device-resident leapfrog, benchmark_leapfrog.cpp:252-384 -------------------------------------------------
    // init: velocities given in the ORIGINAL particle order are re-ordered with the tree's permutation (the `reorder`
    // helper, 252-267) and the initial accelerations (+ potentials) are computed (282).
    void leapfrog_init(const void *vx, const void *vy, const void *vz, int where, double theta, double G, double eps,
                       bool track)
    {
        use();
        const size_t n = m_b.n;
        if (!n) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "rk_tree_leapfrog_init: the tree is empty");
        }
        for (int j = 0; j < 3; ++j) {
            m_lf_v[j].reserve(n, 1.05);
            m_lf_kv[j].reserve(n, 1.05);
        }
        for (int j = 0; j < 4; ++j) {
            m_lf_acc[j].reserve(n, 1.05);
        }
        m_lf_scratch.reserve(lf_scratch_doubles());
        if (!m_lf_ev[0]) {
            for (auto &e : m_lf_ev) {
                RK_CUDA_CHECK(cudaEventCreate(&e));
            }
        }
        const F *dx, *dy, *dz, *dm;
        upload4(vx, vy, vz, nullptr, n, where, dx, dy, dz, dm); // (staging buffers: free once the tree is built)
        const F *in[3] = {dx, dy, dz};
        F *out[3] = {m_lf_v[0].p, m_lf_v[1].p, m_lf_v[2].p};
        launch_lf_reorder<F>(in, m_b.perm.p, out, n, m_stream);
        m_lf_theta = theta;
        m_lf_G = G;
        m_lf_eps = eps;
        m_lf_track = track;
        m_lf_ready = true;
        leapfrog_eval(nullptr);
    }
    void leapfrog_step(double dt_d, rk_leapfrog_info *info)
    {
        use();
        if (!m_lf_ready || !m_b.n) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "rk_tree_leapfrog_step: call rk_tree_leapfrog_init first");
        }
        const size_t n = m_b.n;
        const F dt = static_cast<F>(dt_d), half_dt = dt / F(2);
        if (info) {
            std::memset(info, 0, sizeof(*info));
        }
        const F *acc[3] = {m_lf_acc[0].p, m_lf_acc[1].p, m_lf_acc[2].p};
        F *v[3] = {m_lf_v[0].p, m_lf_v[1].p, m_lf_v[2].p};
        F *kv[3] = {m_lf_kv[0].p, m_lf_kv[1].p, m_lf_kv[2].p};
        try {
            RK_CUDA_CHECK(cudaEventRecord(m_lf_ev[0], m_stream));
            // conserved quantities at the beginning of the step (292-347)
            if (m_lf_track) {
                launch_lf_integrals<F>(m_b.psorted.p, v, m_lf_acc[3].p, n, m_lf_scratch.p, m_stream);
                RK_CUDA_CHECK(cudaMemcpyAsync(m_hpin + 40, m_lf_scratch.p, 7 * sizeof(double), cudaMemcpyDeviceToHost,
                                              m_stream));
            }
            RK_CUDA_CHECK(cudaEventRecord(m_lf_ev[1], m_stream));
            // kick + drift (349-370), written into the pre-sort array of the rebuild; then sync() (3678-3743)
            RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[0], m_stream));
            reset_flags();
            launch_lf_kick_drift<F>(acc, v, m_b.psorted.p, half_dt, dt, kv, m_b.pin.p, n,
                                    reinterpret_cast<u64 *>(m_b.d_misc.p), m_stream);
            RK_CUDA_CHECK(cudaEventRecord(m_lf_ev[2], m_stream));
            rk_build_info bi;
            rebuild(false, &bi);
            RK_CUDA_CHECK(cudaEventRecord(m_lf_ev[3], m_stream));
            // accelerations in the new positions, then the second half kick through last_perm (372-383)
            rk_eval_info ei;
            leapfrog_eval(&ei);
            RK_CUDA_CHECK(cudaEventRecord(m_lf_ev[4], m_stream));
            launch_lf_kick_reindex<F>(acc, kv, m_b.last_perm, half_dt, v, n, m_stream);
            RK_CUDA_CHECK(cudaEventRecord(m_lf_ev[5], m_stream));
            RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
            if (info) {
                cudaEventElapsedTime(&info->ms_step, m_lf_ev[0], m_lf_ev[5]);
                cudaEventElapsedTime(&info->ms_integrals, m_lf_ev[0], m_lf_ev[1]);
                cudaEventElapsedTime(&info->ms_kick_drift, m_lf_ev[1], m_lf_ev[2]);
                cudaEventElapsedTime(&info->ms_rebuild, m_lf_ev[2], m_lf_ev[3]);
                cudaEventElapsedTime(&info->ms_traverse, m_lf_ev[3], m_lf_ev[4]);
                cudaEventElapsedTime(&info->ms_reindex, m_lf_ev[4], m_lf_ev[5]);
                info->interactions = ei.interactions;
                info->n_nodes = bi.n_nodes;
                if (m_lf_track) {
                    const double *h = reinterpret_cast<const double *>(m_hpin + 40);
                    for (int j = 0; j < 3; ++j) {
                        info->com[j] = h[j] / double(n);
                        info->com_v[j] = h[3 + j] / double(n);
                    }
                    info->energy = h[6];
                }
            }
        } catch (...) {
            m_lf_ready = false;
            throw;
        }
    }
    // what: 0 velocities, 1 accelerations of the last evaluation, 2 kicked velocities of the last step (previous order),
    // 3 potentials (track_integrals only); all in the tree's internal order.
    void leapfrog_get(int what, void *a, void *b, void *c, int where)
    {
        use();
        if (!m_lf_ready) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "rk_tree_leapfrog_get: call rk_tree_leapfrog_init first");
        }
        const size_t n = m_b.n;
        const F *src[3] = {nullptr, nullptr, nullptr};
        if (what == 0 || what == 1 || what == 2) {
            dbuf<F> *arr = what == 0 ? m_lf_v : (what == 1 ? m_lf_acc : m_lf_kv);
            for (int j = 0; j < 3; ++j) {
                src[j] = arr[j].p;
            }
        } else if (what == 3 && m_lf_track) {
            src[0] = m_lf_acc[3].p;
        } else {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "rk_tree_leapfrog_get: invalid selector");
        }
        void *dst[3] = {a, b, c};
        for (int j = 0; j < 3; ++j) {
            if (dst[j] && src[j]) {
                RK_CUDA_CHECK(cudaMemcpyAsync(dst[j], src[j], n * sizeof(F),
                                              where == RK_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                              m_stream));
            }
        }
        RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
    }

    // out[j][perm[i]] = in[j][i] on the device: results from Morton order to the original particle order
    // (tree.hpp:3320-3330) for callers that assembled the Morton-order arrays themselves (sharded evaluation).
    void to_original_order(int nres, const void *const in[4], void *const out[4])
    {
        use();
        if (nres < 1 || nres > 4) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "rk_tree_to_original_order: 1 to 4 arrays");
        }
        for (int j = 0; j < nres; ++j) {
            if (!in[j] || !out[j] || in[j] == out[j]) {
                throw api_error(RK_ERR_INVALID_ARGUMENT, "rk_tree_to_original_order: distinct non-null device arrays");
            }
            launch_scatter_perm<F>(static_cast<const F *>(in[j]), m_b.perm.p, static_cast<F *>(out[j]), m_b.n, m_stream);
        }
        RK_CUDA_CHECK(cudaGetLastError());
    }
    const char *last_kernel() const { return m_kernel_name; }
    // Copies of the output arrays (device-accessible: peer memory, mapped host memory) that the following evaluations
    // with DEVICE outputs write as well, result by result, from inside the traversal kernel.
    void set_output_mirrors(unsigned n, void *const *ptrs, unsigned multicast_mask)
    {
        m_mirror_multicast = multicast_mask;
        if (n > TRAV_MAX_MIRRORS || (n && !ptrs)) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "rk_tree_set_output_mirrors: at most " + std::to_string(TRAV_MAX_MIRRORS)
                                                         + " mirrors of 4 pointers each");
        }
        for (unsigned r = 0; r < n; ++r) {
            for (int j = 0; j < 4; ++j) {
                m_out_mirror[r][j] = static_cast<F *>(ptrs[4 * r + j]);
            }
        }
        m_n_mirror = n;
    }
    void set_option(const std::string &name, long long value)
    {
        if (name == "props_bottom_up") {
            m_b.props_bottom_up = value < 0 ? -1 : (value != 0);
        } else if (name == "zero_copy_out") {
            m_zero_copy_out = value != 0;
        } else {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "rk_tree_set_option: unknown option '" + name + "'");
        }
    }
    // Order-independent fingerprints of the device arrays (multi-GPU parity checks without moving the tree to the host).
    void digest(uint64_t out[8])
    {
        use();
        const size_t n = m_b.n, M = m_b.n_nodes, C = m_b.n_crit;
        u64 *d = reinterpret_cast<u64 *>(m_counters.p);
        RK_CUDA_CHECK(cudaMemsetAsync(d, 0, 8 * sizeof(u64), m_stream));
        if (n) {
            launch_digest(m_b.codes, n * sizeof(u64), d + 0, m_stream);
            launch_digest(m_b.perm.p, n * sizeof(u32), d + 1, m_stream);
            launch_digest(m_b.psorted.p, n * sizeof(vec4<F>), d + 2, m_stream);
            launch_digest(m_b.nodeB.p, M * sizeof(uint4), d + 3, m_stream);
            launch_digest(m_b.nodeA.p, M * sizeof(vec4<F>), d + 4, m_stream);
            launch_digest(m_b.crit_node.p, C * sizeof(u32), d + 5, m_stream);
            launch_digest(m_b.crit_begin.p, (C + 1) * sizeof(u32), d + 6, m_stream);
        }
        RK_CUDA_CHECK(cudaMemcpyAsync(m_hpin + 24, d, 8 * sizeof(u64), cudaMemcpyDeviceToHost, m_stream));
        RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
        std::memcpy(out, m_hpin + 24, 7 * sizeof(u64));
        out[7] = (static_cast<u64>(M) << 32) ^ C;
    }
    // First critical node whose first particle is >= pidx[j] (ncrit if none): re-snaps particle-index cuts to the
    // critical nodes of a rebuilt tree.
    void crit_lower_bound(const uint64_t *pidx, size_t k, uint64_t *out)
    {
        use();
        if (k > 16) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "rk_tree_crit_lower_bound: at most 16 values per call");
        }
        u64 *d = reinterpret_cast<u64 *>(m_counters.p);
        std::memcpy(m_hpin + 32, pidx, k * sizeof(u64));
        RK_CUDA_CHECK(cudaMemcpyAsync(d, m_hpin + 32, k * sizeof(u64), cudaMemcpyHostToDevice, m_stream));
        launch_lower_bound(m_b.crit_begin.p, m_b.n_crit, d, k, d + 16, m_stream);
        RK_CUDA_CHECK(cudaMemcpyAsync(m_hpin + 32, d + 16, k * sizeof(u64), cudaMemcpyDeviceToHost, m_stream));
        RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
        std::memcpy(out, m_hpin + 32, k * sizeof(u64));
    }

private:
    void leapfrog_eval(rk_eval_info *ei)
    {
        void *out[4] = {m_lf_acc[0].p, m_lf_acc[1].p, m_lf_acc[2].p, m_lf_acc[3].p};
        acc_pot(m_lf_track ? 2 : 0, false, m_lf_theta, m_lf_G, m_lf_eps, nullptr, 0, false, 0, 0, out, RK_DEVICE, ei);
    }
    void reserve_particles(size_t n)
    {
        m_b.pin.reserve(n, 1.05);
        m_b.psorted.reserve(n, 1.05);
        m_b.keys_a.reserve(n, 1.05);
        m_b.keys_b.reserve(n, 1.05);
        m_b.idx_a.reserve(n, 1.05);
        m_b.idx_b.reserve(n, 1.05);
        m_b.perm.reserve(n, 1.05);
        m_b.perm_tmp.reserve(n, 1.05);
        m_b.inv_perm.reserve(n, 1.05);
    }
    void reset_flags()
    {
        RK_CUDA_CHECK(cudaMemsetAsync(m_b.d_err.p, 0xff, 2 * sizeof(u64), m_stream));
        RK_CUDA_CHECK(cudaMemsetAsync(m_b.d_misc.p, 0, 8 * sizeof(u32), m_stream));
    }
    // Make x, y, z, m available on the device (NULL stays NULL).
    void upload4(const void *x, const void *y, const void *z, const void *m, size_t n, int where, const F *&dx,
                 const F *&dy, const F *&dz, const F *&dm)
    {
        const void *in[4] = {x, y, z, m};
        const F *d[4];
        for (int j = 0; j < 4; ++j) {
            if (!in[j] || !n) {
                d[j] = in[j] ? static_cast<const F *>(in[j]) : nullptr;
                continue;
            }
            if (where == RK_DEVICE) {
                d[j] = static_cast<const F *>(in[j]);
            } else {
                m_b.stage[j].reserve(n, 1.05);
                RK_CUDA_CHECK(cudaMemcpyAsync(m_b.stage[j].p, in[j], n * sizeof(F), cudaMemcpyHostToDevice, m_stream));
                d[j] = m_b.stage[j].p;
            }
        }
        dx = d[0];
        dy = d[1];
        dz = d[2];
        dm = d[3];
    }

    // Encode -> sort -> permute -> topology -> properties, shared by build() and update_positions().
    void rebuild(bool first, rk_build_info *info)
    {
        const size_t n = m_b.n;
        if (info) {
            std::memset(info, 0, sizeof(*info));
        }
        if (!n) {
            if (m_box_deduced) {
                m_box = F(0);
            }
            if (info) {
                info->box_size = m_box;
            }
            return;
        }
        // ---- box size, determine_box_size tree.hpp:1278-1319 ----
        if (m_box_deduced) {
            RK_CUDA_CHECK(cudaMemcpyAsync(m_hpin, m_b.d_misc.p, sizeof(u64), cudaMemcpyDeviceToHost, m_stream));
            RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
            F mx;
            if (sizeof(F) == 4) {
                const u32 bits = static_cast<u32>(m_hpin[0]);
                std::memcpy(&mx, &bits, 4);
            } else {
                std::memcpy(&mx, &m_hpin[0], 8);
            }
            if (!std::isfinite(mx)) {
                throw api_error(RK_ERR_INVALID_ARGUMENT, "While trying to automatically determine the domain size, a "
                                                         "non-finite coordinate with absolute value "
                                                             + std::to_string(std::abs(mx)) + " was encountered");
            }
            F b = mx * F(2);
            b = std::fma(b, F(1) / F(20), b);
            if (!std::isfinite(b)) {
                throw api_error(RK_ERR_INVALID_ARGUMENT,
                                "The automatic deduction of the domain size produced the non-finite value "
                                    + std::to_string(b));
            }
            m_box = b;
        }
        const F inv_box = F(1) / m_box;
        // ---- Morton encoding ----
        launch_encode<F>(m_b.pin.p, m_b.keys_a.p, n, inv_box, m_b.d_err.p, m_stream);
        RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[1], m_stream));
        // ---- sort (includes one sync for the varying-bits mask) ----
        u64 *codes;
        u32 *lperm;
        const int passes
            = radix_sort_pairs(m_b.keys_a.p, m_b.keys_b.p, m_b.idx_a.p, m_b.idx_b.p, n, m_sc, m_stream, &codes, &lperm);
        m_b.codes = codes;
        m_b.last_perm = lperm;
        RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[2], m_stream));
        // ---- permute ----
        if (m_late_m) {
            RK_CUDA_CHECK(cudaStreamWaitEvent(m_stream, m_copy_ev, 0));
        }
        launch_gather<F>(m_b.pin.p, lperm, m_b.psorted.p, n, m_stream, m_late_m);
        m_late_m = nullptr;
        if (first) {
            launch_perm_first(lperm, m_b.perm.p, nullptr, n, m_stream);
        } else {
            launch_perm_compose(m_b.perm.p, lperm, m_b.perm_tmp.p, nullptr, n, m_stream);
            std::swap(m_b.perm.p, m_b.perm_tmp.p);
            std::swap(m_b.perm.cap, m_b.perm_tmp.cap);
        }
        RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[3], m_stream));
        m_pending_inv_box = inv_box;
        m_pending_check_encode = true;
        finish_build(passes, info, [] {});
    }

    // Topology + node properties on sorted codes / particles (events 0..3 already recorded).
    template <typename Hook>
    void finish_build(int passes, rk_build_info *info, Hook &&before_props)
    {
        const size_t n = m_b.n;
        if (info) {
            std::memset(info, 0, sizeof(*info));
        }
        if (!n) {
            return;
        }
        const F inv_box = m_pending_inv_box;
        // ---- topology: count, size, emit ----
        topology_count<F>(m_b, m_max_leaf_n, m_ncrit, m_stream);
        RK_CUDA_CHECK(cudaMemcpyAsync(m_hpin, m_b.d_err.p, 2 * sizeof(u64), cudaMemcpyDeviceToHost, m_stream));
        RK_CUDA_CHECK(
            cudaMemcpyAsync(m_hpin + 2, m_b.rowtot.p, (NLEVELS + 1) * sizeof(u32), cudaMemcpyDeviceToHost, m_stream));
        RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
        if (m_pending_check_encode) {
            m_pending_check_encode = false;
            check_encode_error(m_hpin[0], inv_box);
        }
        const u32 *rt = reinterpret_cast<const u32 *>(m_hpin + 2);
        u64 M = 0;
        for (int l = 0; l < NLEVELS; ++l) {
            m_b.levels.base[l] = static_cast<u32>(M);
            M += rt[l];
        }
        if (M > 0xfffffff0ull / 8) {
            throw api_error(RK_ERR_OVERFLOW, "The size of the tree (" + std::to_string(M)
                                                 + ") is too large, and it results in an overflow condition");
        }
        m_b.levels.base[NLEVELS] = static_cast<u32>(M);
        m_b.levels.base[NLEVELS + 1] = static_cast<u32>(M);
        const u64 C = rt[NLEVELS];
        m_b.n_nodes = M;
        m_b.n_crit = C;
        m_b.nodeA.reserve(M, 1.1);
        m_b.nodeB.reserve(M, 1.1);
        m_b.node_dfs.reserve(M, 1.1);
        m_b.node_ndesc.reserve(M, 1.1);
        if (m_mac == RK_MAC_BH_GEOM) {
            m_b.node_delta.reserve(M, 1.1);
        }
        m_b.crit_node.reserve(C, 1.1);
        m_b.crit_begin.reserve(C + 1, 1.1);
        topology_emit<F>(m_b, m_stream);
        RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[4], m_stream));
        // ---- node properties ----
        before_props();
        node_properties<F>(m_b, m_mac, m_box, m_stream);
        RK_CUDA_CHECK(cudaEventRecord(m_ev.ev[5], m_stream));
        RK_CUDA_CHECK(cudaMemcpyAsync(m_hpin, m_b.d_err.p, 2 * sizeof(u64), cudaMemcpyDeviceToHost, m_stream));
        RK_CUDA_CHECK(cudaMemcpyAsync(m_hpin + 2, m_b.d_misc.p, 4 * sizeof(u32), cudaMemcpyDeviceToHost, m_stream));
        // first particles of the critical nodes at the 40/70/90 % marks: where acc_pot() cuts a host-output
        // evaluation into pipelined launches (read back here so that the evaluation needs no extra sync)
        u32 *hcut = reinterpret_cast<u32 *>(m_hpin + 4);
        for (int k = 0; k < NCHUNK - 1; ++k) {
            m_cut_crit[k] = C * chunk_mark(k + 1) / 10;
            RK_CUDA_CHECK(cudaMemcpyAsync(hcut + k, m_b.crit_begin.p + m_cut_crit[k], sizeof(u32),
                                          cudaMemcpyDeviceToHost, m_stream));
        }
        RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
        check_props_error(m_hpin[1]);
        m_max_group = reinterpret_cast<const u32 *>(m_hpin + 2)[2];
        for (int k = 0; k < NCHUNK - 1; ++k) {
            m_cut_begin[k] = hcut[k];
        }
        m_cuts_valid = true;
        m_costs_valid = false;
        ++m_epoch;
        m_have_inv = false;
        m_h_crit_begin.clear();
        if (info) {
            info->box_size = m_box;
            info->n_nodes = M;
            info->n_crit = C;
            info->max_group = m_max_group;
            info->sort_passes = static_cast<uint32_t>(passes);
            cudaEventElapsedTime(&info->ms_total, m_ev.ev[0], m_ev.ev[5]);
            cudaEventElapsedTime(&info->ms_encode, m_ev.ev[0], m_ev.ev[1]);
            cudaEventElapsedTime(&info->ms_sort, m_ev.ev[1], m_ev.ev[2]);
            cudaEventElapsedTime(&info->ms_permute, m_ev.ev[2], m_ev.ev[3]);
            cudaEventElapsedTime(&info->ms_topology, m_ev.ev[3], m_ev.ev[4]);
            cudaEventElapsedTime(&info->ms_props, m_ev.ev[4], m_ev.ev[5]);
        }
    }

    // Messages of disc_single_coord, tree.hpp:398-426.
    void check_encode_error(u64 key, F inv_box)
    {
        if (key == ~0ull) {
            return;
        }
        const u32 cls = static_cast<u32>(key & 0xf), dim = static_cast<u32>((key >> 4) & 0xf);
        const size_t idx = static_cast<size_t>(key >> 8);
        vec4<F> v;
        RK_CUDA_CHECK(cudaMemcpy(&v, m_b.pin.p + idx, sizeof(v), cudaMemcpyDeviceToHost));
        const F x = dim == 0 ? v.x : (dim == 1 ? v.y : v.z);
        F tmp = std::fma(x, inv_box, F(1) / F(2));
        tmp *= F(u64(1) << CBITS);
        const std::string head = "The discretisation of the input coordinate " + std::to_string(x) + " in a box of size "
                                 + std::to_string(F(1) / inv_box);
        if (cls == 1) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, "While trying to discretise the input coordinate "
                                                         + std::to_string(x) + " in a box of size "
                                                         + std::to_string(F(1) / inv_box) + ", the non-finite value "
                                                         + std::to_string(tmp) + " was generated");
        }
        if (cls == 2) {
            throw api_error(RK_ERR_INVALID_ARGUMENT, head + " produced the floating-point value " + std::to_string(tmp)
                                                         + ", which is outside the allowed bounds");
        }
        throw api_error(RK_ERR_INVALID_ARGUMENT, head + " produced the integral value "
                                                     + std::to_string(static_cast<u64>(tmp))
                                                     + ", which is outside the allowed bounds");
    }
    // Messages of compute_node_properties, tree.hpp:1195-1235.
    void check_props_error(u64 key)
    {
        if (key == ~0ull) {
            return;
        }
        const u32 cls = static_cast<u32>(key & 0xff);
        if (cls == 5) {
            throw api_error(RK_ERR_INVALID_ARGUMENT,
                            "The computation of the centre of mass of a node produced a non-finite value");
        }
        if (cls == 6) {
            throw api_error(RK_ERR_INVALID_ARGUMENT,
                            "The computation of the total mass in a node produced the non-finite value inf");
        }
        throw api_error(RK_ERR_INVALID_ARGUMENT, "The computation of the distance between the centre of mass "
                                                 "and the geometric centre of a node produced the non-finite value inf");
    }
    void ensure_copy_stream()
    {
        if (!m_copy_stream) {
            RK_CUDA_CHECK(cudaStreamCreateWithFlags(&m_copy_stream, cudaStreamNonBlocking));
            for (auto &s : m_aux_stream) {
                RK_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
            }
            for (auto &e : m_chunk_ev) {
                RK_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            }
            RK_CUDA_CHECK(cudaEventCreateWithFlags(&m_copy_ev, cudaEventDisableTiming));
        }
    }
    void host_crit_begin()
    {
        if (m_h_crit_begin.size() == m_b.n_crit + 1) {
            return;
        }
        m_h_crit_begin.resize(m_b.n_crit + 1);
        RK_CUDA_CHECK(cudaMemcpyAsync(m_h_crit_begin.data(), m_b.crit_begin.p, (m_b.n_crit + 1) * sizeof(u32),
                                      cudaMemcpyDeviceToHost, m_stream));
        RK_CUDA_CHECK(cudaStreamSynchronize(m_stream));
    }

    int m_mac, m_device, m_sm_count = 148;
    cudaStream_t m_stream = nullptr, m_own_stream = nullptr, m_copy_stream = nullptr;
    static constexpr int NCHUNK = 4; // pipelined host-output evaluation: launches of 40/30/20/10 % of the groups
    static constexpr size_t chunk_mark(int k) { return k == 1 ? 4 : k == 2 ? 7 : 9; } // tenths
    cudaStream_t m_aux_stream[NCHUNK - 1] = {};
    cudaEvent_t m_chunk_ev[NCHUNK] = {}, m_copy_ev = nullptr;
    const F *m_late_m = nullptr;
    size_t m_cut_crit[NCHUNK - 1] = {}, m_cut_begin[NCHUNK - 1] = {};
    bool m_cuts_valid = false;
    timer_events m_ev;
    build_arrays<F> m_b;
    sort_scratch m_sc;
    F m_box = F(0);
    bool m_box_deduced = false;
    size_t m_max_leaf_n = 16, m_ncrit = 128, m_max_group = 0;
    dbuf<F> m_out[4], m_out_ord[4];
    // multi-device evaluation: mirrors of this tree on the other devices; on a mirror, what it mirrors
    std::vector<std::unique_ptr<tree>> m_mirror;
    const tree *m_mirror_of = nullptr;
    unsigned long long m_epoch = 1, m_mirror_epoch = 0;
    size_t m_range_groups = 0;
    dbuf<u64> m_group_cost, m_counters;
    dbuf<u32> m_work, m_steal;
    F *m_out_mirror[TRAV_MAX_MIRRORS][4] = {};
    unsigned m_n_mirror = 0, m_mirror_multicast = 0; // rk_tree_set_output_mirrors
    int m_zero_copy_out = 1; // rk_tree_set_option("zero_copy_out"): final results straight into pinned host outputs
    dbuf<u64> m_ids_sorted; // partition_shard scratch
    bool m_costs_valid = false, m_have_inv = false, m_pending_check_encode = false;
    F m_pending_inv_box = F(0);
    u64 *m_hpin = nullptr; // pinned scratch for small read-backs
    std::vector<u32> m_h_crit_begin;
    char m_kernel_name[96] = "";
    // leapfrog state: velocities / kicked velocities / accelerations (+ potentials) in the tree's internal order
    dbuf<F> m_lf_v[3], m_lf_kv[3], m_lf_acc[4];
    dbuf<double> m_lf_scratch;
    cudaEvent_t m_lf_ev[6] = {};
    double m_lf_theta = 0.75, m_lf_G = 1, m_lf_eps = 0;
    bool m_lf_track = false, m_lf_ready = false;
};

} // namespace rk

struct rk_tree {
    int fp = 32, mac = 0;
    rk::tree<float> *t32 = nullptr;
    rk::tree<double> *t64 = nullptr;
    std::string err;
    std::mutex mu;
};

namespace
{
thread_local std::string g_create_error;

template <typename Fn>
int guarded(rk_tree *t, Fn &&fn)
{
    if (!t) {
        return RK_ERR_INVALID_ARGUMENT;
    }
    std::lock_guard<std::mutex> lock(t->mu);
    try {
        fn();
        return RK_OK;
    } catch (const rk::api_error &e) {
        t->err = e.what();
        return e.status;
    } catch (const rk::cuda_error &e) {
        t->err = e.what();
        return e.status;
    } catch (const std::bad_alloc &) {
        t->err = "bad_alloc";
        return RK_ERR_BAD_ALLOC;
    } catch (const std::exception &e) {
        t->err = e.what();
        return RK_ERR_RUNTIME;
    }
}
} // namespace

#define RK_WITH(t, expr)                                                                                               \
    do {                                                                                                               \
        if ((t)->fp == 32) {                                                                                           \
            auto &T = *(t)->t32;                                                                                       \
            expr;                                                                                                      \
        } else {                                                                                                       \
            auto &T = *(t)->t64;                                                                                       \
            expr;                                                                                                      \
        }                                                                                                              \
    } while (0)

extern "C" {

unsigned rk_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n < 0 ? 0u : static_cast<unsigned>(n);
}

unsigned rk_min_size(void)
{
    return 1000u; // same threshold as the reference, rakau_cuda.cu:26-29
}

rk_tree *rk_tree_create(int fp_bits, int mac, int device)
{
    g_create_error.clear();
    if ((fp_bits != 32 && fp_bits != 64) || (mac != RK_MAC_BH && mac != RK_MAC_BH_GEOM)) {
        g_create_error = "rk_tree_create: fp_bits must be 32 or 64 and mac RK_MAC_BH or RK_MAC_BH_GEOM";
        return nullptr;
    }
    if (rk_device_count() == 0) {
        g_create_error = "rk_tree_create: no CUDA device is available (librakau_b200 has no CPU fallback)";
        return nullptr;
    }
    rk_tree *t = nullptr;
    try {
        t = new rk_tree;
        t->fp = fp_bits;
        t->mac = mac;
        if (fp_bits == 32) {
            t->t32 = new rk::tree<float>(mac, device);
        } else {
            t->t64 = new rk::tree<double>(mac, device);
        }
        return t;
    } catch (const std::exception &e) {
        g_create_error = e.what();
        delete t;
        return nullptr;
    }
}

void rk_tree_destroy(rk_tree *t)
{
    if (t) {
        delete t->t32;
        delete t->t64;
        delete t;
    }
}

const char *rk_last_error(const rk_tree *t)
{
    return t ? t->err.c_str() : "null tree handle";
}
const char *rk_create_error(void)
{
    return g_create_error.c_str();
}

int rk_tree_set_stream(rk_tree *t, void *s)
{
    return guarded(t, [&]() { RK_WITH(t, T.set_stream(s)); });
}
int rk_tree_synchronize(rk_tree *t)
{
    return guarded(t, [&]() { RK_WITH(t, T.synchronize()); });
}

int rk_tree_build(rk_tree *t, const void *x, const void *y, const void *z, const void *m, size_t n, int where,
                  double box_size, int deduce_box, size_t max_leaf_n, size_t ncrit, rk_build_info *info)
{
    return guarded(
        t, [&]() { RK_WITH(t, T.build(x, y, z, m, n, where, box_size, deduce_box != 0, max_leaf_n, ncrit, info)); });
}
int rk_tree_update_positions(rk_tree *t, const void *x, const void *y, const void *z, const void *m, int where,
                             rk_build_info *info)
{
    return guarded(t, [&]() { RK_WITH(t, T.update_positions(x, y, z, m, where, info)); });
}
int rk_tree_update_masses(rk_tree *t, const void *m, int where)
{
    return guarded(t, [&]() { RK_WITH(t, T.update_masses(m, where)); });
}
int rk_tree_sort_shard(rk_tree *t, const void *x, const void *y, const void *z, const void *m, const uint64_t *codes,
                       size_t n, double box_size)
{
    return guarded(t, [&]() { RK_WITH(t, T.sort_shard(x, y, z, m, codes, n, box_size)); });
}
int rk_tree_encode_shard(rk_tree *t, const void *x, const void *y, const void *z, const void *m, size_t n, double box_size)
{
    return guarded(t, [&]() { RK_WITH(t, T.encode_shard(x, y, z, m, n, box_size)); });
}
int rk_tree_partition_shard(rk_tree *t, const uint64_t *splitters, unsigned nsplit, uint64_t *counts)
{
    return guarded(t, [&]() { RK_WITH(t, T.partition_shard(splitters, nsplit, counts)); });
}
int rk_tree_get_codes_device(rk_tree *t, uint64_t *out)
{
    return guarded(t, [&]() { RK_WITH(t, T.get_codes_device(out)); });
}
int rk_tree_build_presorted(rk_tree *t, const void *x, const void *y, const void *z, const void *m,
                            const uint64_t *codes, const uint32_t *perm, size_t n, double box_size, size_t max_leaf_n,
                            size_t ncrit, void *parts_ready_event, rk_build_info *info)
{
    return guarded(t, [&]() {
        RK_WITH(t, T.build_presorted(x, y, z, m, codes, perm, n, box_size, max_leaf_n, ncrit, parts_ready_event, info));
    });
}
double rk_deduce_box(int fp_bits, double absmax)
{
    // determine_box_size, tree.hpp:1309-1312: 2 * max|coord| plus a 5 % slack, in the tree's precision
    if (fp_bits == 32) {
        float b = static_cast<float>(absmax) * 2.f;
        b = std::fma(b, 1.f / 20.f, b);
        return b;
    }
    double b = absmax * 2.;
    return std::fma(b, 1. / 20., b);
}
int rk_tree_clone(rk_tree *dst, const rk_tree *src)
{
    if (!src || !dst || src->fp != dst->fp || src->mac != dst->mac) {
        return RK_ERR_INVALID_ARGUMENT;
    }
    return guarded(dst, [&]() {
        if (dst->fp == 32) {
            dst->t32->clone_from(*src->t32);
        } else {
            dst->t64->clone_from(*src->t64);
        }
    });
}
int rk_tree_clear(rk_tree *t)
{
    return guarded(t, [&]() { RK_WITH(t, T.clear()); });
}

size_t rk_tree_nparts(const rk_tree *t)
{
    return t->fp == 32 ? t->t32->nparts() : t->t64->nparts();
}
size_t rk_tree_nnodes(const rk_tree *t)
{
    return t->fp == 32 ? t->t32->nnodes() : t->t64->nnodes();
}
size_t rk_tree_ncrit(const rk_tree *t)
{
    return t->fp == 32 ? t->t32->ncrit_nodes() : t->t64->ncrit_nodes();
}
double rk_tree_box_size(const rk_tree *t)
{
    return t->fp == 32 ? t->t32->box_size() : t->t64->box_size();
}
int rk_tree_get_parts(rk_tree *t, void *x, void *y, void *z, void *m)
{
    return guarded(t, [&]() { RK_WITH(t, T.get_parts(x, y, z, m)); });
}
int rk_tree_get_parts_device(rk_tree *t, void *x, void *y, void *z, void *m)
{
    return guarded(t, [&]() { RK_WITH(t, T.get_parts_device(x, y, z, m)); });
}
int rk_tree_get_perm_device(rk_tree *t, int which, uint32_t *out)
{
    return guarded(t, [&]() { RK_WITH(t, T.get_perm_device(which, out)); });
}
int rk_tree_get_codes(rk_tree *t, uint64_t *codes)
{
    return guarded(t, [&]() { RK_WITH(t, T.get_codes(codes)); });
}
int rk_tree_get_perm(rk_tree *t, int which, uint64_t *out)
{
    return guarded(t, [&]() { RK_WITH(t, T.get_perm(which, out)); });
}
int rk_tree_get_nodes(rk_tree *t, void *nodes)
{
    return guarded(t, [&]() { RK_WITH(t, T.get_nodes(nodes)); });
}
int rk_tree_get_crit(rk_tree *t, rk_cnode *crit)
{
    return guarded(t, [&]() { RK_WITH(t, T.get_crit(crit)); });
}
int rk_tree_crit_begin_at(rk_tree *t, const size_t *idx, size_t k, uint64_t *out)
{
    return guarded(t, [&]() { RK_WITH(t, T.crit_begin_at(idx, k, out)); });
}
int rk_tree_get_group_costs(rk_tree *t, uint64_t *costs)
{
    return guarded(t, [&]() { RK_WITH(t, T.get_group_costs(costs)); });
}

int rk_tree_acc_pot(rk_tree *t, int Q, int ordered, double theta, double G, double eps, const double *split,
                    size_t nsplit, void *const out[4], int where, rk_eval_info *info)
{
    return guarded(t, [&]() {
        RK_WITH(t, T.acc_pot(Q, ordered != 0, theta, G, eps, split, nsplit, false, 0, 0, out, where, info));
    });
}
int rk_tree_acc_pot_range(rk_tree *t, int Q, int ordered, double theta, double G, double eps, size_t crit_begin,
                          size_t crit_end, void *const out[4], int where, rk_eval_info *info)
{
    return guarded(t, [&]() {
        RK_WITH(t, T.acc_pot(Q, ordered != 0, theta, G, eps, nullptr, 0, true, crit_begin, crit_end, out, where, info));
    });
}
int rk_tree_exact(rk_tree *t, size_t idx, int ordered, double G, double eps, double out4[4])
{
    return guarded(t, [&]() { RK_WITH(t, T.exact(idx, ordered != 0, G, eps, out4)); });
}

const void *rk_tree_group_costs_device(rk_tree *t)
{
    if (!t) {
        return nullptr;
    }
    return t->fp == 32 ? t->t32->group_costs_device() : t->t64->group_costs_device();
}

int rk_tree_leapfrog_init(rk_tree *t, const void *vx, const void *vy, const void *vz, int where, double theta, double G,
                          double eps, int track_integrals)
{
    return guarded(t, [&]() { RK_WITH(t, T.leapfrog_init(vx, vy, vz, where, theta, G, eps, track_integrals != 0)); });
}
int rk_tree_leapfrog_step(rk_tree *t, double dt, rk_leapfrog_info *info)
{
    return guarded(t, [&]() { RK_WITH(t, T.leapfrog_step(dt, info)); });
}
int rk_tree_leapfrog_get(rk_tree *t, int what, void *a, void *b, void *c, int where)
{
    return guarded(t, [&]() { RK_WITH(t, T.leapfrog_get(what, a, b, c, where)); });
}
int rk_tree_set_output_mirrors(rk_tree *t, unsigned n, void *const *ptrs, unsigned multicast_mask)
{
    return guarded(t, [&]() { RK_WITH(t, T.set_output_mirrors(n, ptrs, multicast_mask)); });
}
int rk_tree_set_option(rk_tree *t, const char *name, long long value)
{
    return guarded(t, [&]() { RK_WITH(t, T.set_option(name ? name : "", value)); });
}
int rk_tree_to_original_order(rk_tree *t, int nres, const void *const in[4], void *const out[4])
{
    return guarded(t, [&]() { RK_WITH(t, T.to_original_order(nres, in, out)); });
}
int rk_tree_digest(rk_tree *t, uint64_t out[8])
{
    return guarded(t, [&]() { RK_WITH(t, T.digest(out)); });
}
int rk_tree_crit_lower_bound(rk_tree *t, const uint64_t *particle_idx, size_t k, uint64_t *out)
{
    return guarded(t, [&]() { RK_WITH(t, T.crit_lower_bound(particle_idx, k, out)); });
}
const char *rk_tree_last_kernel(const rk_tree *t)
{
    if (!t) {
        return "";
    }
    return t->fp == 32 ? t->t32->last_kernel() : t->t64->last_kernel();
}

int rk_device_copy_async(void *dst, const void *src, size_t bytes, void *stream)
{
    if (!bytes) {
        return RK_OK;
    }
    if (!dst || !src) {
        return RK_ERR_INVALID_ARGUMENT;
    }
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)) == cudaSuccess
               ? RK_OK
               : RK_ERR_RUNTIME;
}
int rk_device_bcast_copy(void *const *dst, unsigned ndst, const void *src, size_t bytes, void *stream, int multicast)
{
    if (!bytes || !ndst) {
        return RK_OK;
    }
    if (multicast && (ndst != 1 || bytes % 8)) {
        return RK_ERR_INVALID_ARGUMENT;
    }
    if (!dst || !src || ndst > 8 || (reinterpret_cast<uintptr_t>(src) & 7u)) {
        return RK_ERR_INVALID_ARGUMENT;
    }
    for (unsigned k = 0; k < ndst; ++k) {
        if (!dst[k] || (reinterpret_cast<uintptr_t>(dst[k]) & 7u)) {
            return RK_ERR_INVALID_ARGUMENT;
        }
    }
    try {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
            return RK_ERR_RUNTIME;
        }
        rk::launch_bcast_copy(dst, static_cast<int>(ndst), src, bytes, sms, static_cast<cudaStream_t>(stream), multicast != 0);
        return cudaGetLastError() == cudaSuccess ? RK_OK : RK_ERR_RUNTIME;
    } catch (...) {
        return RK_ERR_RUNTIME;
    }
}
unsigned long long rk_kernel_launch_count(void)
{
    return __atomic_load_n(&rk::g_kernel_launches, __ATOMIC_RELAXED);
}

int rk_measure_fp32_peak(int device, double *tflops, double *ms)
{
    try {
        RK_CUDA_CHECK(cudaSetDevice(device));
        float t = 0;
        const double flops = rk::ffma_microbench(&t);
        if (tflops) {
            *tflops = flops / (double(t) * 1e-3) / 1e12;
        }
        if (ms) {
            *ms = t;
        }
        return RK_OK;
    } catch (...) {
        return RK_ERR_RUNTIME;
    }
}

int rk_measure_fp64_peak(int device, double *tflops, double *ms)
{
    try {
        RK_CUDA_CHECK(cudaSetDevice(device));
        float t = 0;
        const double flops = rk::dfma_microbench(&t);
        if (tflops) {
            *tflops = flops / (double(t) * 1e-3) / 1e12;
        }
        if (ms) {
            *ms = t;
        }
        return RK_OK;
    } catch (...) {
        return RK_ERR_RUNTIME;
    }
}

int rk_traverse_external_tree(int fp_bits, int mac, int Q, void *const out[4], const uint64_t *split_indices,
                              size_t nsplit, const void *tree, size_t tree_size, const void *const parts[4],
                              const uint64_t *codes, size_t nparts, double mac_value, double G, double eps2,
                              int offset_output, size_t ncrit, rk_eval_info *info, char *errbuf, size_t errbuf_len)
{
    (void)codes; // the grouped traversal identifies ancestors by particle ranges, not by Morton codes
    rk_tree *t = rk_tree_create(fp_bits, mac, 0);
    auto fail = [&](int rc, const char *msg) {
        if (errbuf && errbuf_len) {
            std::strncpy(errbuf, msg, errbuf_len - 1);
            errbuf[errbuf_len - 1] = 0;
        }
        return rc;
    };
    if (!t) {
        return fail(RK_ERR_RUNTIME, rk_create_error());
    }
    const int rc = guarded(t, [&]() {
        if (fp_bits == 32) {
            t->t32->traverse_external(Q, out, split_indices, nsplit, static_cast<const rk::host_node_t<float> *>(tree),
                                      tree_size, parts, nparts, static_cast<float>(mac_value), static_cast<float>(G),
                                      static_cast<float>(eps2), offset_output != 0, ncrit, info);
        } else {
            t->t64->traverse_external(Q, out, split_indices, nsplit, static_cast<const rk::host_node_t<double> *>(tree),
                                      tree_size, parts, nparts, mac_value, G, eps2, offset_output != 0, ncrit, info);
        }
    });
    if (rc != RK_OK) {
        fail(rc, t->err.c_str());
    }
    rk_tree_destroy(t);
    return rc;
}

} // extern "C"
