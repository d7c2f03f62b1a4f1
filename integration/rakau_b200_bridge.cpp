// rakau_b200_bridge.cpp — the seam-level drop-in of INTEGRATION.md §2 as a real translation unit.
//
// It replaces the reference's src/rakau_cuda.cu: the reference header include/rakau/tree.hpp stays untouched, is
// compiled with RAKAU_WITH_CUDA, and its accelerator branch (tree.hpp:3131-3257) calls the three functions declared in
// include/rakau/detail/cuda_fwd.hpp:23-30, which are defined here on top of the C ABI of librakau_b200.so.
// Built and exercised by the test suite: oracle/Makefile compiles it together with the unmodified reference header
// into oracle/_ref/libref_bridge.so, tests/test_gpu_bridge.py runs the reference's accs_u(..., split = {0, 1}).
#include <array>
#include <cstddef>
#include <cstdint>
#include <new>
#include <stdexcept>
#include <type_traits>
#include <vector>

#include <rakau/detail/cuda_fwd.hpp>
#include <rakau/detail/tree_fwd.hpp>

#include <rakau_b200.h>

namespace rakau
{
inline namespace detail
{

unsigned cuda_min_size() { return rk_min_size(); }         // src/rakau_cuda.cu:26-29
unsigned cuda_device_count() { return rk_device_count(); } // src/rakau_cuda.cu:32-44

template <unsigned Q, std::size_t NDim, typename F, typename UInt, mac MAC>
void cuda_acc_pot_impl(const std::array<F *, tree_nvecs_res<Q, NDim>> &out,
                       const std::vector<tree_size_t<F>> &split_indices, const tree_node_t<NDim, F, UInt, MAC> *tree,
                       tree_size_t<F> tree_size, const std::array<const F *, NDim + 1u> &parts, const UInt *codes,
                       tree_size_t<F> nparts, F mac_value, F G, F eps2, bool offset_output)
{
    static_assert(NDim == 3 && sizeof(UInt) == 8, "rakau_b200 implements the 3-D, 64-bit path");
    // tree_node_t<3, F, u64, bh> is {u64 begin, end, n_children, code, level; F props[4]; F dim2}; rk_node_f32/f64 has
    // one more trailing F (delta): repack once, O(tree_size).
    using rk_node = std::conditional_t<std::is_same_v<F, float>, rk_node_f32, rk_node_f64>;
    std::vector<rk_node> nodes(tree_size);
    for (tree_size_t<F> i = 0; i < tree_size; ++i) {
        rk_node &n = nodes[i];
        n.begin = tree[i].begin;
        n.end = tree[i].end;
        n.n_children = tree[i].n_children;
        n.code = tree[i].code;
        n.level = tree[i].level;
        for (int j = 0; j < 4; ++j) {
            n.props[j] = tree[i].props[j];
        }
        if constexpr (MAC == mac::bh) {
            n.dim = tree[i].dim2;
            n.delta = 0;
        } else {
            n.dim = tree[i].dim;
            n.delta = tree[i].delta;
        }
    }
    void *o[4] = {nullptr, nullptr, nullptr, nullptr};
    for (std::size_t j = 0; j < out.size(); ++j) {
        o[j] = out[j];
    }
    const void *p[4] = {parts[0], parts[1], parts[2], parts[3]};
    const std::vector<std::uint64_t> si(split_indices.begin(), split_indices.end());
    char err[512] = "";
    // ncrit = 0: the reference's default (128). A maintainer passes m_ncrit through the call at tree.hpp:3207/3220.
    const int rc = rk_traverse_external_tree(sizeof(F) * 8, MAC == mac::bh ? RK_MAC_BH : RK_MAC_BH_GEOM, int(Q), o,
                                             si.data(), si.size(), nodes.data(), tree_size, p,
                                             reinterpret_cast<const std::uint64_t *>(codes), nparts, mac_value, G, eps2,
                                             offset_output ? 1 : 0, 0, nullptr, err, sizeof(err));
    if (rc == RK_ERR_BAD_ALLOC) {
        throw std::bad_alloc{}; // src/rakau_cuda.cu:47-55
    }
    if (rc != RK_OK) {
        throw std::runtime_error(err); // src/rakau_cuda.cu:83-127
    }
}

// explicit instantiations as src/rakau_cuda.cu:536-568, restricted to NDim = 3, UInt = std::uint64_t
#define RK_BRIDGE_INST(Q, F, MAC)                                                                                       \
    template void cuda_acc_pot_impl<Q, 3, F, std::uint64_t, MAC>(                                                      \
        const std::array<F *, tree_nvecs_res<Q, 3>> &, const std::vector<tree_size_t<F>> &,                            \
        const tree_node_t<3, F, std::uint64_t, MAC> *, tree_size_t<F>, const std::array<const F *, 4> &,               \
        const std::uint64_t *, tree_size_t<F>, F, F, F, bool);
RK_BRIDGE_INST(0, float, mac::bh)
RK_BRIDGE_INST(1, float, mac::bh)
RK_BRIDGE_INST(2, float, mac::bh)
RK_BRIDGE_INST(0, float, mac::bh_geom)
RK_BRIDGE_INST(1, float, mac::bh_geom)
RK_BRIDGE_INST(2, float, mac::bh_geom)
RK_BRIDGE_INST(0, double, mac::bh)
RK_BRIDGE_INST(1, double, mac::bh)
RK_BRIDGE_INST(2, double, mac::bh)
RK_BRIDGE_INST(0, double, mac::bh_geom)
RK_BRIDGE_INST(1, double, mac::bh_geom)
RK_BRIDGE_INST(2, double, mac::bh_geom)
#undef RK_BRIDGE_INST

} // namespace detail
} // namespace rakau
