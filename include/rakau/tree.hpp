// rakau/tree.hpp — the reference's public API (`rakau::tree<NDim, F, UInt, MAC>`, `rakau::octree<F, MAC>`,
// the keyword arguments of namespace rakau::kwargs) on top of the B200 library librakau_b200.so.
//
// This header is the host side of the drop-in: same names, argument meaning and error behaviour as
// /root/reference/include/rakau/tree.hpp for the Barnes-Hut path (constructors 1573-1733, accs/pots
// 3405-3497, exact_* 3571-3616, iterators/permutations/nodes 3637-3673, update_* 3767-3817, getters
// 3818-3837, clear 1882-1904, pprint 1906-1958). Everything heavy — box deduction, Morton encoding, sort,
// octree build, node properties, traversal — runs on the GPU through the C ABI of include/rakau_b200.h; the
// host keeps lazily fetched mirrors only because the reference API hands out raw host pointers.
//
// Deliberate deviations: only NDim = 3, F in {float, double} and a 64-bit UInt are instantiable;
// the collision-graph members (compute_cgraph_*) are not part of this path; there is no CPU share, so every
// entry of `split` maps to the tree's GPU (the vector is validated exactly as tree.hpp:2857-2868, 3134-3141).
#ifndef RAKAU_B200_TREE_HPP
#define RAKAU_B200_TREE_HPP

#include <algorithm>
#include <array>
#include <bitset>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <iterator>
#include <limits>
#include <mutex>
#include <new>
#include <ostream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include "../rakau_b200.h"
#include "detail/aligned_allocator.hpp"
#include "detail/kwargs.hpp"
#include "detail/simple_timer.hpp"

namespace rakau
{

// Multipole acceptance criteria (reference detail/tree_fwd.hpp:46).
enum class mac { bh, bh_geom };

inline namespace detail
{

template <typename F>
using tree_size_t = std::size_t;

// Node layouts of the reference (detail/tree_fwd.hpp:76-116): same members, same order.
template <std::size_t NDim, typename F, typename UInt>
struct base_tree_node_t {
    tree_size_t<F> begin, end, n_children;
    UInt code, level;
    friend bool operator==(const base_tree_node_t &a, const base_tree_node_t &b)
    {
        return a.begin == b.begin && a.end == b.end && a.n_children == b.n_children && a.code == b.code
               && a.level == b.level;
    }
};

template <std::size_t NDim, typename F, typename UInt, mac MAC>
struct tree_node_t;

template <std::size_t NDim, typename F, typename UInt>
struct tree_node_t<NDim, F, UInt, mac::bh> : base_tree_node_t<NDim, F, UInt> {
    F props[NDim + 1u], dim2;
    friend bool operator==(const tree_node_t &a, const tree_node_t &b)
    {
        using base = base_tree_node_t<NDim, F, UInt>;
        return std::equal(std::begin(a.props), std::end(a.props), std::begin(b.props)) && a.dim2 == b.dim2
               && static_cast<const base &>(a) == static_cast<const base &>(b);
    }
    friend bool operator!=(const tree_node_t &a, const tree_node_t &b) { return !(a == b); }
};

template <std::size_t NDim, typename F, typename UInt>
struct tree_node_t<NDim, F, UInt, mac::bh_geom> : base_tree_node_t<NDim, F, UInt> {
    F props[NDim + 1u], dim, delta;
    friend bool operator==(const tree_node_t &a, const tree_node_t &b)
    {
        using base = base_tree_node_t<NDim, F, UInt>;
        return std::equal(std::begin(a.props), std::end(a.props), std::begin(b.props)) && a.dim == b.dim
               && a.delta == b.delta && static_cast<const base &>(a) == static_cast<const base &>(b);
    }
    friend bool operator!=(const tree_node_t &a, const tree_node_t &b) { return !(a == b); }
};

// Critical node (detail/tree_fwd.hpp:119-125).
template <typename F, typename UInt>
struct tree_cnode_t {
    UInt code;
    tree_size_t<F> begin, end;
};

template <unsigned Q, std::size_t NDim>
inline constexpr std::size_t tree_nvecs_res = (Q == 0u ? NDim : (Q == 1u ? std::size_t(1) : NDim + 1u));

template <typename UInt, std::size_t NDim>
inline constexpr UInt cbits_v
    = static_cast<UInt>(std::numeric_limits<UInt>::digits / NDim - !(std::numeric_limits<UInt>::digits % NDim));

// Level of a nodal code (position of the leading 1 divided by NDim), detail/tree_fwd.hpp:208-226.
template <std::size_t NDim, typename UInt>
inline UInt tree_level(UInt n)
{
    unsigned hb = 0;
    for (UInt v = n; v >>= 1;) {
        ++hb;
    }
    return static_cast<UInt>(hb / NDim);
}

inline constexpr unsigned default_max_leaf_n = 16; // tree.hpp:584
inline constexpr unsigned default_ncrit = 128;     // tree.hpp:589-595 (the non-AVX-512 value)

// Scalar FMA as the reference's fma_wrap with FP_FAST_FMA defined (tree.hpp:181-207).
inline float fma_wrap(float x, float y, float z) { return std::fma(x, y, z); }
inline double fma_wrap(double x, double y, double z) { return std::fma(x, y, z); }

// 3-D Morton decode of a 63-bit code: x from bits 0,3,..., y from 1,4,..., z from 2,5,...
inline std::uint64_t morton_compact3(std::uint64_t v)
{
    v &= 0x1249249249249249ull;
    v = (v ^ (v >> 2)) & 0x10c30c30c30c30c3ull;
    v = (v ^ (v >> 4)) & 0x100f00f00f00f00full;
    v = (v ^ (v >> 8)) & 0x1f0000ff0000ffull;
    v = (v ^ (v >> 16)) & 0x1f00000000ffffull;
    v = (v ^ (v >> 32)) & 0x1fffffull;
    return v;
}
inline std::uint64_t morton_spread3(std::uint64_t v)
{
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

// morton_encoder / morton_decoder functors for the one supported instantiation (tree.hpp:216-372).
template <std::size_t NDim, typename UInt>
struct morton_encoder {
    static_assert(NDim == 3 && std::numeric_limits<UInt>::digits == 64, "only the 3-D, 64-bit encoder is provided");
    template <typename It>
    UInt operator()(It it) const
    {
        return static_cast<UInt>(morton_spread3(*it) | (morton_spread3(*(it + 1)) << 1)
                                 | (morton_spread3(*(it + 2)) << 2));
    }
};
template <std::size_t NDim, typename UInt>
struct morton_decoder {
    static_assert(NDim == 3 && std::numeric_limits<UInt>::digits == 64, "only the 3-D, 64-bit decoder is provided");
    template <typename It>
    void operator()(It it, UInt code) const
    {
        *it = static_cast<UInt>(morton_compact3(code));
        *(it + 1) = static_cast<UInt>(morton_compact3(code >> 1));
        *(it + 2) = static_cast<UInt>(morton_compact3(code >> 2));
    }
};

// get_node_dim / get_node_centre, tree.hpp:443-482 (host-side helpers used by the reference's tests).
template <typename UInt, typename F>
inline F get_node_dim(UInt node_level, F box_size)
{
    return box_size / static_cast<F>(UInt(1) << node_level);
}
template <typename F, std::size_t NDim, typename UInt>
inline void get_node_centre(F (&out)[NDim], UInt node_code, F box_size)
{
    static_assert(NDim == 3, "only NDim = 3 is supported");
    constexpr UInt cbits = cbits_v<UInt, NDim>;
    const UInt lvl = tree_level<NDim>(node_code);
    const UInt first = static_cast<UInt>((node_code - (UInt(1) << (lvl * NDim))) << ((cbits - lvl) * NDim));
    const F half_dim = get_node_dim(lvl, box_size) * (F(1) / F(2));
    const F cell = box_size * (F(1) / static_cast<F>(UInt(1) << cbits));
    const UInt d[3] = {static_cast<UInt>(morton_compact3(first)), static_cast<UInt>(morton_compact3(first >> 1)),
                       static_cast<UInt>(morton_compact3(first >> 2))};
    for (std::size_t j = 0; j < NDim; ++j) {
        out[j] = fma_wrap(static_cast<F>(d[j]), cell, half_dim - box_size * (F(1) / F(2)));
    }
}

// Random-access iterator visiting base[idx[0]], base[idx[1]], ... (stands in for the
// boost::permutation_iterator the reference returns from p_its_o()/c_it_o(), tree.hpp:3621-3657).
template <typename BaseIt, typename IdxIt>
class perm_iterator
{
    BaseIt m_base{};
    IdxIt m_idx{};

public:
    using iterator_category = std::random_access_iterator_tag;
    using value_type = typename std::iterator_traits<BaseIt>::value_type;
    using difference_type = typename std::iterator_traits<IdxIt>::difference_type;
    using reference = typename std::iterator_traits<BaseIt>::reference;
    using pointer = typename std::iterator_traits<BaseIt>::pointer;

    perm_iterator() = default;
    perm_iterator(BaseIt b, IdxIt i) : m_base(b), m_idx(i) {}
    reference operator*() const { return *(m_base + static_cast<difference_type>(*m_idx)); }
    reference operator[](difference_type n) const { return *(m_base + static_cast<difference_type>(*(m_idx + n))); }
    perm_iterator &operator++()
    {
        ++m_idx;
        return *this;
    }
    perm_iterator operator++(int)
    {
        auto t = *this;
        ++m_idx;
        return t;
    }
    perm_iterator &operator--()
    {
        --m_idx;
        return *this;
    }
    perm_iterator operator--(int)
    {
        auto t = *this;
        --m_idx;
        return t;
    }
    perm_iterator &operator+=(difference_type n)
    {
        m_idx += n;
        return *this;
    }
    perm_iterator &operator-=(difference_type n)
    {
        m_idx -= n;
        return *this;
    }
    friend perm_iterator operator+(perm_iterator a, difference_type n) { return a += n; }
    friend perm_iterator operator+(difference_type n, perm_iterator a) { return a += n; }
    friend perm_iterator operator-(perm_iterator a, difference_type n) { return a -= n; }
    friend difference_type operator-(const perm_iterator &a, const perm_iterator &b) { return a.m_idx - b.m_idx; }
    friend bool operator==(const perm_iterator &a, const perm_iterator &b) { return a.m_idx == b.m_idx; }
    friend bool operator!=(const perm_iterator &a, const perm_iterator &b) { return a.m_idx != b.m_idx; }
    friend bool operator<(const perm_iterator &a, const perm_iterator &b) { return a.m_idx < b.m_idx; }
    friend bool operator>(const perm_iterator &a, const perm_iterator &b) { return a.m_idx > b.m_idx; }
    friend bool operator<=(const perm_iterator &a, const perm_iterator &b) { return a.m_idx <= b.m_idx; }
    friend bool operator>=(const perm_iterator &a, const perm_iterator &b) { return a.m_idx >= b.m_idx; }
};

// Range detection (begin/end found by ADL or std).
namespace adl_probe
{
using std::begin;
using std::end;
template <typename T>
auto test(int) -> decltype(begin(std::declval<T>()), end(std::declval<T>()), std::true_type{});
template <typename>
std::false_type test(...);
} // namespace adl_probe
template <typename T>
inline constexpr bool is_range_v = decltype(adl_probe::test<T>(0))::value;

template <typename T>
using uncvref_t = std::remove_cv_t<std::remove_reference_t<T>>;

// Checked numeric conversion for keyword-argument values (the reference uses boost::numeric_cast).
template <typename To, typename From>
inline To num_cast(const From &x)
{
    if constexpr (std::is_floating_point_v<To>) {
        return static_cast<To>(x);
    } else {
        if constexpr (std::is_floating_point_v<From>) {
            if (!(x >= From(0)) || x >= static_cast<From>(std::numeric_limits<To>::max())) {
                throw std::overflow_error("numeric conversion out of range");
            }
        } else if constexpr (std::is_signed_v<From>) {
            if (x < 0) {
                throw std::overflow_error("negative value in a numeric conversion to an unsigned type");
            }
        }
        return static_cast<To>(x);
    }
}

} // namespace detail

// Keyword arguments (reference tree.hpp:599-626).
namespace kwargs
{
struct box_size_tag;
struct max_leaf_n_tag;
struct ncrit_tag;
struct masses_tag;
struct nparts_tag;
struct G_tag;
struct eps_tag;
struct split_tag;
template <std::size_t>
struct coords_tag {
};

inline constexpr kw::keyword<box_size_tag> box_size{};
inline constexpr kw::keyword<max_leaf_n_tag> max_leaf_n{};
inline constexpr kw::keyword<ncrit_tag> ncrit{};
template <std::size_t N>
inline constexpr kw::keyword<coords_tag<N>> coords{};
inline constexpr const kw::keyword<coords_tag<0>> &x_coords = coords<0>;
inline constexpr const kw::keyword<coords_tag<1>> &y_coords = coords<1>;
inline constexpr const kw::keyword<coords_tag<2>> &z_coords = coords<2>;
inline constexpr kw::keyword<masses_tag> masses{};
inline constexpr kw::keyword<nparts_tag> nparts{};
inline constexpr kw::keyword<G_tag> G{};
inline constexpr kw::keyword<eps_tag> eps{};
inline constexpr kw::keyword<split_tag> split{};
} // namespace kwargs

// Vector type for floating-point data: default-initialising, 64-byte aligned allocator (tree.hpp:628-631).
template <typename F>
using f_vector = std::vector<F, di_aligned_allocator<F, 64>>;

template <std::size_t NDim, typename F, typename UInt, mac MAC>
class tree
{
    static_assert(NDim == 3, "rakau_b200 implements the 3-dimensional (octree) path only.");
    static_assert(std::is_same_v<F, float> || std::is_same_v<F, double>,
                  "rakau_b200 supports the float and double floating-point types only.");
    static_assert(std::is_integral_v<UInt> && std::is_unsigned_v<UInt> && std::numeric_limits<UInt>::digits == 64,
                  "rakau_b200 supports 64-bit unsigned Morton codes only.");
    static_assert(MAC >= mac::bh && MAC <= mac::bh_geom, "The selected MAC does not exist.");
    static constexpr int fp_bits = std::is_same_v<F, float> ? 32 : 64;
    static constexpr int mac_id = MAC == mac::bh ? RK_MAC_BH : RK_MAC_BH_GEOM;

public:
    using size_type = tree_size_t<F>;

private:
    using node_type = tree_node_t<NDim, F, UInt, MAC>;
    using tree_type = std::vector<node_type, di_aligned_allocator<node_type>>;
    using cnode_type = tree_cnode_t<F, UInt>;
    using cnode_list_type = std::vector<cnode_type, di_aligned_allocator<cnode_type>>;
    using idx_vector = std::vector<size_type, di_aligned_allocator<size_type>>;
    using code_vector = std::vector<UInt, di_aligned_allocator<UInt>>;
    using raw_node = std::conditional_t<fp_bits == 32, rk_node_f32, rk_node_f64>;
    template <unsigned Q>
    static constexpr std::size_t nvecs_res = tree_nvecs_res<Q, NDim>;

    // ---- C ABI plumbing ----------------------------------------------------------------------------------
    [[noreturn]] static void throw_status(int rc, const std::string &msg)
    {
        switch (rc) {
            case RK_ERR_INVALID_ARGUMENT:
                throw std::invalid_argument(msg);
            case RK_ERR_DOMAIN:
                throw std::domain_error(msg);
            case RK_ERR_OVERFLOW:
                throw std::overflow_error(msg);
            case RK_ERR_BAD_ALLOC:
                throw std::bad_alloc{};
            default:
                throw std::runtime_error(msg);
        }
    }
    void check(int rc) const
    {
        if (rc != RK_OK) {
            throw_status(rc, rk_last_error(m_h));
        }
    }
    void ensure_handle()
    {
        if (!m_h) {
            m_h = rk_tree_create(fp_bits, mac_id, 0);
            if (!m_h) {
                throw std::runtime_error(rk_create_error());
            }
        }
    }
    void invalidate_mirrors() const noexcept
    {
        m_have_parts = m_have_codes = m_have_perms = m_have_nodes = false;
    }
    // Lazy host mirrors (the reference API returns raw host pointers / references to host vectors).
    void fetch_parts() const
    {
        std::lock_guard<std::mutex> lock(m_mut);
        if (m_have_parts) {
            return;
        }
        const auto n = nparts();
        for (auto &p : m_parts) {
            p.resize(n);
        }
        if (n) {
            check(rk_tree_get_parts(m_h, m_parts[0].data(), m_parts[1].data(), m_parts[2].data(), m_parts[3].data()));
        }
        m_have_parts = true;
    }
    void fetch_codes() const
    {
        std::lock_guard<std::mutex> lock(m_mut);
        if (m_have_codes) {
            return;
        }
        m_codes.resize(nparts());
        if (nparts()) {
            static_assert(sizeof(UInt) == sizeof(std::uint64_t));
            check(rk_tree_get_codes(m_h, reinterpret_cast<std::uint64_t *>(m_codes.data())));
        }
        m_have_codes = true;
    }
    void fetch_perms() const
    {
        std::lock_guard<std::mutex> lock(m_mut);
        if (m_have_perms) {
            return;
        }
        static_assert(sizeof(size_type) == sizeof(std::uint64_t));
        const auto n = nparts();
        m_perm.resize(n);
        m_last_perm.resize(n);
        m_inv_perm.resize(n);
        if (n) {
            check(rk_tree_get_perm(m_h, RK_PERM, reinterpret_cast<std::uint64_t *>(m_perm.data())));
            check(rk_tree_get_perm(m_h, RK_LAST_PERM, reinterpret_cast<std::uint64_t *>(m_last_perm.data())));
            check(rk_tree_get_perm(m_h, RK_INV_PERM, reinterpret_cast<std::uint64_t *>(m_inv_perm.data())));
        }
        m_have_perms = true;
    }
    void fetch_nodes() const
    {
        std::lock_guard<std::mutex> lock(m_mut);
        if (m_have_nodes) {
            return;
        }
        const std::size_t M = m_h ? rk_tree_nnodes(m_h) : 0u, C = m_h ? rk_tree_ncrit(m_h) : 0u;
        m_tree.resize(M);
        m_crit_nodes.resize(C);
        if (M) {
            std::vector<raw_node> raw(M);
            check(rk_tree_get_nodes(m_h, raw.data()));
            for (std::size_t i = 0; i < M; ++i) {
                node_type &n = m_tree[i];
                n.begin = raw[i].begin;
                n.end = raw[i].end;
                n.n_children = raw[i].n_children;
                n.code = static_cast<UInt>(raw[i].code);
                n.level = static_cast<UInt>(raw[i].level);
                for (std::size_t j = 0; j < NDim + 1u; ++j) {
                    n.props[j] = raw[i].props[j];
                }
                if constexpr (MAC == mac::bh) {
                    n.dim2 = raw[i].dim;
                } else {
                    n.dim = raw[i].dim;
                    n.delta = raw[i].delta;
                }
            }
            std::vector<rk_cnode> cr(C);
            check(rk_tree_get_crit(m_h, cr.data()));
            for (std::size_t i = 0; i < C; ++i) {
                m_crit_nodes[i] = cnode_type{static_cast<UInt>(cr[i].code), cr[i].begin, cr[i].end};
            }
        }
        m_have_nodes = true;
    }

    // Contiguous F data for an iterator range: pointers are used in place, anything else is copied.
    template <typename It>
    static const F *contiguous(It it, size_type n, std::vector<F> &tmp)
    {
        if constexpr (std::is_pointer_v<It> && std::is_same_v<uncvref_t<decltype(*it)>, F>) {
            (void)n;
            (void)tmp;
            return it;
        } else {
            tmp.resize(n);
            for (size_type i = 0; i < n; ++i) {
                tmp[i] = static_cast<F>(*(it + static_cast<typename std::iterator_traits<It>::difference_type>(i)));
            }
            return tmp.data();
        }
    }

    // construct_impl, tree.hpp:1329-1487.
    template <typename It>
    void construct_impl(F box_size, bool box_size_deduced, const std::array<It, NDim + 1u> &its, size_type N,
                        size_type max_leaf_n, size_type ncrit)
    {
        simple_timer st("overall tree construction");
        m_box_size = box_size;
        m_box_size_deduced = box_size_deduced;
        m_max_leaf_n = max_leaf_n;
        m_ncrit = ncrit;
        ensure_handle();
        invalidate_mirrors();
        std::vector<F> tmp[NDim + 1u];
        const F *ptrs[NDim + 1u];
        for (std::size_t j = 0; j < NDim + 1u; ++j) {
            ptrs[j] = contiguous(its[j], N, tmp[j]);
        }
        rk_build_info info;
        const int rc = rk_tree_build(m_h, ptrs[0], ptrs[1], ptrs[2], ptrs[3], N, RK_HOST, static_cast<double>(box_size),
                                     box_size_deduced ? 1 : 0, max_leaf_n, ncrit, &info);
        if (rc != RK_OK) {
            const std::string msg = rk_last_error(m_h);
            clear();
            throw_status(rc, msg);
        }
        m_box_size = static_cast<F>(info.box_size);
        report_build(info);
    }
    // Device-side phases of one build (CUDA-event times), under the reference's phase names.
    static void report_build(const rk_build_info &info)
    {
        simple_timer::report_ms("morton encoding", info.ms_encode);
        simple_timer::report_ms("indirect code sorting", info.ms_sort);
        simple_timer::report_ms("permute", info.ms_permute);
        simple_timer::report_ms("node building", info.ms_topology + info.ms_props);
        (void)info;
    }

    template <typename... KwArgs>
    struct generic_ctor_enabled : std::true_type {
    };
    template <typename T>
    struct generic_ctor_enabled<T> : std::negation<std::is_same<tree, uncvref_t<T>>> {
    };

public:
    // Default constructor (tree.hpp:1523-1527).
    tree() : m_box_size(0), m_box_size_deduced(false), m_max_leaf_n(default_max_leaf_n), m_ncrit(default_ncrit) {}

    // Generic constructor with keyword arguments (tree.hpp:1573-1733).
    template <typename... KwArgs, std::enable_if_t<generic_ctor_enabled<KwArgs &&...>::value, int> = 0>
    explicit tree(KwArgs &&... args) : tree()
    {
        kw::parser p{args...};
        static_assert(!p.has_duplicates(), "The generic constructor cannot have duplicate keyword arguments.");
        static_assert(!p.has_unnamed_arguments(),
                      "All the arguments for the generic constructor must be keyword arguments.");
        static_assert(p.has_all(kwargs::coords<0>, kwargs::coords<1>, kwargs::coords<2>) && p.has(kwargs::masses),
                      "The generic tree constructor needs particle coordinates for every dimension, and particle "
                      "masses.");
        F bsize(0);
        bool deduced = true;
        if constexpr (p.has(kwargs::box_size)) {
            bsize = num_cast<F>(p(kwargs::box_size));
            deduced = false;
        }
        size_type mln = default_max_leaf_n, nc = default_ncrit;
        if constexpr (p.has(kwargs::max_leaf_n)) {
            mln = num_cast<size_type>(p(kwargs::max_leaf_n));
        }
        if constexpr (p.has(kwargs::ncrit)) {
            nc = num_cast<size_type>(p(kwargs::ncrit));
        }
        using p_data_t = uncvref_t<decltype(p(kwargs::coords<0>))>;
        static_assert(std::is_same_v<p_data_t, uncvref_t<decltype(p(kwargs::coords<1>))>>
                          && std::is_same_v<p_data_t, uncvref_t<decltype(p(kwargs::coords<2>))>>,
                      "All particle data in the generic tree constructor must be passed in as the same type.");
        static_assert(std::is_same_v<p_data_t, uncvref_t<decltype(p(kwargs::masses))>>,
                      "The type of the particle masses data is not consistent with the type of the particle "
                      "coordinates data.");
        if constexpr (is_range_v<const p_data_t &>) {
            static_assert(!p.has(kwargs::nparts), "If the particle coordinates are provided as ranges, the "
                                                  "'nparts' keyword argument must not be provided.");
            using std::begin;
            using std::end;
            const p_data_t &rx = p(kwargs::coords<0>), &ry = p(kwargs::coords<1>), &rz = p(kwargs::coords<2>),
                           &rm = p(kwargs::masses);
            const auto N = num_cast<size_type>(std::distance(begin(rx), end(rx)));
            if (num_cast<size_type>(std::distance(begin(ry), end(ry))) != N
                || num_cast<size_type>(std::distance(begin(rz), end(rz))) != N) {
                throw std::invalid_argument("The input ranges for the particle coordinates have inconsistent sizes");
            }
            const auto msize = num_cast<size_type>(std::distance(begin(rm), end(rm)));
            if (msize != N) {
                throw std::invalid_argument("The size of the input range for the particle masses ("
                                            + std::to_string(msize)
                                            + ") is different from the size of "
                                              "the input ranges for the particle coordinates ("
                                            + std::to_string(N) + ")");
            }
            construct_impl(bsize, deduced, std::array{begin(rx), begin(ry), begin(rz), begin(rm)}, N, mln, nc);
        } else {
            static_assert(p.has(kwargs::nparts), "If the particle coordinates are provided as iterators, the "
                                                 "'nparts' keyword argument must also be provided.");
            const auto N = num_cast<size_type>(p(kwargs::nparts));
            construct_impl(bsize, deduced,
                           std::array<p_data_t, NDim + 1u>{p(kwargs::coords<0>), p(kwargs::coords<1>),
                                                           p(kwargs::coords<2>), p(kwargs::masses)},
                           N, mln, nc);
        }
    }

    // Copy / move (tree.hpp:1735-1829).
    tree(const tree &other)
        : m_box_size(other.m_box_size), m_box_size_deduced(other.m_box_size_deduced), m_max_leaf_n(other.m_max_leaf_n),
          m_ncrit(other.m_ncrit)
    {
        if (other.m_h) {
            ensure_handle();
            const int rc = rk_tree_clone(m_h, other.m_h);
            if (rc != RK_OK) {
                const std::string msg = rk_last_error(m_h);
                rk_tree_destroy(m_h);
                m_h = nullptr;
                throw_status(rc, msg);
            }
        }
    }
    tree(tree &&other) noexcept
        : m_h(other.m_h), m_box_size(other.m_box_size), m_box_size_deduced(other.m_box_size_deduced),
          m_max_leaf_n(other.m_max_leaf_n), m_ncrit(other.m_ncrit)
    {
        other.m_h = nullptr;
        other.clear();
    }
    tree &operator=(const tree &other)
    {
        if (this != &other) {
            try {
                tree tmp(other);
                *this = std::move(tmp);
            } catch (...) {
                clear();
                throw;
            }
        }
        return *this;
    }
    tree &operator=(tree &&other) noexcept
    {
        if (this != &other) {
            if (m_h) {
                rk_tree_destroy(m_h);
            }
            m_h = other.m_h;
            other.m_h = nullptr;
            m_box_size = other.m_box_size;
            m_box_size_deduced = other.m_box_size_deduced;
            m_max_leaf_n = other.m_max_leaf_n;
            m_ncrit = other.m_ncrit;
            invalidate_mirrors();
            other.clear();
        }
        return *this;
    }
    ~tree()
    {
        if (m_h) {
            rk_tree_destroy(m_h);
        }
    }

    // Reset to the default-constructed state (tree.hpp:1882-1904).
    void clear() noexcept
    {
        m_box_size = F(0);
        m_box_size_deduced = false;
        m_max_leaf_n = default_max_leaf_n;
        m_ncrit = default_ncrit;
        if (m_h) {
            rk_tree_clear(m_h);
        }
        for (auto &p : m_parts) {
            p.clear();
        }
        m_codes.clear();
        m_perm.clear();
        m_last_perm.clear();
        m_inv_perm.clear();
        m_tree.clear();
        m_crit_nodes.clear();
        invalidate_mirrors();
    }

    // Pretty printing (tree.hpp:1906-1958).
    std::ostream &pprint(std::ostream &os, size_type max_nodes = 0) const
    {
        fetch_nodes();
        const auto n_nodes = m_tree.size();
        os << "Box size                 : " << m_box_size << (m_box_size_deduced ? " (deduced)" : "") << '\n';
        os << "Total number of particles: " << nparts() << '\n';
        os << "Total number of nodes    : " << n_nodes << "\n\n";
        if (!n_nodes) {
            return os;
        }
        os << ((max_nodes && max_nodes < n_nodes) ? "First " + std::to_string(max_nodes) + " nodes:\n" : "Nodes:\n");
        size_type shown = 0;
        for (const auto &nd : m_tree) {
            os << std::bitset<std::numeric_limits<UInt>::digits>(nd.code) << '|' << nd.begin << ',' << nd.end << ','
               << nd.n_children << "|" << nd.props[NDim] << "|[" << nd.props[0] << ", " << nd.props[1] << ", "
               << nd.props[2] << "]\n";
            if (++shown == max_nodes) {
                break;
            }
        }
        if (shown < n_nodes) {
            os << "...\n";
        }
        return os;
    }
    friend std::ostream &operator<<(std::ostream &os, const tree &t) { return t.pprint(os, 20); }

private:
    // parse_accpot_kwargs, tree.hpp:3376-3403.
    template <typename... Args>
    static auto parse_accpot_kwargs(Args &&... args)
    {
        kw::parser p{args...};
        static_assert(!p.has_duplicates(), "The functions for the computation of accelerations and/or potentials "
                                           "cannot have duplicate keyword arguments.");
        static_assert(!p.has_unnamed_arguments(), "Only keyword arguments can be passed in the parameter pack of the "
                                                  "functions for the computation of accelerations and potentials");
        F G(1), eps(0);
        if constexpr (p.has(kwargs::G)) {
            G = num_cast<F>(p(kwargs::G));
        }
        if constexpr (p.has(kwargs::eps)) {
            eps = num_cast<F>(p(kwargs::eps));
        }
        if constexpr (p.has(kwargs::split)) {
            return std::tuple<F, F, std::vector<double>>{G, eps, std::vector<double>(p(kwargs::split))};
        } else {
            return std::tuple<F, F, std::vector<double>>{G, eps, std::vector<double>{}};
        }
    }

    // acc_pot_dispatch, tree.hpp:3293-3334: all argument checks live behind the C ABI (same messages).
    template <bool Ordered, unsigned Q, typename It>
    void acc_pot_dispatch(const std::array<It, nvecs_res<Q>> &out, F theta, F G, F eps,
                          const std::vector<double> &split) const
    {
        simple_timer st("vector accs/pots computation");
        if (!m_h) {
            const_cast<tree *>(this)->ensure_handle(); // an empty tree still validates its arguments
        }
        const auto n = nparts();
        void *ptrs[4] = {nullptr, nullptr, nullptr, nullptr};
        if constexpr (std::is_same_v<It, F *>) {
            for (std::size_t j = 0; j < nvecs_res<Q>; ++j) {
                ptrs[j] = out[j];
            }
            check(rk_tree_acc_pot(m_h, static_cast<int>(Q), Ordered ? 1 : 0, theta, G, eps, split.data(), split.size(),
                                  ptrs, RK_HOST, nullptr));
        } else {
            std::array<std::vector<F>, nvecs_res<Q>> tmp;
            for (std::size_t j = 0; j < nvecs_res<Q>; ++j) {
                tmp[j].resize(n);
                ptrs[j] = tmp[j].data();
            }
            check(rk_tree_acc_pot(m_h, static_cast<int>(Q), Ordered ? 1 : 0, theta, G, eps, split.data(), split.size(),
                                  ptrs, RK_HOST, nullptr));
            for (std::size_t j = 0; j < nvecs_res<Q>; ++j) {
                auto o = out[j];
                for (size_type i = 0; i < n; ++i, ++o) {
                    *o = tmp[j][i];
                }
            }
        }
    }
    template <bool Ordered, unsigned Q, typename Allocator>
    void acc_pot_dispatch(std::array<std::vector<F, Allocator>, nvecs_res<Q>> &out, F theta, F G, F eps,
                          const std::vector<double> &split) const
    {
        std::array<F *, nvecs_res<Q>> ptrs;
        for (std::size_t j = 0; j < nvecs_res<Q>; ++j) {
            out[j].resize(nparts());
            ptrs[j] = out[j].data();
        }
        acc_pot_dispatch<Ordered, Q>(ptrs, theta, G, eps, split);
    }
    template <bool Ordered, unsigned Q, typename Allocator>
    void acc_pot_dispatch(std::vector<F, Allocator> &out, F theta, F G, F eps, const std::vector<double> &split) const
    {
        static_assert(Q == 1u);
        out.resize(nparts());
        acc_pot_dispatch<Ordered, Q>(std::array<F *, 1>{out.data()}, theta, G, eps, split);
    }
    template <unsigned Q, typename It>
    static auto ilist_to_array(std::initializer_list<It> il)
    {
        if (il.size() != nvecs_res<Q>) {
            throw std::invalid_argument(
                "An initializer list containing " + std::to_string(il.size())
                + " iterators was used as the output for the computation of the accelerations/potentials in a "
                + std::to_string(NDim) + "-dimensional tree, but a list with " + std::to_string(nvecs_res<Q>)
                + " iterators is required instead");
        }
        std::array<It, nvecs_res<Q>> r;
        std::copy(il.begin(), il.end(), r.begin());
        return r;
    }

public:
    // Accelerations / potentials, unordered (Morton order) and ordered (original order), tree.hpp:3405-3497.
#define RAKAU_B200_ACCPOT(NAME, ORDERED, Q, NOUT)                                                                      \
    template <typename Allocator, typename... KwArgs>                                                                  \
    void NAME(std::array<std::vector<F, Allocator>, NOUT> &out, F mac_value, KwArgs &&... args) const                  \
    {                                                                                                                  \
        const auto [G, eps, split] = parse_accpot_kwargs(std::forward<KwArgs>(args)...);                               \
        acc_pot_dispatch<ORDERED, Q>(out, mac_value, G, eps, split);                                                   \
    }                                                                                                                  \
    template <typename It, typename... KwArgs>                                                                         \
    void NAME(const std::array<It, NOUT> &out, F mac_value, KwArgs &&... args) const                                   \
    {                                                                                                                  \
        const auto [G, eps, split] = parse_accpot_kwargs(std::forward<KwArgs>(args)...);                               \
        acc_pot_dispatch<ORDERED, Q>(out, mac_value, G, eps, split);                                                   \
    }                                                                                                                  \
    template <typename It, typename... KwArgs>                                                                         \
    void NAME(std::initializer_list<It> out, F mac_value, KwArgs &&... args) const                                     \
    {                                                                                                                  \
        NAME(ilist_to_array<Q>(out), mac_value, std::forward<KwArgs>(args)...);                                        \
    }
    RAKAU_B200_ACCPOT(accs_u, false, 0, NDim)
    RAKAU_B200_ACCPOT(accs_o, true, 0, NDim)
    RAKAU_B200_ACCPOT(accs_pots_u, false, 2, NDim + 1u)
    RAKAU_B200_ACCPOT(accs_pots_o, true, 2, NDim + 1u)
#undef RAKAU_B200_ACCPOT
#define RAKAU_B200_POT(NAME, ORDERED)                                                                                  \
    template <typename Allocator, typename... KwArgs>                                                                  \
    void NAME(std::vector<F, Allocator> &out, F mac_value, KwArgs &&... args) const                                    \
    {                                                                                                                  \
        const auto [G, eps, split] = parse_accpot_kwargs(std::forward<KwArgs>(args)...);                               \
        acc_pot_dispatch<ORDERED, 1>(out, mac_value, G, eps, split);                                                   \
    }                                                                                                                  \
    template <typename It, typename... KwArgs, std::enable_if_t<!is_range_v<It &>, int> = 0>                           \
    void NAME(It out, F mac_value, KwArgs &&... args) const                                                            \
    {                                                                                                                  \
        const auto [G, eps, split] = parse_accpot_kwargs(std::forward<KwArgs>(args)...);                               \
        acc_pot_dispatch<ORDERED, 1>(std::array<It, 1>{out}, mac_value, G, eps, split);                                \
    }
    RAKAU_B200_POT(pots_u, false)
    RAKAU_B200_POT(pots_o, true)
#undef RAKAU_B200_POT

private:
    // exact_acc_pot_impl, tree.hpp:3531-3569 (direct summation on the GPU).
    template <bool Ordered, typename... KwArgs>
    std::array<F, 4> exact_impl(size_type idx, KwArgs &&... args) const
    {
        const auto [G, eps, split] = parse_accpot_kwargs(std::forward<KwArgs>(args)...);
        (void)split;
        if (!m_h) {
            throw std::invalid_argument("exact_*: the tree is empty");
        }
        simple_timer st("exact acc/pot computation");
        double out[4];
        check(rk_tree_exact(m_h, idx, Ordered ? 1 : 0, G, eps, out));
        return {static_cast<F>(out[0]), static_cast<F>(out[1]), static_cast<F>(out[2]), static_cast<F>(out[3])};
    }

public:
    template <typename... KwArgs>
    std::array<F, NDim> exact_acc_u(size_type idx, KwArgs &&... args) const
    {
        const auto r = exact_impl<false>(idx, std::forward<KwArgs>(args)...);
        return {r[0], r[1], r[2]};
    }
    template <typename... KwArgs>
    F exact_pot_u(size_type idx, KwArgs &&... args) const
    {
        return exact_impl<false>(idx, std::forward<KwArgs>(args)...)[3];
    }
    template <typename... KwArgs>
    std::array<F, NDim + 1u> exact_acc_pot_u(size_type idx, KwArgs &&... args) const
    {
        return exact_impl<false>(idx, std::forward<KwArgs>(args)...);
    }
    template <typename... KwArgs>
    std::array<F, NDim> exact_acc_o(size_type idx, KwArgs &&... args) const
    {
        const auto r = exact_impl<true>(idx, std::forward<KwArgs>(args)...);
        return {r[0], r[1], r[2]};
    }
    template <typename... KwArgs>
    F exact_pot_o(size_type idx, KwArgs &&... args) const
    {
        return exact_impl<true>(idx, std::forward<KwArgs>(args)...)[3];
    }
    template <typename... KwArgs>
    std::array<F, NDim + 1u> exact_acc_pot_o(size_type idx, KwArgs &&... args) const
    {
        return exact_impl<true>(idx, std::forward<KwArgs>(args)...);
    }

    // Iterators into the particle data and codes, permutations, nodes (tree.hpp:3637-3673).
    std::array<const F *, NDim + 1u> p_its_u() const
    {
        fetch_parts();
        return {m_parts[0].data(), m_parts[1].data(), m_parts[2].data(), m_parts[3].data()};
    }
    auto p_its_o() const
    {
        fetch_parts();
        fetch_perms();
        using it_t = perm_iterator<const F *, const size_type *>;
        return std::array<it_t, NDim + 1u>{
            it_t(m_parts[0].data(), m_inv_perm.data()), it_t(m_parts[1].data(), m_inv_perm.data()),
            it_t(m_parts[2].data(), m_inv_perm.data()), it_t(m_parts[3].data(), m_inv_perm.data())};
    }
    const UInt *c_it_u() const
    {
        fetch_codes();
        return m_codes.data();
    }
    auto c_it_o() const
    {
        fetch_codes();
        fetch_perms();
        return perm_iterator<const UInt *, const size_type *>(m_codes.data(), m_inv_perm.data());
    }
    const idx_vector &perm() const
    {
        fetch_perms();
        return m_perm;
    }
    const idx_vector &last_perm() const
    {
        fetch_perms();
        return m_last_perm;
    }
    const idx_vector &inv_perm() const
    {
        fetch_perms();
        return m_inv_perm;
    }
    const tree_type &nodes() const
    {
        fetch_nodes();
        return m_tree;
    }
    // Not part of the reference's public API (its critical-node list is private): exposed for parity tests.
    const cnode_list_type &crit_nodes() const
    {
        fetch_nodes();
        return m_crit_nodes;
    }

private:
    // update_particles_dispatch + sync, tree.hpp:3678-3765.
    template <bool Ordered, typename Func>
    void update_particles_dispatch(Func &&f)
    {
        simple_timer st("overall update_particles");
        try {
            fetch_parts();
            if constexpr (Ordered) {
                fetch_perms();
                using it_t = perm_iterator<F *, const size_type *>;
                std::forward<Func>(f)(std::array<it_t, NDim + 1u>{
                    it_t(m_parts[0].data(), m_inv_perm.data()), it_t(m_parts[1].data(), m_inv_perm.data()),
                    it_t(m_parts[2].data(), m_inv_perm.data()), it_t(m_parts[3].data(), m_inv_perm.data())});
            } else {
                std::forward<Func>(f)(std::array<F *, NDim + 1u>{m_parts[0].data(), m_parts[1].data(),
                                                                 m_parts[2].data(), m_parts[3].data()});
            }
            if (m_h && nparts()) {
                rk_build_info info;
                check(rk_tree_update_positions(m_h, m_parts[0].data(), m_parts[1].data(), m_parts[2].data(),
                                               m_parts[3].data(), RK_HOST, &info));
                m_box_size = static_cast<F>(info.box_size);
                report_build(info);
            }
            invalidate_mirrors();
        } catch (...) {
            clear();
            throw;
        }
    }
    // update_masses_dispatch, tree.hpp:3782-3805.
    template <bool Ordered, typename Func>
    void update_masses_dispatch(Func &&f)
    {
        simple_timer st("overall update_masses");
        try {
            fetch_parts();
            if constexpr (Ordered) {
                fetch_perms();
                std::forward<Func>(f)(perm_iterator<F *, const size_type *>(m_parts[3].data(), m_inv_perm.data()));
            } else {
                std::forward<Func>(f)(m_parts[3].data());
            }
            if (m_h && nparts()) {
                check(rk_tree_update_masses(m_h, m_parts[3].data(), RK_HOST));
            }
            m_have_nodes = false;
        } catch (...) {
            clear();
            throw;
        }
    }

public:
    template <typename Func>
    void update_particles_u(Func &&f)
    {
        update_particles_dispatch<false>(std::forward<Func>(f));
    }
    template <typename Func>
    void update_particles_o(Func &&f)
    {
        update_particles_dispatch<true>(std::forward<Func>(f));
    }
    template <typename Func>
    void update_masses_u(Func &&f)
    {
        update_masses_dispatch<false>(std::forward<Func>(f));
    }
    template <typename Func>
    void update_masses_o(Func &&f)
    {
        update_masses_dispatch<true>(std::forward<Func>(f));
    }

    // ---- extension (no counterpart in the reference's class): the time loop of the reference's
    // benchmark/benchmark_leapfrog.cpp:252-384 with positions and velocities resident on the GPU -------------------
    // Velocities in the ORIGINAL particle order; theta / G / eps apply to every later step.
    template <typename... KwArgs>
    void leapfrog_init(const F *vx, const F *vy, const F *vz, F mac_value, bool track_integrals, KwArgs &&... args)
    {
        const auto [G, eps, split] = parse_accpot_kwargs(std::forward<KwArgs>(args)...);
        (void)split;
        if (!m_h) {
            throw std::invalid_argument("leapfrog_init: the tree is empty");
        }
        check(rk_tree_leapfrog_init(m_h, vx, vy, vz, RK_HOST, mac_value, G, eps, track_integrals ? 1 : 0));
    }
    // One kick-drift-kick step (tree rebuild included); returns the phase times and, with track_integrals, the centre
    // of mass, its velocity and the total energy at the beginning of the step.
    rk_leapfrog_info leapfrog_step(F timestep)
    {
        rk_leapfrog_info info;
        const int rc = rk_tree_leapfrog_step(m_h, timestep, &info);
        invalidate_mirrors();
        if (rc != RK_OK) {
            const std::string msg = rk_last_error(m_h);
            clear();
            throw_status(rc, msg);
        }
        return info;
    }
    // Velocities in the tree's internal (Morton) order, like p_its_u().
    std::array<std::vector<F>, NDim> leapfrog_velocities_u() const
    {
        std::array<std::vector<F>, NDim> v;
        for (auto &a : v) {
            a.resize(nparts());
        }
        check(rk_tree_leapfrog_get(m_h, 0, v[0].data(), v[1].data(), v[2].data(), RK_HOST));
        return v;
    }

    // Getters (tree.hpp:3818-3837).
    F box_size() const { return m_box_size; }
    bool box_size_deduced() const { return m_box_size_deduced; }
    size_type max_leaf_n() const { return m_max_leaf_n; }
    size_type ncrit() const { return m_ncrit; }
    size_type nparts() const { return m_h ? rk_tree_nparts(m_h) : 0u; }

private:
    rk_tree *m_h = nullptr;
    F m_box_size;
    bool m_box_size_deduced;
    size_type m_max_leaf_n, m_ncrit;
    // host mirrors, fetched on demand
    mutable std::mutex m_mut;
    mutable bool m_have_parts = false, m_have_codes = false, m_have_perms = false, m_have_nodes = false;
    mutable std::array<f_vector<F>, NDim + 1u> m_parts;
    mutable code_vector m_codes;
    mutable idx_vector m_perm, m_last_perm, m_inv_perm;
    mutable tree_type m_tree;
    mutable cnode_list_type m_crit_nodes;
};

template <typename F, mac MAC = mac::bh>
using octree = tree<3, F, std::size_t, MAC>;

} // namespace rakau

#endif
