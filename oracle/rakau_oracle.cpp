// rakau_oracle.cpp — CPU restatement of rakau's Barnes-Hut hot path.
//
// *** TEST INFRASTRUCTURE ONLY ***  This file is the parity oracle for the CUDA
// path in rakau_b200/csrc. Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load it. The product
// (librakau_b200.so, include/rakau/tree.hpp) never links or calls it.
//
// It restates, in scalar C++17, the algorithm of the reference
// (/root/reference/include/rakau/tree.hpp, "tree.hpp" below). Each function cites
// the reference lines it follows. Arithmetic contract: built with
// -ffp-contract=off; std::fma is used exactly where the reference writes
// fma_wrap()/xsimd_fma() (tree.hpp:181-207 — a real FMA when FP_FAST_FMA[F] is
// defined, i.e. any -mfma build), plain operators elsewhere. Where the reference
// has a SIMD and a scalar branch, the SCALAR branch is the one restated.
// Sort ties: the reference uses the unstable tbb::parallel_sort (tree.hpp:1271);
// the canonical order here is the stable one (a legal outcome of the reference).
//
// Parity pin: tests/test_oracle_golden.py checks this file against the known-answer
// vectors of the reference's own tests (test/node_centre.cpp, test/basic.cpp,
// test/auto_box_size.cpp, test/morton.cpp, ...); tests/test_oracle_vs_ref.py checks
// it BIT FOR BIT against oracle/_ref/libref_scalar.so, the unmodified reference
// header compiled with -DRAKAU_DISABLE_SIMD against the stand-ins in oracle/ref_shim.
//
// The oracle additionally carries the counters the reference lacks: per target
// group #MAC tests, #accepted nodes, #leaf P2P pairs, #self pairs.

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace
{

using u64 = std::uint64_t;
constexpr unsigned CBITS = 21; // cbits_v<uint64_t,3>, detail/tree_fwd.hpp:141-150

// ---------------------------------------------------------------------------
// Morton encoding. Bit layout of libmorton's m3D_e_sLUT (detail/libmorton/
// morton3D.h:38-49): x -> bits 0,3,6,..., y -> bits 1,4,7,..., z -> bits 2,5,8,...
// ---------------------------------------------------------------------------
inline u64 spread3(u64 v)
{
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
inline u64 compact3(u64 v)
{
    v &= 0x1249249249249249ull;
    v = (v ^ (v >> 2)) & 0x10c30c30c30c30c3ull;
    v = (v ^ (v >> 4)) & 0x100f00f00f00f00full;
    v = (v ^ (v >> 8)) & 0x1f0000ff0000ffull;
    v = (v ^ (v >> 16)) & 0x1f00000000ffffull;
    v = (v ^ (v >> 32)) & 0x1fffffull;
    return v;
}
inline u64 morton3(u64 x, u64 y, u64 z)
{
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}

// tree_level(), detail/tree_fwd.hpp:208-226.
inline unsigned node_level(u64 nodal_code)
{
    return (63u - static_cast<unsigned>(__builtin_clzll(nodal_code))) / 3u;
}

template <typename F>
struct node_t {
    // Field order/meaning of base_tree_node_t + tree_node_t, detail/tree_fwd.hpp:76-116.
    u64 begin, end, n_children, code, level;
    F props[4]; // com x,y,z, mass
    F dim;      // dim2 (mac::bh) or dim (mac::bh_geom)
    F delta;    // bh_geom only (0 for bh)
};

struct cnode_t { // tree_cnode_t, detail/tree_fwd.hpp:119-125
    u64 code, begin, end;
};

struct counters_t {
    u64 mac_tests = 0, accepted = 0, p2p_pairs = 0, self_pairs = 0, interactions = 0, leaves_opened = 0;
};

struct oracle_error : std::runtime_error {
    int code; // 1 invalid_argument, 2 domain_error, 3 overflow
    oracle_error(int c, const std::string &s) : std::runtime_error(s), code(c) {}
};

template <typename F>
struct tree_t {
    int mac = 0; // 0 bh, 1 bh_geom
    F box_size = 0;
    bool box_deduced = false;
    std::size_t max_leaf_n = 16, ncrit = 128;
    std::vector<F> parts[4]; // x,y,z,m in Morton order
    std::vector<u64> codes, perm, last_perm, inv_perm;
    std::vector<node_t<F>> nodes;
    std::vector<cnode_t> crit;
    std::vector<u64> crit_node_idx; // index in `nodes` of each critical node

    std::size_t n() const { return parts[0].size(); }

    // ---- disc_single_coord, tree.hpp:381-429 (Clamp = false) ----
    static u64 disc_coord(F x, F inv_box)
    {
        constexpr u64 factor = u64(1) << CBITS;
        F tmp = std::fma(x, inv_box, F(1) / F(2));
        tmp *= F(factor);
        if (!std::isfinite(tmp)) {
            throw oracle_error(1, "While trying to discretise the input coordinate " + std::to_string(x)
                                      + " in a box of size " + std::to_string(F(1) / inv_box)
                                      + ", the non-finite value " + std::to_string(tmp) + " was generated");
        }
        if (tmp < F(0) || tmp >= F(factor)) {
            throw oracle_error(1, "The discretisation of the input coordinate " + std::to_string(x)
                                      + " in a box of size " + std::to_string(F(1) / inv_box)
                                      + " produced the floating-point value " + std::to_string(tmp)
                                      + ", which is outside the allowed bounds");
        }
        const u64 r = static_cast<u64>(tmp);
        if (r >= factor) {
            throw oracle_error(1, "The discretisation of the input coordinate " + std::to_string(x)
                                      + " in a box of size " + std::to_string(F(1) / inv_box)
                                      + " produced the integral value " + std::to_string(r)
                                      + ", which is outside the allowed bounds");
        }
        return r;
    }

    // ---- determine_box_size, tree.hpp:1278-1319 ----
    F deduce_box() const
    {
        F mx = 0;
        for (int j = 0; j < 3; ++j) {
            for (F v : parts[j]) {
                const F a = std::abs(v);
                if (!std::isfinite(a)) {
                    throw oracle_error(1, "While trying to automatically determine the domain size, a "
                                          "non-finite coordinate with absolute value "
                                              + std::to_string(a) + " was encountered");
                }
                mx = std::max(mx, a);
            }
        }
        F b = mx * F(2);
        b = std::fma(b, F(1) / F(20), b);
        if (!std::isfinite(b)) {
            throw oracle_error(1, "The automatic deduction of the domain size produced the non-finite value "
                                      + std::to_string(b));
        }
        return b;
    }

    // ---- get_node_dim / get_node_centre, tree.hpp:443-482 ----
    F level_dim(u64 level) const { return box_size / static_cast<F>(u64(1) << level); }
    void node_centre(F out[3], u64 ncode) const
    {
        const unsigned lvl = node_level(ncode);
        const u64 first_cell = (ncode - (u64(1) << (lvl * 3u))) << ((CBITS - lvl) * 3u);
        const F half_dim = level_dim(lvl) * (F(1) / F(2));
        const F cell = box_size * (F(1) / static_cast<F>(u64(1) << CBITS));
        const u64 d[3] = {compact3(first_cell), compact3(first_cell >> 1), compact3(first_cell >> 2)};
        for (int j = 0; j < 3; ++j) {
            out[j] = std::fma(static_cast<F>(d[j]), cell, half_dim - box_size * (F(1) / F(2)));
        }
    }

    // ---- compute_node_properties, tree.hpp:1116-1237 (scalar tail 1162-1168) ----
    void node_props(node_t<F> &nd) const
    {
        F tot = 0, com[3] = {0, 0, 0};
        for (u64 i = nd.begin; i < nd.end; ++i) {
            const F mass = parts[3][i];
            tot += mass;
            for (int j = 0; j < 3; ++j) {
                com[j] = std::fma(mass, parts[j][i], com[j]);
            }
        }
        F geo[3] = {0, 0, 0};
        if (mac == 1) {
            node_centre(geo, nd.code);
        }
        if (tot == F(0)) {
            if (mac == 0) {
                node_centre(com, nd.code);
            } else {
                std::copy(geo, geo + 3, com);
            }
        } else {
            const F inv = F(1) / tot;
            for (int j = 0; j < 3; ++j) {
                com[j] *= inv;
            }
        }
        for (int j = 0; j < 3; ++j) {
            if (!std::isfinite(com[j])) {
                throw oracle_error(1, "The computation of the centre of mass of a node produced a non-finite value");
            }
            nd.props[j] = com[j];
        }
        if (!std::isfinite(tot)) {
            throw oracle_error(1, "The computation of the total mass in a node produced the non-finite value "
                                      + std::to_string(tot));
        }
        nd.props[3] = tot;
        const F nd_dim = level_dim(nd.level);
        if (mac == 0) {
            nd.dim = nd_dim * nd_dim;
            nd.delta = 0;
            if (!std::isfinite(nd.dim)) {
                throw oracle_error(
                    1, "The computation of the square of the dimension of a node produced the non-finite value "
                           + std::to_string(nd.dim));
            }
        } else {
            nd.dim = nd_dim;
            if (!std::isfinite(nd.dim)) {
                throw oracle_error(1, "The computation of the dimension of a node produced the non-finite value "
                                          + std::to_string(nd.dim));
            }
            F d2 = (com[0] - geo[0]) * (com[0] - geo[0]);
            for (int j = 1; j < 3; ++j) {
                d2 = std::fma(com[j] - geo[j], com[j] - geo[j], d2);
            }
            nd.delta = std::sqrt(d2);
            if (!std::isfinite(nd.delta)) {
                throw oracle_error(1, "The computation of the distance between the centre of mass "
                                      "and the geometric centre of a node produced the non-finite value "
                                          + std::to_string(nd.delta));
            }
        }
    }

    // ---- build_tree_ser_impl, tree.hpp:724-833: children of the node at index `parent_idx` ----
    // Returns the number of descendants appended.
    u64 build_children(std::size_t parent_idx, bool crit_ancestor)
    {
        const u64 pcode = nodes[parent_idx].code, plevel = nodes[parent_idx].level;
        if (plevel >= CBITS) {
            return 0;
        }
        const u64 pbegin = nodes[parent_idx].begin, pend = nodes[parent_idx].end;
        const unsigned shift = (CBITS - static_cast<unsigned>(plevel) - 1u) * 3u;
        const u64 prefix = pcode - (u64(1) << (plevel * 3u));
        u64 total = 0;
        const u64 *cb = codes.data() + pbegin, *ce = codes.data() + pend;
        for (u64 c = 0; c < 8; ++c) {
            const u64 want = (prefix << 3) + c;
            // std::equal_range on the shifted codes, tree.hpp:763-764.
            const u64 *lo = std::partition_point(cb, ce, [&](u64 v) { return (v >> shift) < want; });
            const u64 *hi = std::partition_point(lo, ce, [&](u64 v) { return (v >> shift) <= want; });
            const u64 np = static_cast<u64>(hi - lo);
            if (!np) {
                continue;
            }
            node_t<F> nd{};
            nd.begin = static_cast<u64>(lo - codes.data());
            nd.end = static_cast<u64>(hi - codes.data());
            nd.code = (pcode << 3) + c;
            nd.level = plevel + 1u;
            node_props(nd);
            nodes.push_back(nd);
            const std::size_t me = nodes.size() - 1u;
            // Critical-node rule, tree.hpp:801-807.
            const bool is_crit = !crit_ancestor && (np <= ncrit || np <= max_leaf_n || plevel + 1u == CBITS);
            if (is_crit) {
                crit.push_back({nd.code, nd.begin, nd.end});
                crit_node_idx.push_back(me);
            }
            if (np > max_leaf_n) {
                const u64 nc = build_children(me, is_crit || crit_ancestor);
                nodes[me].n_children = nc;
            }
            total += nodes[me].n_children + 1u;
        }
        return total;
    }

    // ---- build_tree, tree.hpp:932-1111 (root handling 956-979) ----
    void build_nodes()
    {
        nodes.clear();
        crit.clear();
        crit_node_idx.clear();
        const std::size_t N = n();
        if (!N) {
            return;
        }
        node_t<F> root{};
        root.begin = 0;
        root.end = N;
        root.code = 1;
        root.level = 0;
        node_props(root);
        nodes.push_back(root);
        const bool root_crit = N <= ncrit || N <= max_leaf_n;
        if (root_crit) {
            crit.push_back({u64(1), 0, static_cast<u64>(N)});
            crit_node_idx.push_back(0);
        }
        if (N > max_leaf_n) {
            const u64 nc = build_children(0, root_crit);
            nodes[0].n_children = nc;
        }
    }

    // Encode + stable indirect sort + permute; shared by construct_impl (tree.hpp:1436-1483)
    // and sync (tree.hpp:3678-3743).
    void encode_sort_permute(bool first_time)
    {
        const std::size_t N = n();
        if (box_deduced) {
            box_size = deduce_box();
        }
        const F inv_box = F(1) / box_size;
        for (std::size_t i = 0; i < N; ++i) {
            const u64 dx = disc_coord(parts[0][i], inv_box), dy = disc_coord(parts[1][i], inv_box),
                      dz = disc_coord(parts[2][i], inv_box);
            codes[i] = morton3(dx, dy, dz);
        }
        std::vector<u64> idx(N);
        std::iota(idx.begin(), idx.end(), u64(0));
        std::stable_sort(idx.begin(), idx.end(), [this](u64 a, u64 b) { return codes[a] < codes[b]; });
        auto gather = [&](auto &vec) {
            auto tmp = vec;
            for (std::size_t i = 0; i < N; ++i) {
                tmp[i] = vec[idx[i]];
            }
            vec.swap(tmp);
        };
        gather(codes);
        for (auto &p : parts) {
            gather(p);
        }
        if (first_time) {
            perm = idx;
        } else {
            gather(perm); // apply_isort(m_perm, m_last_perm), tree.hpp:3725-3727
        }
        last_perm = idx;
        for (std::size_t i = 0; i < N; ++i) {
            inv_perm[perm[i]] = i; // perm_to_inv_perm, tree.hpp:1248-1262
        }
    }

    // ---- construct_impl, tree.hpp:1329-1487 ----
    void construct(const F *x, const F *y, const F *z, const F *m, std::size_t N, F bsize, bool deduce,
                   std::size_t mln, std::size_t nc)
    {
        box_size = bsize;
        box_deduced = deduce;
        max_leaf_n = mln;
        ncrit = nc;
        if (!std::isfinite(box_size) || box_size < F(0)) {
            throw oracle_error(1, "The box size must be a finite non-negative value, but it is "
                                      + std::to_string(box_size) + " instead");
        }
        if (!max_leaf_n) {
            throw oracle_error(1, "The maximum number of particles per leaf must be nonzero");
        }
        if (!ncrit) {
            throw oracle_error(1, "The critical number of particles for the vectorised computation of the "
                                  "potentials/accelerations must be nonzero");
        }
        parts[0].assign(x, x + N);
        parts[1].assign(y, y + N);
        parts[2].assign(z, z + N);
        parts[3].assign(m, m + N);
        codes.assign(N, 0);
        perm.assign(N, 0);
        last_perm.assign(N, 0);
        inv_perm.assign(N, 0);
        encode_sort_permute(true);
        build_nodes();
    }

    void clear()
    {
        for (auto &p : parts) {
            p.clear();
        }
        codes.clear();
        perm.clear();
        last_perm.clear();
        inv_perm.clear();
        nodes.clear();
        crit.clear();
        crit_node_idx.clear();
        box_size = 0;
        box_deduced = false;
    }

    // ---- tree_self_interactions, scalar branch tree.hpp:2258-2320 ----
    template <unsigned Q>
    void self_interactions(F eps2, u64 tsize, const F *const tp[4], F *const res[4]) const
    {
        for (u64 i1 = 0; i1 < tsize; ++i1) {
            const F p1[3] = {tp[0][i1], tp[1][i1], tp[2][i1]};
            const F m1 = tp[3][i1];
            F a1[4] = {0, 0, 0, 0};
            for (u64 i2 = i1 + 1u; i2 < tsize; ++i2) {
                F d[3], dist2 = eps2;
                for (int j = 0; j < 3; ++j) {
                    d[j] = tp[j][i2] - p1[j];
                    dist2 = std::fma(d[j], d[j], dist2);
                }
                const F dist = std::sqrt(dist2), m2 = tp[3][i2];
                if constexpr (Q == 0u || Q == 2u) {
                    const F dist3 = dist2 * dist, m2d3 = m2 / dist3, m1d3 = m1 / dist3;
                    for (int j = 0; j < 3; ++j) {
                        a1[j] = std::fma(m2d3, d[j], a1[j]);
                        res[j][i2] = std::fma(m1d3, -d[j], res[j][i2]);
                    }
                }
                if constexpr (Q == 1u || Q == 2u) {
                    constexpr int pi = (Q == 1u) ? 0 : 3;
                    const F mut = m1 / dist * m2;
                    a1[pi] -= mut;
                    res[pi][i2] -= mut;
                }
            }
            if constexpr (Q == 0u || Q == 2u) {
                for (int j = 0; j < 3; ++j) {
                    res[j][i1] += a1[j];
                }
            }
            if constexpr (Q == 1u || Q == 2u) {
                constexpr int pi = (Q == 1u) ? 0 : 3;
                res[pi][i1] += a1[pi];
            }
        }
    }

    // ---- tree_acc_pot_leaf, scalar branch tree.hpp:2432-2470 ----
    template <unsigned Q>
    void leaf_p2p(F eps2, const node_t<F> &src, u64 tsize, const F *const tp[4], F *const res[4]) const
    {
        for (u64 i1 = 0; i1 < tsize; ++i1) {
            const F p1[3] = {tp[0][i1], tp[1][i1], tp[2][i1]};
            const F m1 = tp[3][i1];
            for (u64 i2 = src.begin; i2 < src.end; ++i2) {
                F d[3], dist2 = eps2;
                for (int j = 0; j < 3; ++j) {
                    d[j] = parts[j][i2] - p1[j];
                    dist2 = std::fma(d[j], d[j], dist2);
                }
                const F dist = std::sqrt(dist2), m2 = parts[3][i2];
                if constexpr (Q == 0u || Q == 2u) {
                    const F dist3 = dist * dist2, md3 = m2 / dist3;
                    for (int j = 0; j < 3; ++j) {
                        res[j][i1] = std::fma(d[j], md3, res[j][i1]);
                    }
                }
                if constexpr (Q == 1u || Q == 2u) {
                    constexpr int pi = (Q == 1u) ? 0 : 3;
                    res[pi][i1] = std::fma(-m1, m2 / dist, res[pi][i1]);
                }
            }
        }
    }

    // ---- tree_acc_pot_mac_check (scalar branch tree.hpp:2741-2777) + tree_acc_pot_src_com
    //      (scalar branch tree.hpp:2564-2589). Returns the next node index in the DFS. ----
    template <unsigned Q>
    u64 mac_check(u64 src_idx, F mac_value, F eps2, u64 tsize, const F *const tp[4], F *const res[4], F *tmp[5],
                  counters_t &cnt) const
    {
        const node_t<F> &src = nodes[src_idx];
        // mac_lh, tree.hpp:2632-2642.
        F mac_lh;
        if (mac == 0) {
            mac_lh = src.dim * mac_value;
        } else {
            const F t = std::fma(src.dim, mac_value, src.delta);
            mac_lh = t * t;
        }
        ++cnt.mac_tests;
        bool ok = true;
        for (u64 i = 0; i < tsize; ++i) {
            F dist2 = 0;
            for (int j = 0; j < 3; ++j) {
                const F diff = src.props[j] - tp[j][i];
                if constexpr (Q == 0u || Q == 2u) {
                    tmp[j][i] = diff;
                }
                dist2 = std::fma(diff, diff, dist2);
            }
            if (mac_lh >= dist2) {
                ok = false;
                break;
            }
            dist2 += eps2;
            const F dist = std::sqrt(dist2);
            if constexpr (Q == 0u || Q == 2u) {
                tmp[3][i] = dist * dist2;
            }
            if constexpr (Q == 1u || Q == 2u) {
                tmp[4][i] = dist;
            }
        }
        if (ok) {
            const F msrc = src.props[3];
            for (u64 i = 0; i < tsize; ++i) {
                if constexpr (Q == 0u || Q == 2u) {
                    const F md3 = msrc / tmp[3][i];
                    for (int j = 0; j < 3; ++j) {
                        res[j][i] = std::fma(tmp[j][i], md3, res[j][i]);
                    }
                }
                if constexpr (Q == 1u || Q == 2u) {
                    constexpr int pi = (Q == 1u) ? 0 : 3;
                    res[pi][i] = std::fma(-tp[3][i], msrc / tmp[4][i], res[pi][i]);
                }
            }
            ++cnt.accepted;
            return src_idx + src.n_children + 1u;
        }
        if (!src.n_children) {
            leaf_p2p<Q>(eps2, src, tsize, tp, res);
            ++cnt.leaves_opened;
            cnt.p2p_pairs += tsize * (src.end - src.begin);
        }
        return src_idx + 1u;
    }

    // ---- tree_acc_pot, tree.hpp:2798-2849 ----
    template <unsigned Q>
    void traverse_group(F mac_value, F eps2, u64 tsize, u64 tcode, const F *const tp[4], F *const res[4], F *tmp[5],
                        counters_t &cnt) const
    {
        const unsigned tlevel = node_level(tcode);
        const u64 nn = nodes.size();
        for (u64 s = 0; s < nn;) {
            const node_t<F> &src = nodes[s];
            // NOTE: the reference shifts by (tgt_level - src_level)*NDim unconditionally and relies on
            // the x86 shift-count wrap when src_level > tgt_level (tree.hpp:2828); in that case the
            // shifted code can never equal src.code, which is what the explicit guard states.
            const bool same_branch = src.level <= tlevel && (tcode >> ((tlevel - src.level) * 3u)) == src.code;
            if (same_branch) {
                s += 1u + ((src.code == tcode) ? src.n_children : 0u);
            } else {
                s = mac_check<Q>(s, mac_value, eps2, tsize, tp, res, tmp, cnt);
            }
        }
        self_interactions<Q>(eps2, tsize, tp, res);
        cnt.self_pairs += tsize * (tsize - 1u) / 2u;
    }

    // ---- acc_pot_dispatch (tree.hpp:3293-3334) + acc_pot_impl/cpu_run (tree.hpp:2853-3022) ----
    template <unsigned Q>
    void acc_pot(F theta, F G, F eps, F *const out[4], counters_t &total, u64 *per_group, int nthreads,
                 std::size_t stride = 1, std::size_t offset = 0) const
    {
        if (!std::isfinite(theta) || theta <= F(0)) {
            throw oracle_error(2,
                               "The MAC value must be finite and positive, but it is " + std::to_string(theta)
                                   + " instead");
        }
        const F mac_value = (mac == 0) ? F(1) / (theta * theta) : F(1) / theta;
        if (!std::isfinite(mac_value) || mac_value <= F(0)) {
            throw oracle_error(2, "The transformed MAC value must be finite and positive, but it is "
                                      + std::to_string(mac_value) + " instead");
        }
        const F eps2 = compute_eps2(eps);
        check_G(G);
        constexpr int NR = (Q == 0u) ? 3 : (Q == 1u ? 1 : 4);
        // Groups offset, offset+stride, ... (stride > 1: bounded sample for CPU-baseline timing).
        const std::size_t C = crit.size() > offset ? (crit.size() - offset + stride - 1) / stride : 0;
        std::atomic<std::size_t> next{0};
        const int nt = std::max(1, nthreads);
        std::vector<counters_t> tcnt(nt);
        auto worker = [&](int tid) {
            std::vector<F> tgt[4], resv[4], tmpv[5];
            counters_t &cnt = tcnt[tid];
            for (;;) {
                const std::size_t c0 = next.fetch_add(16);
                if (c0 >= C) {
                    break;
                }
                for (std::size_t cs = c0; cs < std::min(C, c0 + 16); ++cs) {
                    const std::size_t ci = offset + cs * stride;
                    const u64 tb = crit[ci].begin, ts = crit[ci].end - tb;
                    const F *tp[4];
                    F *res[4] = {nullptr, nullptr, nullptr, nullptr}, *tmp[5];
                    for (int j = 0; j < 4; ++j) {
                        tgt[j].assign(parts[j].begin() + tb, parts[j].begin() + tb + ts);
                        tp[j] = tgt[j].data();
                    }
                    for (int j = 0; j < NR; ++j) {
                        resv[j].assign(ts, F(0));
                        res[j] = resv[j].data();
                    }
                    for (int j = 0; j < 5; ++j) {
                        tmpv[j].resize(ts);
                        tmp[j] = tmpv[j].data();
                    }
                    // The scalar kernels index the potential at res[3] when Q == 2 and res[0] when Q == 1.
                    counters_t g;
                    traverse_group<Q>(mac_value, eps2, ts, crit[ci].code, tp, res, tmp, g);
                    g.interactions = g.p2p_pairs + g.accepted * ts + ts * (ts - 1u);
                    if (per_group) {
                        per_group[ci] = g.interactions;
                    }
                    cnt.mac_tests += g.mac_tests;
                    cnt.accepted += g.accepted;
                    cnt.p2p_pairs += g.p2p_pairs;
                    cnt.self_pairs += g.self_pairs;
                    cnt.interactions += g.interactions;
                    cnt.leaves_opened += g.leaves_opened;
                    // G scaling (tree.hpp:2986-3002) and write-out (3004-3007).
                    for (int j = 0; j < NR; ++j) {
                        if (G != F(1)) {
                            for (u64 k = 0; k < ts; ++k) {
                                res[j][k] *= G;
                            }
                        }
                        std::copy(res[j], res[j] + ts, out[j] + tb);
                    }
                }
            }
        };
        if (nt == 1) {
            worker(0);
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t) {
                th.emplace_back(worker, t);
            }
            for (auto &t : th) {
                t.join();
            }
        }
        for (auto &c : tcnt) {
            total.mac_tests += c.mac_tests;
            total.accepted += c.accepted;
            total.p2p_pairs += c.p2p_pairs;
            total.self_pairs += c.self_pairs;
            total.interactions += c.interactions;
            total.leaves_opened += c.leaves_opened;
        }
    }

    // compute_eps2 / check_G_const, tree.hpp:3268-3289.
    static F compute_eps2(F eps)
    {
        if (!std::isfinite(eps) || eps < F(0)) {
            throw oracle_error(2, "The softening length must be finite and non-negative, but it is "
                                      + std::to_string(eps) + " instead");
        }
        const F e2 = eps * eps;
        if (!std::isfinite(e2) || e2 < F(0)) {
            throw oracle_error(2, "The square of the softening length must be finite and non-negative, but it is "
                                      + std::to_string(e2) + " instead");
        }
        return e2;
    }
    static void check_G(F G)
    {
        if (!std::isfinite(G)) {
            throw oracle_error(2, "The value of the gravitational constant G must be finite, but it is "
                                      + std::to_string(G) + " instead");
        }
    }

    // ---- exact_acc_pot_impl, tree.hpp:3531-3569 (idx in Morton order) ----
    void exact(u64 idx, F G, F eps, F out[4]) const
    {
        const F eps2 = compute_eps2(eps);
        check_G(G);
        F acc[3] = {0, 0, 0}, pot = 0;
        const std::size_t N = n();
        for (std::size_t i = 0; i < N; ++i) {
            if (i == idx) {
                continue;
            }
            F d[3], dist2 = eps2;
            for (int j = 0; j < 3; ++j) {
                d[j] = parts[j][i] - parts[j][idx];
                dist2 = std::fma(d[j], d[j], dist2);
            }
            const F inv = F(1) / std::sqrt(dist2), gmi = G * parts[3][i] * inv;
            const F gmi3 = inv * inv * gmi;
            for (int j = 0; j < 3; ++j) {
                acc[j] = std::fma(d[j], gmi3, acc[j]);
            }
            pot = std::fma(-gmi, parts[3][idx], pot);
        }
        out[0] = acc[0];
        out[1] = acc[1];
        out[2] = acc[2];
        out[3] = pot;
    }

    // ---- sync(), tree.hpp:3678-3743: positions were overwritten in Morton order ----
    void sync_positions()
    {
        try {
            encode_sort_permute(false);
            build_nodes();
        } catch (...) {
            clear();
            throw;
        }
    }
    // ---- update_masses_dispatch, tree.hpp:3782-3805 ----
    void sync_masses()
    {
        try {
            for (auto &nd : nodes) {
                node_props(nd);
            }
        } catch (...) {
            clear();
            throw;
        }
    }
};

struct handle_t {
    int fp; // 32 or 64
    tree_t<float> t32;
    tree_t<double> t64;
    std::string err;
};

template <typename Fn>
int guarded(handle_t *h, Fn &&fn)
{
    try {
        fn();
        return 0;
    } catch (const oracle_error &e) {
        h->err = e.what();
        return e.code;
    } catch (const std::exception &e) {
        h->err = e.what();
        return 99;
    }
}

// Plummer sphere, sequential branch of benchmark/common.hpp:96-126 (std::mt19937 default seed,
// libstdc++ distributions), layout [m | x | y | z].
template <typename F>
void plummer_seq(std::size_t n, F a, F size, std::mt19937 &rng, F *out)
{
    const F pi = static_cast<F>(3.141592653589793238462643383279502884L);
    const F lim = (size > F(0)) ? (size / F(2) - size / F(100)) : std::numeric_limits<F>::infinity();
    std::uniform_real_distribution<F> udist(F(0), F(1));
    std::uniform_real_distribution<F> mdist(F(0.1), F(1.9));
    for (std::size_t i = 0; i < n; ++i) {
        out[i] = mdist(rng);
    }
    for (std::size_t i = 0; i < n;) {
        F r;
        do {
            r = a / std::sqrt(std::pow(udist(rng), F(-2) / F(3)) - F(1));
        } while (!std::isfinite(r));
        const F u = udist(rng), v = udist(rng);
        const F lon = std::clamp(F(2) * pi * u, F(0), F(2) * pi);
        const F colat = std::acos(std::clamp(F(2) * v - F(1), F(-1), F(1)));
        const F x = r * std::cos(lon) * std::sin(colat), y = r * std::sin(lon) * std::sin(colat),
                z = r * std::cos(colat);
        if (x >= -lim && x < lim && y >= -lim && y < lim && z >= -lim && z < lim) {
            out[n + i] = x;
            out[2 * n + i] = y;
            out[3 * n + i] = z;
            ++i;
        }
    }
}

// Chunked deterministic variant of the parallel branch (benchmark/common.hpp:60-95): fixed chunks,
// each chunk's rng seeded with its first index. Masses are interleaved with
// the position draws, exactly as in the reference's parallel branch.
template <typename F>
void plummer_chunked(std::size_t n, F a, F size, std::size_t chunk, int nthreads, F *out)
{
    const F pi = static_cast<F>(3.141592653589793238462643383279502884L);
    const F lim = (size > F(0)) ? (size / F(2) - size / F(100)) : std::numeric_limits<F>::infinity();
    const std::size_t nchunks = (n + chunk - 1) / chunk;
    std::atomic<std::size_t> next{0};
    auto worker = [&]() {
        for (;;) {
            const std::size_t c = next.fetch_add(1);
            if (c >= nchunks) {
                break;
            }
            const std::size_t b = c * chunk, e = std::min(n, b + chunk);
            std::mt19937 rng;
            rng.seed(static_cast<std::mt19937::result_type>(b));
            std::uniform_real_distribution<F> udist(F(0), F(1));
            std::uniform_real_distribution<F> mdist(F(0.1), F(1.9));
            for (std::size_t i = b; i < e;) {
                out[i] = mdist(rng);
                F r;
                do {
                    r = a / std::sqrt(std::pow(udist(rng), F(-2) / F(3)) - F(1));
                } while (!std::isfinite(r));
                const F u = udist(rng), v = udist(rng);
                const F lon = std::clamp(F(2) * pi * u, F(0), F(2) * pi);
                const F colat = std::acos(std::clamp(F(2) * v - F(1), F(-1), F(1)));
                const F x = r * std::cos(lon) * std::sin(colat), y = r * std::sin(lon) * std::sin(colat),
                        z = r * std::cos(colat);
                if (x >= -lim && x < lim && y >= -lim && y < lim && z >= -lim && z < lim) {
                    out[n + i] = x;
                    out[2 * n + i] = y;
                    out[3 * n + i] = z;
                    ++i;
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < std::max(1, nthreads); ++t) {
        th.emplace_back(worker);
    }
    for (auto &t : th) {
        t.join();
    }
}

// get_uniform_particles, test/test_utils.hpp:41-59: masses U[0,1) then coords U[-size/2,size/2),
// layout [m | x | y | z], engine state carried by the caller through `seed`/`discard`.
template <typename F>
void uniform_fixture(std::size_t n, F size, std::mt19937 &rng, F *out)
{
    std::uniform_real_distribution<F> mdist(F(0), F(1));
    for (std::size_t i = 0; i < n; ++i) {
        out[i] = mdist(rng);
    }
    std::uniform_real_distribution<F> rdist(-size / F(2), size / F(2));
    for (std::size_t i = n; i < 4 * n; ++i) {
        out[i] = rdist(rng);
    }
}

} // namespace

#define ORC_API extern "C" __attribute__((visibility("default")))

ORC_API void *orc_create(int fp, int mac)
{
    if ((fp != 32 && fp != 64) || (mac != 0 && mac != 1)) {
        return nullptr;
    }
    auto *h = new handle_t{};
    h->fp = fp;
    h->t32.mac = mac;
    h->t64.mac = mac;
    return h;
}
ORC_API void orc_destroy(void *p)
{
    delete static_cast<handle_t *>(p);
}
ORC_API const char *orc_last_error(void *p)
{
    return static_cast<handle_t *>(p)->err.c_str();
}

#define DISPATCH(h, expr32, expr64)                                                                                    \
    do {                                                                                                               \
        if ((h)->fp == 32) {                                                                                           \
            auto &t = (h)->t32;                                                                                        \
            using F = float;                                                                                           \
            (void)sizeof(F);                                                                                           \
            expr32;                                                                                                    \
        } else {                                                                                                       \
            auto &t = (h)->t64;                                                                                        \
            using F = double;                                                                                          \
            (void)sizeof(F);                                                                                           \
            expr64;                                                                                                    \
        }                                                                                                              \
    } while (0)

ORC_API int orc_build(void *p, const void *x, const void *y, const void *z, const void *m, std::size_t n,
                      double box_size, int deduce, std::size_t max_leaf_n, std::size_t ncrit)
{
    auto *h = static_cast<handle_t *>(p);
    return guarded(h, [&]() {
        DISPATCH(h,
                 t.construct(static_cast<const F *>(x), static_cast<const F *>(y), static_cast<const F *>(z),
                             static_cast<const F *>(m), n, static_cast<F>(box_size), deduce != 0, max_leaf_n, ncrit),
                 t.construct(static_cast<const F *>(x), static_cast<const F *>(y), static_cast<const F *>(z),
                             static_cast<const F *>(m), n, static_cast<F>(box_size), deduce != 0, max_leaf_n, ncrit));
    });
}

ORC_API std::size_t orc_nparts(void *p)
{
    auto *h = static_cast<handle_t *>(p);
    return h->fp == 32 ? h->t32.n() : h->t64.n();
}
ORC_API std::size_t orc_nnodes(void *p)
{
    auto *h = static_cast<handle_t *>(p);
    return h->fp == 32 ? h->t32.nodes.size() : h->t64.nodes.size();
}
ORC_API std::size_t orc_ncrit_nodes(void *p)
{
    auto *h = static_cast<handle_t *>(p);
    return h->fp == 32 ? h->t32.crit.size() : h->t64.crit.size();
}
ORC_API double orc_box_size(void *p)
{
    auto *h = static_cast<handle_t *>(p);
    return h->fp == 32 ? static_cast<double>(h->t32.box_size) : h->t64.box_size;
}
ORC_API void orc_get_codes(void *p, u64 *out)
{
    auto *h = static_cast<handle_t *>(p);
    DISPATCH(h, std::copy(t.codes.begin(), t.codes.end(), out), std::copy(t.codes.begin(), t.codes.end(), out));
}
// which: 0 perm, 1 last_perm, 2 inv_perm
ORC_API void orc_get_perm(void *p, int which, u64 *out)
{
    auto *h = static_cast<handle_t *>(p);
    auto pick = [which](auto &t) -> const std::vector<u64> & {
        return which == 0 ? t.perm : (which == 1 ? t.last_perm : t.inv_perm);
    };
    DISPATCH(h, { auto &v = pick(t); std::copy(v.begin(), v.end(), out); },
             { auto &v = pick(t); std::copy(v.begin(), v.end(), out); });
}
ORC_API void orc_get_parts(void *p, void *x, void *y, void *z, void *m)
{
    auto *h = static_cast<handle_t *>(p);
    void *o[4] = {x, y, z, m};
    DISPATCH(h,
             for (int j = 0; j < 4; ++j) std::copy(t.parts[j].begin(), t.parts[j].end(), static_cast<F *>(o[j])),
             for (int j = 0; j < 4; ++j) std::copy(t.parts[j].begin(), t.parts[j].end(), static_cast<F *>(o[j])));
}
// Node AoS: {u64 begin,end,n_children,code,level; F props[4]; F dim; F delta} (64 B f32, 88 B f64).
ORC_API std::size_t orc_node_stride(void *p)
{
    auto *h = static_cast<handle_t *>(p);
    return h->fp == 32 ? sizeof(node_t<float>) : sizeof(node_t<double>);
}
ORC_API void orc_get_nodes(void *p, void *out)
{
    auto *h = static_cast<handle_t *>(p);
    DISPATCH(h, std::memcpy(out, t.nodes.data(), t.nodes.size() * sizeof(node_t<F>)),
             std::memcpy(out, t.nodes.data(), t.nodes.size() * sizeof(node_t<F>)));
}
// Critical nodes as (code, begin, end) u64 triplets; node_idx (nullable) gets the node-array index.
ORC_API void orc_get_crit(void *p, u64 *out, u64 *node_idx)
{
    auto *h = static_cast<handle_t *>(p);
    auto cp = [&](auto &t) {
        for (std::size_t i = 0; i < t.crit.size(); ++i) {
            out[3 * i] = t.crit[i].code;
            out[3 * i + 1] = t.crit[i].begin;
            out[3 * i + 2] = t.crit[i].end;
            if (node_idx) {
                node_idx[i] = t.crit_node_idx[i];
            }
        }
    };
    DISPATCH(h, cp(t), cp(t));
}

// out0..3: Morton-order outputs. Q=0: ax,ay,az ; Q=1: pot in out0 ; Q=2: ax,ay,az,pot.
// counters[6]: mac_tests, accepted, p2p_pairs, self_pairs, interactions, leaves_opened.
static int orc_acc_pot_impl(void *p, int Q, double theta, double G, double eps, void *o0, void *o1, void *o2, void *o3,
                            u64 *counters, u64 *per_group, int nthreads, std::size_t stride, std::size_t offset)
{
    auto *h = static_cast<handle_t *>(p);
    return guarded(h, [&]() {
        counters_t c;
        auto run = [&](auto &t, auto fzero) {
            using F = decltype(fzero);
            F *out[4] = {static_cast<F *>(o0), static_cast<F *>(o1), static_cast<F *>(o2), static_cast<F *>(o3)};
            if (Q == 0) {
                t.template acc_pot<0>(F(theta), F(G), F(eps), out, c, per_group, nthreads, stride, offset);
            } else if (Q == 1) {
                t.template acc_pot<1>(F(theta), F(G), F(eps), out, c, per_group, nthreads, stride, offset);
            } else {
                t.template acc_pot<2>(F(theta), F(G), F(eps), out, c, per_group, nthreads, stride, offset);
            }
        };
        DISPATCH(h, run(t, 0.f), run(t, 0.));
        if (counters) {
            counters[0] = c.mac_tests;
            counters[1] = c.accepted;
            counters[2] = c.p2p_pairs;
            counters[3] = c.self_pairs;
            counters[4] = c.interactions;
            counters[5] = c.leaves_opened;
        }
    });
}

ORC_API int orc_acc_pot(void *p, int Q, double theta, double G, double eps, void *o0, void *o1, void *o2, void *o3,
                        u64 *counters, u64 *per_group, int nthreads)
{
    return orc_acc_pot_impl(p, Q, theta, G, eps, o0, o1, o2, o3, counters, per_group, nthreads, 1, 0);
}
// Bounded sample: only critical nodes offset, offset+stride, ... are evaluated (outputs of the others untouched).
ORC_API int orc_acc_pot_sample(void *p, int Q, double theta, double G, double eps, void *o0, void *o1, void *o2,
                               void *o3, u64 *counters, int nthreads, std::size_t stride, std::size_t offset)
{
    return orc_acc_pot_impl(p, Q, theta, G, eps, o0, o1, o2, o3, counters, nullptr, nthreads, stride ? stride : 1,
                            offset);
}

// Direct sum for the particle at Morton index idx; out4 = ax, ay, az, pot (as double).
ORC_API int orc_exact(void *p, std::size_t idx, double G, double eps, double *out4)
{
    auto *h = static_cast<handle_t *>(p);
    return guarded(h, [&]() {
        auto run = [&](auto &t, auto fzero) {
            using F = decltype(fzero);
            F o[4];
            t.exact(idx, F(G), F(eps), o);
            for (int j = 0; j < 4; ++j) {
                out4[j] = o[j];
            }
        };
        DISPATCH(h, run(t, 0.f), run(t, 0.));
    });
}

// New positions (Morton order) -> sync. Any of x,y,z may be NULL (unchanged).
ORC_API int orc_update_positions(void *p, const void *x, const void *y, const void *z)
{
    auto *h = static_cast<handle_t *>(p);
    return guarded(h, [&]() {
        const void *in[3] = {x, y, z};
        auto run = [&](auto &t, auto fzero) {
            using F = decltype(fzero);
            for (int j = 0; j < 3; ++j) {
                if (in[j]) {
                    std::copy(static_cast<const F *>(in[j]), static_cast<const F *>(in[j]) + t.n(),
                              t.parts[j].begin());
                }
            }
            t.sync_positions();
        };
        DISPATCH(h, run(t, 0.f), run(t, 0.));
    });
}
ORC_API int orc_update_masses(void *p, const void *m)
{
    auto *h = static_cast<handle_t *>(p);
    return guarded(h, [&]() {
        auto run = [&](auto &t, auto fzero) {
            using F = decltype(fzero);
            std::copy(static_cast<const F *>(m), static_cast<const F *>(m) + t.n(), t.parts[3].begin());
            t.sync_masses();
        };
        DISPATCH(h, run(t, 0.f), run(t, 0.));
    });
}

// Node centre (get_node_centre) for tests.
ORC_API void orc_node_centre(void *p, u64 code, double *out3)
{
    auto *h = static_cast<handle_t *>(p);
    auto run = [&](auto &t, auto fzero) {
        using F = decltype(fzero);
        F o[3];
        t.node_centre(o, code);
        for (int j = 0; j < 3; ++j) {
            out3[j] = o[j];
        }
    };
    DISPATCH(h, run(t, 0.f), run(t, 0.));
}

ORC_API u64 orc_morton_encode(u64 x, u64 y, u64 z)
{
    return morton3(x, y, z);
}
ORC_API void orc_morton_decode(u64 code, u64 *xyz)
{
    xyz[0] = compact3(code);
    xyz[1] = compact3(code >> 1);
    xyz[2] = compact3(code >> 2);
}

// Generators. out has 4*n elements of the requested precision, layout [m | x | y | z].
ORC_API void orc_plummer(int fp, std::size_t n, double a, double size, void *out)
{
    std::mt19937 rng; // default seed 5489, as the thread_local rng of benchmark/common.hpp:36
    if (fp == 32) {
        plummer_seq<float>(n, float(a), float(size), rng, static_cast<float *>(out));
    } else {
        plummer_seq<double>(n, a, size, rng, static_cast<double *>(out));
    }
}
ORC_API void orc_plummer_chunked(int fp, std::size_t n, double a, double size, std::size_t chunk, int nthreads,
                                 void *out)
{
    if (fp == 32) {
        plummer_chunked<float>(n, float(a), float(size), chunk, nthreads, static_cast<float *>(out));
    } else {
        plummer_chunked<double>(n, a, size, chunk, nthreads, static_cast<double *>(out));
    }
}
// Persistent engine so that consecutive fixture draws reproduce a test file's rng stream.
ORC_API void *orc_rng_create(std::uint32_t seed)
{
    return new std::mt19937(seed);
}
ORC_API void orc_rng_destroy(void *r)
{
    delete static_cast<std::mt19937 *>(r);
}
ORC_API void orc_uniform(int fp, std::size_t n, double size, void *rng, void *out)
{
    auto &r = *static_cast<std::mt19937 *>(rng);
    if (fp == 32) {
        uniform_fixture<float>(n, float(size), r, static_cast<float *>(out));
    } else {
        uniform_fixture<double>(n, size, r, static_cast<double *>(out));
    }
}
