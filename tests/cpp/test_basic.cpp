// Mirrors the reference's test/basic.cpp (constructors, keyword arguments, exception messages, copy/move,
// code iterators), test/morton.cpp, test/node_centre.cpp and test/auto_box_size.cpp against
// include/rakau/tree.hpp of this repository (GPU-backed).
#include <algorithm>
#include <array>
#include <cmath>
#include <initializer_list>
#include <iterator>
#include <limits>
#include <numeric>
#include <random>
#include <sstream>
#include <tuple>
#include <type_traits>
#include <vector>

#include <rakau/tree.hpp>

#include "mini_test.hpp"
#include "test_utils.hpp"

using namespace rakau;
using namespace rakau::kwargs;
using namespace rakau_test;
using mini_test::tuple_for_each;

using fp_types = std::tuple<float, double>;
using macs = std::tuple<std::integral_constant<mac, mac::bh>, std::integral_constant<mac, mac::bh_geom>>;

static std::mt19937 rng;

TEST_CASE("ctors")
{
    tuple_for_each(macs{}, [](auto mac_type) {
        tuple_for_each(fp_types{}, [](auto x) {
            using fp_type = decltype(x);
            constexpr fp_type bsize = 10;
            constexpr unsigned N = 100;
            using tree_t = octree<fp_type, decltype(mac_type)::value>;
            tree_t t0;
            REQUIRE(t0.box_size() == fp_type(0));
            REQUIRE(!t0.box_size_deduced());
            REQUIRE(t0.ncrit() == default_ncrit);
            REQUIRE(t0.max_leaf_n() == default_max_leaf_n);
            REQUIRE(t0.perm().empty());
            REQUIRE(t0.last_perm().empty());
            REQUIRE(t0.inv_perm().empty());
            REQUIRE(t0.nparts() == 0u);
            REQUIRE(t0.nodes().empty());
            auto parts = get_uniform_particles<3>(N, bsize, rng);
            // iterators + nparts, default and non-default parameters, any keyword order
            tree_t t1{x_coords = parts.begin() + N, y_coords = parts.begin() + 2u * N, z_coords = parts.begin() + 3u * N,
                      masses = parts.begin(),       nparts = N,                        box_size = bsize};
            REQUIRE(t1.box_size() == bsize);
            REQUIRE(!t1.box_size_deduced());
            REQUIRE(t1.max_leaf_n() == default_max_leaf_n);
            REQUIRE(t1.ncrit() == default_ncrit);
            REQUIRE(t1.perm() == t1.last_perm());
            REQUIRE(t1.inv_perm().size() == N);
            REQUIRE(t1.nparts() == N);
            tree_t t2{masses = parts.begin(), nparts = N, max_leaf_n = 4, ncrit = 5, box_size = bsize,
                      z_coords = parts.begin() + 3u * N, x_coords = parts.begin() + N, y_coords = parts.begin() + 2u * N};
            REQUIRE(t2.box_size() == bsize);
            REQUIRE(t2.max_leaf_n() == 4u);
            REQUIRE(t2.ncrit() == 5u);
            REQUIRE(t2.perm() == t2.last_perm());
            // the ordered view returns the input
            REQUIRE(std::equal(parts.begin() + N, parts.begin() + 2u * N, t1.p_its_o()[0]));
            REQUIRE(std::equal(parts.begin(), parts.begin() + N, t1.p_its_o()[3]));
            // ranges (vectors)
            std::array<std::vector<fp_type>, 4> arr_vec;
            for (auto &vec : arr_vec) {
                vec.resize(N);
                std::uniform_real_distribution<fp_type> urd(-fp_type(1), fp_type(1));
                std::generate(vec.begin(), vec.end(), [&urd]() { return urd(rng); });
            }
            tree_t tvec1{x_coords = arr_vec[0], y_coords = arr_vec[1], z_coords = arr_vec[2], masses = arr_vec[3],
                         box_size = 100};
            REQUIRE(tvec1.nparts() == N);
            REQUIRE(tvec1.box_size() == fp_type(100));
            for (std::size_t j = 0; j < 4; ++j) {
                REQUIRE(std::equal(arr_vec[j].begin(), arr_vec[j].end(), tvec1.p_its_o()[j]));
            }
            REQUIRE_THROWS_WITH(
                (tree_t{x_coords = arr_vec[0], y_coords = arr_vec[1], z_coords = arr_vec[2],
                        masses = std::vector<fp_type>{}, box_size = 3, max_leaf_n = 4, ncrit = 5}),
                "The size of the input range for the particle masses (0) is different from the size of "
                "the input ranges for the particle coordinates ("
                    + std::to_string(arr_vec[0].size()) + ")");
            auto short_z = arr_vec[2];
            short_z.clear();
            REQUIRE_THROWS_WITH((tree_t{x_coords = arr_vec[0], y_coords = arr_vec[1], z_coords = short_z,
                                        masses = arr_vec[3], box_size = 3, max_leaf_n = 4, ncrit = 5}),
                                "The input ranges for the particle coordinates have inconsistent sizes");
            // deduced box size (test/basic.cpp:144-147)
            fp_type xcoords[] = {-10, 1, 2, 10}, ycoords[] = {-10, 1, 2, 10}, zcoords[] = {-10, 1, 2, 10},
                    pmasses[] = {1, 1, 1, 1};
            tree_t t3{x_coords = xcoords, y_coords = ycoords, z_coords = zcoords, masses = pmasses};
            REQUIRE(t3.box_size() == fp_type(21));
            REQUIRE(t3.box_size_deduced());
            REQUIRE(t3.max_leaf_n() == default_max_leaf_n);
            REQUIRE(t3.ncrit() == default_ncrit);
            REQUIRE(t3.perm() == t3.last_perm());
            REQUIRE(t3.inv_perm().size() == 4u);
            tree_t t4{x_coords = xcoords, y_coords = ycoords, z_coords = zcoords,
                      masses = pmasses,   max_leaf_n = 4,     ncrit = 5};
            REQUIRE(t4.box_size() == fp_type(21));
            REQUIRE(t4.box_size_deduced());
            REQUIRE(t4.max_leaf_n() == 4u);
            REQUIRE(t4.ncrit() == 5u);
            // error messages (test/basic.cpp:178-202)
            REQUIRE_THROWS_WITH((tree_t{x_coords = xcoords, y_coords = ycoords, z_coords = zcoords, masses = pmasses,
                                        box_size = 0., max_leaf_n = 4, ncrit = 5}),
                                "While trying to discretise the input coordinate");
            REQUIRE_THROWS_WITH((tree_t{x_coords = xcoords, y_coords = ycoords, z_coords = zcoords, masses = pmasses,
                                        box_size = 3, max_leaf_n = 4, ncrit = 5}),
                                "produced the floating-point value");
            REQUIRE_THROWS_WITH((tree_t{x_coords = xcoords, y_coords = ycoords, z_coords = zcoords, masses = pmasses,
                                        box_size = -3, max_leaf_n = 4, ncrit = 5}),
                                "The box size must be a finite non-negative value, but it is");
            REQUIRE_THROWS_WITH(
                (tree_t{x_coords = xcoords, y_coords = ycoords, z_coords = zcoords, masses = pmasses,
                        box_size = std::numeric_limits<fp_type>::infinity(), max_leaf_n = 4, ncrit = 5}),
                "The box size must be a finite non-negative value, but it is");
            REQUIRE_THROWS_AS((tree_t{x_coords = xcoords, y_coords = ycoords, z_coords = zcoords, masses = pmasses,
                                      box_size = -3}),
                              std::invalid_argument);
            REQUIRE_THROWS_WITH((tree_t{x_coords = xcoords, y_coords = ycoords, z_coords = zcoords, masses = pmasses,
                                        max_leaf_n = 0, ncrit = 5}),
                                "The maximum number of particles per leaf must be nonzero");
            REQUIRE_THROWS_WITH((tree_t{x_coords = xcoords, y_coords = ycoords, z_coords = zcoords, masses = pmasses,
                                        max_leaf_n = 4, ncrit = 0}),
                                "The critical number of particles for the vectorised computation of the");
            // copy / move semantics (test/basic.cpp:203-260)
            tree_t t4_copy(t4);
            REQUIRE(t4_copy.box_size() == fp_type(21));
            REQUIRE(t4_copy.box_size_deduced());
            REQUIRE(t4_copy.max_leaf_n() == 4u);
            REQUIRE(t4_copy.ncrit() == 5u);
            REQUIRE(t4_copy.perm() == t4.perm());
            REQUIRE(t4_copy.last_perm() == t4.last_perm());
            REQUIRE(t4_copy.inv_perm() == t4.inv_perm());
            REQUIRE(t4_copy.nodes() == t4.nodes());
            REQUIRE(std::equal(t4.p_its_u()[0], t4.p_its_u()[0] + 4, t4_copy.p_its_u()[0]));
            tree_t t4_move(std::move(t4_copy));
            REQUIRE(t4_move.box_size() == fp_type(21));
            REQUIRE(t4_move.nodes() == t4.nodes());
            // the moved-from tree is in the default state
            REQUIRE(t4_copy.box_size() == fp_type(0));
            REQUIRE(!t4_copy.box_size_deduced());
            REQUIRE(t4_copy.max_leaf_n() == default_max_leaf_n);
            REQUIRE(t4_copy.ncrit() == default_ncrit);
            REQUIRE(t4_copy.perm().empty());
            REQUIRE(t4_copy.nparts() == 0u);
            tree_t t5;
            t5 = t4;
            REQUIRE(t5.nodes() == t4.nodes());
            REQUIRE(t5.ncrit() == 5u);
            t5 = std::move(t4_move);
            REQUIRE(t5.nodes() == t4.nodes());
            REQUIRE(t4_move.nparts() == 0u);
            t5 = *&t5; // self assignment
            REQUIRE(t5.nodes() == t4.nodes());
            std::ostringstream oss;
            oss << t5;
            REQUIRE(oss.str().find("Total number of particles: 4") != std::string::npos);
        });
    });
}

TEST_CASE("code iterators")
{
    // test/basic.cpp:322-354
    tuple_for_each(macs{}, [](auto mac_type) {
        tuple_for_each(fp_types{}, [](auto x) {
            using fp_type = decltype(x);
            constexpr auto bsize = static_cast<fp_type>(1);
            constexpr auto s = 10000u;
            auto parts = get_uniform_particles<3>(s, bsize, rng);
            octree<fp_type, decltype(mac_type)::value> t{x_coords = parts.begin() + s,
                                                         y_coords = parts.begin() + 2u * s,
                                                         z_coords = parts.begin() + 3u * s,
                                                         masses = parts.begin(),
                                                         nparts = s,
                                                         box_size = fp_type(1.25)};
            REQUIRE(std::is_sorted(t.c_it_u(), t.c_it_u() + s));
            auto cit = t.c_it_o();
            const fp_type inv_box = fp_type(1) / fp_type(1.25);
            morton_encoder<3, std::size_t> me;
            for (auto i = 0u; i < s; ++i) {
                std::size_t d[3];
                for (std::size_t j = 0; j < 3; ++j) {
                    auto tmp = fma_wrap(parts[(j + 1u) * s + i], inv_box, fp_type(1) / fp_type(2));
                    tmp *= fp_type(std::size_t(1) << 21);
                    d[j] = static_cast<std::size_t>(tmp);
                }
                REQUIRE(cit[i] == me(&d[0]));
            }
        });
    });
}

TEST_CASE("morton")
{
    // test/morton.cpp:35-58
    morton_encoder<3, std::uint64_t> me;
    morton_decoder<3, std::uint64_t> md;
    std::uniform_int_distribution<std::uint64_t> dist(0, (std::uint64_t(1) << 21) - 1u);
    for (int i = 0; i < 10000; ++i) {
        std::uint64_t in[3] = {dist(rng), dist(rng), dist(rng)}, out[3];
        md(&out[0], me(&in[0]));
        REQUIRE(std::equal(in, in + 3, out));
    }
}

TEST_CASE("node centre")
{
    // test/node_centre.cpp:55-110
    tuple_for_each(macs{}, [](auto mac_type) {
        tuple_for_each(fp_types{}, [](auto x) {
            using fp_type = decltype(x);
            const fp_type eps = std::numeric_limits<fp_type>::epsilon() * 10;
            fp_type xc[] = {1, 1, 1, 1, -1, -1, -1, -1}, yc[] = {1, 1, -1, -1, 1, 1, -1, -1},
                    zc[] = {1, -1, 1, -1, 1, -1, 1, -1}, ms[] = {1, 1, 1, 1, 1, 1, 1, 1};
            octree<fp_type, decltype(mac_type)::value> t{x_coords = xc, y_coords = yc, z_coords = zc, masses = ms,
                                                         box_size = 10, max_leaf_n = 1, ncrit = 1};
            const auto &nodes = t.nodes();
            REQUIRE(nodes.size() == 9u);
            fp_type c[3];
            get_node_centre(c, nodes[0].code, fp_type(10));
            REQUIRE((std::abs(c[0]) <= eps && std::abs(c[1]) <= eps && std::abs(c[2]) <= eps));
            const fp_type q = fp_type(10) / 4;
            const fp_type exp[8][3] = {{-q, -q, -q}, {q, -q, -q}, {-q, q, -q}, {q, q, -q},
                                       {-q, -q, q},  {q, -q, q},  {-q, q, q},  {q, q, q}};
            for (std::size_t i = 0; i < 8; ++i) {
                get_node_centre(c, nodes[i + 1u].code, fp_type(10));
                for (int j = 0; j < 3; ++j) {
                    REQUIRE(std::abs(c[j] - exp[i][j]) <= eps * q);
                    REQUIRE(nodes[i + 1u].props[j] == (exp[i][j] > 0 ? fp_type(1) : fp_type(-1)));
                }
                REQUIRE(nodes[i + 1u].level == 1u);
                REQUIRE(nodes[i + 1u].n_children == 0u);
            }
            REQUIRE(nodes[0].n_children == 8u);
            REQUIRE(nodes[0].props[3] == fp_type(8));
        });
    });
}

TEST_CASE("automatic box size")
{
    // test/auto_box_size.cpp:28-76
    tuple_for_each(macs{}, [](auto mac_type) {
        tuple_for_each(fp_types{}, [](auto x) {
            using fp_type = decltype(x);
            fp_type x_c[] = {0, 1, 2, 3}, y_c[] = {-4, -5, -6, -7}, z_c[] = {4, 5, 3, 1}, p_masses[] = {1, 1, 1, 1};
            octree<fp_type, decltype(mac_type)::value> t{x_coords = x_c,    y_coords = y_c, z_coords = z_c,
                                                         masses = p_masses, max_leaf_n = 1, ncrit = 1};
            REQUIRE(t.box_size_deduced());
            REQUIRE(t.box_size() == 14 + fp_type(0.7));
            t.update_particles_u([](const auto &its) {
                for (std::size_t i = 0; i < 4u; ++i) {
                    for (std::size_t j = 0; j < 3u; ++j) {
                        its[j][i] *= 2;
                    }
                }
            });
            REQUIRE(t.box_size_deduced());
            REQUIRE(t.box_size() == 28 + fp_type(1.4));
            t.update_particles_u([](const auto &its) {
                for (std::size_t i = 0; i < 4u; ++i) {
                    for (std::size_t j = 0; j < 3u; ++j) {
                        its[j][i] /= 4;
                    }
                }
            });
            REQUIRE(t.box_size() == 7 + fp_type(0.35));
            auto its = t.p_its_o();
            for (int i = 0; i < 4; ++i) {
                REQUIRE(its[0][i] == fp_type(i) / 2);
                REQUIRE(its[1][i] == fp_type(-4 - i) / 2);
            }
            REQUIRE(its[2][0] == fp_type(4) / 2);
            REQUIRE(its[2][1] == fp_type(5) / 2);
            REQUIRE(its[2][2] == fp_type(3) / 2);
            REQUIRE(its[2][3] == fp_type(1) / 2);
        });
    });
}

MINI_TEST_MAIN()
