// Mirrors the reference's test/update.cpp and test/update_masses.cpp.
#include <algorithm>
#include <array>
#include <cmath>
#include <iterator>
#include <limits>
#include <numeric>
#include <random>
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

TEST_CASE("update positions")
{
    // test/update.cpp:34-200
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
                                                         box_size = fp_type(10)},
                t2(t);
            REQUIRE(t.perm() == t.last_perm());
            using size_type = typename decltype(t)::size_type;
            std::vector<size_type> track_idx(1000);
            std::uniform_int_distribution<size_type> idist(0, s - 1u);
            std::generate(track_idx.begin(), track_idx.end(), [&idist]() { return idist(rng); });
            auto check_ordered = [&](std::size_t ix, std::size_t iy, std::size_t iz, fp_type add, fp_type div) {
                auto pro = t.p_its_o();
                for (auto idx : track_idx) {
                    const auto d = static_cast<std::ptrdiff_t>(idx);
                    REQUIRE(pro[0][d] == (parts[ix * s + idx] + add) / div);
                    REQUIRE(pro[1][d] == (parts[iy * s + idx] + add) / div);
                    REQUIRE(pro[2][d] == (parts[iz * s + idx] + add) / div);
                    REQUIRE(pro[3][d] == parts[idx]);
                }
            };
            check_ordered(1, 2, 3, 0, 1);
            auto orig_perm = t.perm();
            auto orig_last_perm = t.last_perm();
            auto orig_inv_perm = t.inv_perm();
            t.update_particles_u([](const auto &) {});
            REQUIRE(orig_perm == t.perm());
            std::iota(orig_last_perm.begin(), orig_last_perm.end(), size_type(0));
            REQUIRE(orig_last_perm == t.last_perm());
            REQUIRE(orig_inv_perm == t.inv_perm());
            check_ordered(1, 2, 3, 0, 1);
            t.update_particles_o([](const auto &) {});
            REQUIRE(orig_perm == t.perm());
            REQUIRE(orig_last_perm == t.last_perm());
            REQUIRE(orig_inv_perm == t.inv_perm());
            check_ordered(1, 2, 3, 0, 1);
            {
                auto pru = t.p_its_u();
                auto pru2 = t2.p_its_u();
                for (auto idx : track_idx) {
                    for (int j = 0; j < 4; ++j) {
                        REQUIRE(pru[j][idx] == pru2[j][idx]);
                    }
                }
            }
            REQUIRE(t.nodes() == t2.nodes());
            // x, y, z -> y, z, x through the ordered iterators
            std::vector<fp_type> x_morton_old(t.p_its_u()[0], t.p_its_u()[0] + s), x_morton_orig(x_morton_old);
            t.update_particles_o([](const auto &p_its) {
                auto x_it = p_its[0], y_it = p_its[1], z_it = p_its[2];
                for (std::ptrdiff_t idx = 0; idx < std::ptrdiff_t(s); ++idx) {
                    std::swap(*(x_it + idx), *(y_it + idx));
                    std::swap(*(y_it + idx), *(z_it + idx));
                }
            });
            check_ordered(2, 3, 1, 0, 1);
            auto lp = t.last_perm();
            auto x_morton_new(x_morton_old);
            for (std::size_t i = 0; i < lp.size(); ++i) {
                x_morton_new[i] = x_morton_old[lp[i]];
            }
            REQUIRE(std::equal(x_morton_new.begin(), x_morton_new.end(), t.p_its_u()[2]));
            t.update_particles_o([](const auto &p_its) {
                auto x_it = p_its[0], y_it = p_its[1], z_it = p_its[2];
                for (std::ptrdiff_t idx = 0; idx < std::ptrdiff_t(s); ++idx) {
                    std::swap(*(z_it + idx), *(y_it + idx));
                    std::swap(*(x_it + idx), *(y_it + idx));
                }
            });
            check_ordered(1, 2, 3, 0, 1);
            lp = t.last_perm();
            x_morton_old = x_morton_new;
            for (std::size_t i = 0; i < lp.size(); ++i) {
                x_morton_new[i] = x_morton_old[lp[i]];
            }
            REQUIRE(std::equal(x_morton_new.begin(), x_morton_new.end(), t.p_its_u()[0]));
            REQUIRE(x_morton_new == x_morton_orig);
            // arithmetic updates
            t.update_particles_u([](const auto &p_its) {
                for (size_type idx = 0; idx < s; ++idx) {
                    for (std::size_t j = 0; j < 3; ++j) {
                        p_its[j][idx] += fp_type(1);
                    }
                }
            });
            check_ordered(1, 2, 3, 1, 1);
            lp = t.last_perm();
            x_morton_old = x_morton_new;
            for (std::size_t i = 0; i < lp.size(); ++i) {
                x_morton_new[i] = x_morton_old[lp[i]] + fp_type(1);
            }
            REQUIRE(std::equal(x_morton_new.begin(), x_morton_new.end(), t.p_its_u()[0]));
            t.update_particles_u([](const auto &p_its) {
                for (size_type idx = 0; idx < s; ++idx) {
                    for (std::size_t j = 0; j < 3; ++j) {
                        p_its[j][idx] /= fp_type(2);
                    }
                }
            });
            check_ordered(1, 2, 3, 1, 2);
            // a functor that throws, or moves a particle out of the box, resets the tree (tree.hpp:3760-3764)
            REQUIRE_THROWS_AS(t.update_particles_u([](const auto &) { throw std::runtime_error("boom"); }),
                              std::runtime_error);
            REQUIRE(t.nparts() == 0u);
            REQUIRE(t.box_size() == fp_type(0));
            REQUIRE_THROWS_WITH(t2.update_particles_u([](const auto &p_its) { p_its[0][7] = fp_type(100); }),
                                "outside the allowed bounds");
            REQUIRE(t2.nparts() == 0u);
        });
    });
}

TEST_CASE("update masses")
{
    // test/update_masses.cpp:34-151
    tuple_for_each(macs{}, [](auto mac_type) {
        tuple_for_each(fp_types{}, [](auto x) {
            using fp_type = decltype(x);
            constexpr auto bsize = static_cast<fp_type>(1);
            constexpr auto s = 10000u;
            auto parts = get_uniform_particles<3>(s, bsize, rng);
            using tree_t = octree<fp_type, decltype(mac_type)::value>;
            tree_t t{x_coords = parts.begin() + s,
                     y_coords = parts.begin() + 2u * s,
                     z_coords = parts.begin() + 3u * s,
                     masses = parts.begin(),
                     nparts = s,
                     box_size = fp_type(10)};
            const auto t2(t);
            t.update_masses_u([](auto) {});
            REQUIRE(t.nodes() == t2.nodes());
            t.update_masses_o([](auto) {});
            REQUIRE(t.nodes() == t2.nodes());
            auto dbl = [](auto it) {
                for (auto i = 0u; i < s; ++i) {
                    *(it + i) *= 2;
                }
            };
            auto zero = [](auto it) {
                for (auto i = 0u; i < s; ++i) {
                    *(it + i) = 0;
                }
            };
            auto check_doubled = [&]() {
                REQUIRE(t.nodes() != t2.nodes());
                for (std::size_t i = 0; i < t.nodes().size(); ++i) {
                    REQUIRE(t.nodes()[i].props[3] == t2.nodes()[i].props[3] * 2);
                    REQUIRE(std::equal(t.nodes()[i].props, t.nodes()[i].props + 3, t2.nodes()[i].props));
                }
            };
            auto check_zero = [&]() {
                REQUIRE(t.nodes() != t2.nodes());
                for (std::size_t i = 0; i < t.nodes().size(); ++i) {
                    REQUIRE(t.nodes()[i].props[3] == 0);
                    fp_type c_pos[3];
                    get_node_centre(c_pos, t.nodes()[i].code, fp_type(10));
                    REQUIRE(std::equal(c_pos, c_pos + 3, t.nodes()[i].props));
                }
            };
            t.update_masses_u(dbl);
            check_doubled();
            t = t2;
            t.update_masses_o(dbl);
            check_doubled();
            t = t2;
            t.update_masses_u(zero);
            check_zero();
            t = t2;
            t.update_masses_o(zero);
            check_zero();
            // individual masses through the ordered iterator end up at the right particle
            t = t2;
            const std::vector<unsigned> indices{1u, 100u, 123u, 1045u, 9800u};
            t.update_masses_o([&indices](auto it) {
                for (auto idx : indices) {
                    *(it + idx) = fp_type(42);
                }
            });
            for (auto idx : indices) {
                REQUIRE(t.p_its_o()[3][idx] == fp_type(42));
                REQUIRE(t.p_its_u()[3][t.inv_perm()[idx]] == fp_type(42));
            }
            // exceptions reset the tree; non-finite masses are rejected (update_masses.cpp:134-148)
            REQUIRE_THROWS_AS(t.update_masses_u([](auto) { throw std::invalid_argument("x"); }),
                              std::invalid_argument);
            REQUIRE(t.nparts() == 0u);
            t = t2;
            REQUIRE_THROWS_AS(
                t.update_masses_u([](auto it) { *(it + 5) = std::numeric_limits<fp_type>::infinity(); }),
                std::invalid_argument);
            REQUIRE(t.nparts() == 0u);
        });
    });
}

MINI_TEST_MAIN()
