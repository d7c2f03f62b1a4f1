// The `split` keyword argument across several devices driven by ONE process, as the reference drives its accelerators
// (tree.hpp:3147-3198, src/rakau_cuda.cu:492-527): accs_o / accs_pots_u with split = {0, 1, 1, ...} must reproduce the
// one-device result bit for bit (the runs of the traversal do not depend on how the critical nodes are cut). With fewer
// devices than shares the reference's error is raised (tree.hpp:3134-3141).
#include <array>
#include <random>
#include <string>
#include <vector>

#include <rakau/tree.hpp>

#include "mini_test.hpp"
#include "test_utils.hpp"

using namespace rakau;
using namespace rakau::kwargs;
using namespace rakau_test;

static std::mt19937 rng(7);

TEST_CASE("split over the devices of one process")
{
    constexpr std::size_t N = 300000;
    auto parts = get_uniform_particles<3>(N, 10.f, rng);
    octree<float> t{x_coords = parts.data() + N, y_coords = parts.data() + 2 * N, z_coords = parts.data() + 3 * N,
                    masses = parts.data(), nparts = N};
    const unsigned ndev = rk_device_count();
    std::array<std::vector<float>, 4> ref_o, ref_u;
    t.accs_pots_o(ref_o, 0.75f, eps = 0.01f, G = 2.f);
    t.accs_pots_u(ref_u, 0.75f, eps = 0.01f, G = 2.f);
    // one share per device, equal and unequal weights
    for (const double first : {1., 3.}) {
        if (ndev < 2u) {
            break;
        }
        std::vector<double> sp{0., first};
        for (unsigned d = 1; d < ndev; ++d) {
            sp.push_back(1.);
        }
        std::array<std::vector<float>, 4> got_o, got_u;
        t.accs_pots_o(got_o, 0.75f, eps = 0.01f, G = 2.f, split = sp);
        t.accs_pots_u(got_u, 0.75f, eps = 0.01f, G = 2.f, split = sp);
        for (std::size_t j = 0; j < 4u; ++j) {
            REQUIRE(got_o[j] == ref_o[j]);
            REQUIRE(got_u[j] == ref_u[j]);
        }
        // a rebuilt tree refreshes the mirrors on the other devices
        t.update_particles_u([&](const auto &its) {
            for (std::size_t i = 0; i < N; ++i) {
                its[0][i] += 0.001f;
            }
        });
        t.accs_pots_o(ref_o, 0.75f, eps = 0.01f, G = 2.f);
        t.accs_pots_u(ref_u, 0.75f, eps = 0.01f, G = 2.f);
        t.accs_pots_o(got_o, 0.75f, eps = 0.01f, G = 2.f, split = sp);
        for (std::size_t j = 0; j < 4u; ++j) {
            REQUIRE(got_o[j] == ref_o[j]);
        }
    }
    // more accelerator shares than devices: the reference's message
    std::vector<double> too_many(ndev + 2u, 1.);
    std::array<std::vector<float>, 3> accs;
    REQUIRE_THROWS_WITH(t.accs_u(accs, 0.75f, split = too_many), "accelerators, but only");
}

MINI_TEST_MAIN()
