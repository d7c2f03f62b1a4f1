// Counterpart of the reference's benchmark/benchmark_move.cpp: evaluate, shift every particle by bsize / 1000 along
// each axis through update_particles_u (a full re-sort and rebuild: `sync`, tree.hpp:3678-3743), evaluate again. Same
// command line and printed lines as the reference program (the tree, then tree result / exact result on particle --idx
// before and after the move). Extra: wall-clock timers around the three phases.
#include <array>
#include <iostream>
#include <type_traits>
#include <vector>

#include "common.hpp"

using namespace rakau;
using namespace rakau_benchmark;

template <typename F, mac M>
static void run_move(const accpot_options &o)
{
    const auto n = o.nparts;
    auto parts = get_plummer_sphere(n, static_cast<F>(o.a), static_cast<F>(o.bsize), o.parinit);
    octree<F, M> t{kwargs::x_coords = parts.data() + n,
                   kwargs::y_coords = parts.data() + 2 * n,
                   kwargs::z_coords = parts.data() + 3 * n,
                   kwargs::masses = parts.data(),
                   kwargs::nparts = n,
                   kwargs::max_leaf_n = o.max_leaf_n,
                   kwargs::ncrit = o.ncrit};
    std::cout << t << '\n';
    const F theta = static_cast<F>(o.mac_value);
    std::array<std::vector<F>, 3> accs;
    auto report = [&](const char *what) {
        {
            rakau_benchmark::simple_timer st(what);
            t.accs_u(accs, theta, kwargs::split = o.split);
        }
        const auto i = t.inv_perm()[o.idx];
        std::cout << accs[0][i] << ", " << accs[1][i] << ", " << accs[2][i] << '\n';
        const auto e = t.exact_acc_u(i);
        std::cout << e[0] << ", " << e[1] << ", " << e[2] << '\n';
    };
    report("evaluation before the move");
    {
        rakau_benchmark::simple_timer st("update_particles_u (host functor + re-sort + rebuild)");
        const F shift = static_cast<F>(o.bsize) / F(1000);
        t.update_particles_u([shift, n](const auto &p_its) {
            for (int d = 0; d < 3; ++d) {
                auto it = p_its[d];
                for (std::remove_const_t<decltype(n)> k = 0; k < n; ++k) {
                    it[k] += shift;
                }
            }
        });
    }
    report("evaluation after the move");
}

int main(int argc, char **argv)
{
    std::cout.precision(20);
    try {
        const auto o = parse_accpot_benchmark_options(argc, argv);
        if (o.fp_type == "float") {
            o.mac_type == "bh" ? run_move<float, mac::bh>(o) : run_move<float, mac::bh_geom>(o);
        } else {
            o.mac_type == "bh" ? run_move<double, mac::bh>(o) : run_move<double, mac::bh_geom>(o);
        }
    } catch (const std::exception &e) {
        std::cerr << "error: " << e.what() << '\n';
        return 1;
    }
    return 0;
}
