// Shared driver of benchmark_acc / benchmark_pot / benchmark_acc_pot: counterparts of the reference's
// benchmark/benchmark_acc.cpp, benchmark_pot.cpp and benchmark_acc_pot.cpp on the drop-in header. Same command line
// (benchmark/common.hpp:143-229 of the reference), same printed lines (the tree, the tree result on particle --idx, the
// exact result on it). Extra: wall-clock timers around construction and evaluation (host buffers in, host buffers out)
// and a second evaluation (the first one pays the one-off CUDA module load).
#ifndef RAKAU_B200_BENCHMARK_ACCPOT_MAIN_HPP
#define RAKAU_B200_BENCHMARK_ACCPOT_MAIN_HPP

#include <array>
#include <iostream>
#include <memory>
#include <vector>

#include "common.hpp"

namespace rakau_benchmark
{

// Q: 0 accelerations, 1 potentials, 2 both (detail/tree_fwd.hpp:129-137).
template <unsigned Q, typename F, rakau::mac M>
inline void run_accpot(const accpot_options &o)
{
    using namespace rakau;
    auto parts = get_plummer_sphere(o.nparts, static_cast<F>(o.a), static_cast<F>(o.bsize), o.parinit);
    const auto n = o.nparts;
    std::unique_ptr<octree<F, M>> tp;
    {
        simple_timer st("tree construction (host buffers in)");
        tp = std::make_unique<octree<F, M>>(kwargs::x_coords = parts.data() + n, kwargs::y_coords = parts.data() + 2 * n,
                                             kwargs::z_coords = parts.data() + 3 * n, kwargs::masses = parts.data(),
                                             kwargs::nparts = n, kwargs::max_leaf_n = o.max_leaf_n,
                                             kwargs::ncrit = o.ncrit);
    }
    auto &t = *tp;
    std::cout << t << '\n';
    const F theta = static_cast<F>(o.mac_value);
    constexpr std::size_t NOUT = Q == 0 ? 3 : (Q == 1 ? 1 : 4);
    std::array<std::vector<F>, NOUT> out;
    for (int rep = 0; rep < 2; ++rep) {
        simple_timer st(rep ? "evaluation (host buffers out)" : "first evaluation (CUDA module load)");
        if constexpr (Q == 0) {
            o.ordered ? t.accs_o(out, theta, kwargs::split = o.split) : t.accs_u(out, theta, kwargs::split = o.split);
        } else if constexpr (Q == 1) {
            o.ordered ? t.pots_o(out[0], theta, kwargs::split = o.split) : t.pots_u(out[0], theta, kwargs::split = o.split);
        } else {
            o.ordered ? t.accs_pots_o(out, theta, kwargs::split = o.split)
                      : t.accs_pots_u(out, theta, kwargs::split = o.split);
        }
    }
    const auto i = o.ordered ? o.idx : t.inv_perm()[o.idx];
    if constexpr (Q == 1) {
        std::cout << out[0][i] << '\n';
        std::cout << (o.ordered ? t.exact_pot_o(o.idx) : t.exact_pot_u(i)) << '\n';
    } else {
        std::cout << out[0][i] << ", " << out[1][i] << ", " << out[2][i] << '\n';
        const auto e = o.ordered ? t.exact_acc_pot_o(o.idx) : t.exact_acc_pot_u(i);
        std::cout << e[0] << ", " << e[1] << ", " << e[2] << '\n';
        if constexpr (Q == 2) {
            std::cout << out[3][i] << '\n' << e[3] << '\n';
        }
    }
}

template <unsigned Q>
inline int accpot_main(int argc, char **argv)
{
    std::cout.precision(20);
    try {
        const auto o = parse_accpot_benchmark_options(argc, argv);
        if (o.fp_type == "float") {
            o.mac_type == "bh" ? run_accpot<Q, float, rakau::mac::bh>(o) : run_accpot<Q, float, rakau::mac::bh_geom>(o);
        } else {
            o.mac_type == "bh" ? run_accpot<Q, double, rakau::mac::bh>(o) : run_accpot<Q, double, rakau::mac::bh_geom>(o);
        }
    } catch (const std::exception &e) {
        std::cerr << "error: " << e.what() << '\n';
        return 1;
    }
    return 0;
}

} // namespace rakau_benchmark

#endif
