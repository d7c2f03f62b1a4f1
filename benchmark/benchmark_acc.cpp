// Counterpart of the reference's benchmark/benchmark_acc.cpp: same command line, same printed lines (the tree,
// the tree acceleration on particle --idx, the exact acceleration on it), built on the drop-in header. Extra:
// wall-clock timers around construction and evaluation (host buffers in, host buffers out), a second evaluation
// (the first one pays the one-off CUDA module load), and --pots for accelerations + potentials.
#include <array>
#include <memory>
#include <iostream>
#include <type_traits>
#include <vector>

#include "common.hpp"

using namespace rakau;
using namespace rakau_benchmark;

template <typename F, mac M>
static void run(const accpot_options &o)
{
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
    std::array<std::vector<F>, 3> accs;
    const F theta = static_cast<F>(o.mac_value);
    for (int rep = 0; rep < 2; ++rep) {
        simple_timer st(rep ? "acceleration evaluation (host buffers out)" : "first acceleration evaluation (CUDA module load)");
        if (o.ordered) {
            t.accs_o(accs, theta, kwargs::split = o.split);
        } else {
            t.accs_u(accs, theta, kwargs::split = o.split);
        }
    }
    if (o.ordered) {
        std::cout << accs[0][o.idx] << ", " << accs[1][o.idx] << ", " << accs[2][o.idx] << '\n';
        auto eacc = t.exact_acc_o(o.idx);
        std::cout << eacc[0] << ", " << eacc[1] << ", " << eacc[2] << '\n';
    } else {
        const auto i = t.inv_perm()[o.idx];
        std::cout << accs[0][i] << ", " << accs[1][i] << ", " << accs[2][i] << '\n';
        auto eacc = t.exact_acc_u(i);
        std::cout << eacc[0] << ", " << eacc[1] << ", " << eacc[2] << '\n';
    }
}

int main(int argc, char **argv)
{
    std::cout.precision(20);
    try {
        const auto o = parse_accpot_benchmark_options(argc, argv);
        if (o.fp_type == "float") {
            o.mac_type == "bh" ? run<float, mac::bh>(o) : run<float, mac::bh_geom>(o);
        } else {
            o.mac_type == "bh" ? run<double, mac::bh>(o) : run<double, mac::bh_geom>(o);
        }
    } catch (const std::exception &e) {
        std::cerr << "error: " << e.what() << '\n';
        return 1;
    }
    return 0;
}
