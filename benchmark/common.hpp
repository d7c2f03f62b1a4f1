// Command-line front end shared by the benchmark programs: the options of the reference's
// benchmark/common.hpp:143-229 (parse_accpot_benchmark_options) with the same names, defaults and checks, parsed by
// hand (the reference uses boost::program_options, which this tree does not depend on). --nthreads is accepted
// and ignored (there is no host thread pool on this path); --parinit selects the chunked generator.
#ifndef RAKAU_B200_BENCHMARK_COMMON_HPP
#define RAKAU_B200_BENCHMARK_COMMON_HPP

#include <chrono>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include <rakau/tree.hpp>

namespace rakau_benchmark
{

struct accpot_options {
    unsigned long nparts = 1'000'000ul, idx = 0;
    unsigned max_leaf_n = rakau::default_max_leaf_n, ncrit = rakau::default_ncrit, nthreads = 0;
    double bsize = 0., a = 1., mac_value = 0.75;
    bool parinit = false, ordered = false;
    std::vector<double> split;
    std::string fp_type = "float", mac_type = "bh";
};

inline accpot_options parse_accpot_benchmark_options(int argc, char **argv)
{
    accpot_options o;
    auto value = [&](int &i) -> std::string {
        std::string s = argv[i];
        const auto eq = s.find('=');
        if (eq != std::string::npos) {
            return s.substr(eq + 1);
        }
        if (i + 1 >= argc) {
            throw std::invalid_argument("the option '" + s + "' requires a value");
        }
        return argv[++i];
    };
    for (int i = 1; i < argc; ++i) {
        std::string s = argv[i];
        const std::string name = s.substr(0, s.find('='));
        if (name == "--help") {
            std::cout << "Allowed options:\n  --help\n  --nparts arg (=1000000)\n  --idx arg (=0)\n  --max_leaf_n arg (="
                      << rakau::default_max_leaf_n << ")\n  --ncrit arg (=" << rakau::default_ncrit
                      << ")\n  --a arg (=1)\n  --bsize arg (=0)\n  --nthreads arg (=0)\n  --mac_value arg (=0.75)\n"
                         "  --parinit\n  --split arg...\n  --fp_type arg (=float)\n  --mac_type arg (=bh)\n  --ordered\n";
            std::exit(0);
        } else if (name == "--nparts") {
            o.nparts = std::stoul(value(i));
        } else if (name == "--idx") {
            o.idx = std::stoul(value(i));
        } else if (name == "--max_leaf_n") {
            o.max_leaf_n = static_cast<unsigned>(std::stoul(value(i)));
        } else if (name == "--ncrit") {
            o.ncrit = static_cast<unsigned>(std::stoul(value(i)));
        } else if (name == "--nthreads") {
            o.nthreads = static_cast<unsigned>(std::stoul(value(i)));
        } else if (name == "--a") {
            o.a = std::stod(value(i));
        } else if (name == "--bsize") {
            o.bsize = std::stod(value(i));
        } else if (name == "--mac_value") {
            o.mac_value = std::stod(value(i));
        } else if (name == "--fp_type") {
            o.fp_type = value(i);
        } else if (name == "--mac_type") {
            o.mac_type = value(i);
        } else if (name == "--parinit") {
            o.parinit = true;
        } else if (name == "--ordered") {
            o.ordered = true;
        } else if (name == "--split") { // multitoken: every following token that is not an option
            if (s.find('=') != std::string::npos) {
                o.split.push_back(std::stod(s.substr(s.find('=') + 1)));
            }
            while (i + 1 < argc && std::string(argv[i + 1]).rfind("--", 0) != 0) {
                o.split.push_back(std::stod(argv[++i]));
            }
        } else {
            throw std::invalid_argument("unrecognised option '" + s + "'");
        }
    }
    if (o.nparts == 0u) {
        throw std::invalid_argument("The number of particles cannot be zero");
    }
    if (o.idx >= o.nparts) {
        throw std::invalid_argument(
            "The index of the particle to test against needs to be less-than the total number of particles ("
            + std::to_string(o.nparts) + ")");
    }
    if (o.fp_type != "float" && o.fp_type != "double") {
        throw std::invalid_argument("Only the 'float' and 'double' floating-point types are supported, but the type '"
                                    + o.fp_type + "' was specified instead");
    }
    if (o.mac_type != "bh" && o.mac_type != "bh_geom") {
        throw std::invalid_argument("'" + o.mac_type + "' is not a valid MAC type");
    }
    return o;
}

// Scoped wall-clock timer printing like the reference's simple_timer (detail/simple_timer.hpp).
struct simple_timer {
    explicit simple_timer(const char *d) : desc(d), start(std::chrono::steady_clock::now()) {}
    ~simple_timer()
    {
        const auto us
            = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - start).count();
        std::cout << "Elapsed time for '" << desc << "': " << us << u8"μs\n";
    }
    const char *desc;
    std::chrono::steady_clock::time_point start;
};

// Plummer sphere in the reference's layout: masses first, then x, y, z (benchmark/common.hpp:39-126). The
// generator is the library's (rk_plummer): the sequential default-seeded mt19937 stream of the reference, or the
// deterministic chunked variant for --parinit.
template <typename F>
inline std::vector<F> get_plummer_sphere(unsigned long n, F a, F size, bool parinit)
{
    simple_timer st("plummer init");
    std::vector<F> v(4u * n);
    const int rc = rk_plummer(sizeof(F) * 8, n, 0, n, static_cast<double>(a), static_cast<double>(size), parinit ? 1 : 0,
                              parinit ? (1u << 20) : 0, 0, v.data(), v.data() + n, v.data() + 2 * n, v.data() + 3 * n);
    if (rc) {
        throw std::runtime_error("rk_plummer failed");
    }
    return v;
}

} // namespace rakau_benchmark

#endif
