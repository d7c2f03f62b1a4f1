// Counterpart of the reference's benchmark/benchmark_leapfrog.cpp on the drop-in header: the same options (--nparts
// --max_leaf_n --ncrit --a --nthreads --mac_value --split --fp_type --mac_type --timestep --track-integrals, reference
// lines 118-136), the same initial conditions (Plummer sphere with velocities clipped at 10 core radii, equal masses,
// eps = 0.45 N^-0.73) and the same kick-drift-kick loop. Two ways of running the loop:
//   default         the reference's own structure (286-384): host functors around update_particles_u, accelerations
//                   written to host vectors, velocities updated on the host through last_perm: every step crosses PCIe;
//   --device        the library's device-resident integrator (octree::leapfrog_init / leapfrog_step =
//                   rk_tree_leapfrog_*): positions and velocities stay on the GPU.
// The reference's loop never ends; here --steps (default 10) bounds it and the time per step is printed.
#include <array>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include <rakau/tree.hpp>

using namespace rakau;

namespace
{
struct options {
    unsigned long nparts = 1'000'000ul, steps = 10;
    unsigned max_leaf_n = default_max_leaf_n, ncrit = default_ncrit, nthreads = 0;
    double a = 1., mac_value = 0.75, timestep = 1E-4;
    std::vector<double> split;
    std::string fp_type = "float", mac_type = "bh";
    bool track_integrals = false, device = false;
};

options parse(int argc, char **argv)
{
    options o;
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
        const std::string s = argv[i], name = s.substr(0, s.find('='));
        if (name == "--help") {
            std::cout << "Allowed options:\n  --help\n  --nparts arg (=1000000)\n  --max_leaf_n arg\n  --ncrit arg\n"
                         "  --a arg (=1)\n  --nthreads arg (=0)\n  --mac_value arg (=0.75)\n  --split arg...\n"
                         "  --fp_type arg (=float)\n  --mac_type arg (=bh)\n  --timestep arg (=0.0001)\n"
                         "  --track-integrals\n  --steps arg (=10)\n  --device\n";
            std::exit(0);
        } else if (name == "--nparts") {
            o.nparts = std::stoul(value(i));
        } else if (name == "--steps") {
            o.steps = std::stoul(value(i));
        } else if (name == "--max_leaf_n") {
            o.max_leaf_n = static_cast<unsigned>(std::stoul(value(i)));
        } else if (name == "--ncrit") {
            o.ncrit = static_cast<unsigned>(std::stoul(value(i)));
        } else if (name == "--nthreads") {
            o.nthreads = static_cast<unsigned>(std::stoul(value(i)));
        } else if (name == "--a") {
            o.a = std::stod(value(i));
        } else if (name == "--mac_value") {
            o.mac_value = std::stod(value(i));
        } else if (name == "--timestep") {
            o.timestep = std::stod(value(i));
        } else if (name == "--fp_type") {
            o.fp_type = value(i);
        } else if (name == "--mac_type") {
            o.mac_type = value(i);
        } else if (name == "--track-integrals") {
            o.track_integrals = true;
        } else if (name == "--device") {
            o.device = true;
        } else if (name == "--split") {
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
    // the reference's checks, benchmark_leapfrog.cpp:146-171
    if (o.nparts == 0u) {
        throw std::invalid_argument("The number of particles cannot be zero");
    }
    if (o.fp_type != "float" && o.fp_type != "double") {
        throw std::invalid_argument("Only the 'float' and 'double' floating-point types are supported, but the type '"
                                    + o.fp_type + "' was specified instead");
    }
    if (o.mac_type != "bh" && o.mac_type != "bh_geom") {
        throw std::invalid_argument("'" + o.mac_type + "' is not a valid MAC type");
    }
    if (!std::isfinite(o.a) || o.a <= 0.) {
        throw std::invalid_argument("The Plummer core radius must be finite and positive, but it is "
                                    + std::to_string(o.a) + " instead");
    }
    if (!std::isfinite(o.timestep) || o.timestep <= 0.) {
        throw std::invalid_argument("The integration timestep must be finite and positive, but it is "
                                    + std::to_string(o.timestep) + " instead");
    }
    return o;
}

template <typename F, mac M>
void run(const options &o)
{
    std::cout << "Building the Plummer distribution...\n";
    std::vector<F> xp(o.nparts), yp(o.nparts), zp(o.nparts), xv(o.nparts), yv(o.nparts), zv(o.nparts);
    std::size_t kept = 0;
    if (rk_plummer_leapfrog(sizeof(F) * 8, o.nparts, o.a, xp.data(), yp.data(), zp.data(), xv.data(), yv.data(), zv.data(),
                            &kept)) {
        throw std::runtime_error("rk_plummer_leapfrog failed");
    }
    for (auto *v : {&xp, &yp, &zp, &xv, &yv, &zv}) {
        v->resize(kept);
    }
    const std::size_t nparts = kept;
    std::cout << "After clipping, nparts is " << nparts << "\nDone\n";
    const auto eps = static_cast<F>(0.45) * std::pow(static_cast<F>(nparts), static_cast<F>(-0.73));
    std::cout << "Softening length: " << eps << '\n';
    std::vector<F> masses(nparts, F(1) / nparts);
    octree<F, M> t{kwargs::x_coords = xp, kwargs::y_coords = yp, kwargs::z_coords = zp, kwargs::masses = masses,
                   kwargs::max_leaf_n = o.max_leaf_n, kwargs::ncrit = o.ncrit};
    std::cout << "Box size: " << t.box_size() << '\n';
    const F timestep = static_cast<F>(o.timestep), half_timestep = timestep / F(2), mac_value = static_cast<F>(o.mac_value);
    using clock = std::chrono::steady_clock;

    if (o.device) {
        t.leapfrog_init(xv.data(), yv.data(), zv.data(), mac_value, o.track_integrals, kwargs::eps = eps);
        double tot = 0;
        for (unsigned long s = 0; s < o.steps; ++s) {
            const auto info = t.leapfrog_step(timestep);
            if (o.track_integrals) {
                std::cout << "Centre of mass: " << info.com[0] << ", " << info.com[1] << ", " << info.com[2] << '\n';
                std::cout << "Centre of mass velocity: " << info.com_v[0] << ", " << info.com_v[1] << ", " << info.com_v[2]
                          << '\n';
                std::cout << "Total energy: " << info.energy << '\n';
            }
            std::cout << "step " << s << ": " << info.ms_step << " ms on the device (kick+drift " << info.ms_kick_drift
                      << ", rebuild " << info.ms_rebuild << ", traversal " << info.ms_traverse << ", velocity update "
                      << info.ms_reindex << ")\n";
            tot += info.ms_step;
        }
        std::cout << "Average time per step (device-resident): " << tot / double(o.steps) << " ms\n";
        return;
    }

    // ---- the reference's loop, benchmark_leapfrog.cpp:233-384 ----
    std::vector<F> acc_x(nparts), acc_y(nparts), acc_z(nparts), pots;
    std::array<F *, 3> acc_its{acc_x.data(), acc_y.data(), acc_z.data()};
    std::array<F *, 4> acc_pot_its{};
    if (o.track_integrals) {
        pots.resize(nparts);
        acc_pot_its = {acc_x.data(), acc_y.data(), acc_z.data(), pots.data()};
    }
    std::vector<F> kick_x_vel(nparts), kick_y_vel(nparts), kick_z_vel(nparts);
    auto &tmp_buffer = xp;
    auto reorder = [&t, &tmp_buffer](auto &vec) {
        const auto &lp = t.last_perm();
        for (std::size_t i = 0; i < lp.size(); ++i) {
            tmp_buffer[i] = vec[lp[i]];
        }
        vec.swap(tmp_buffer);
    };
    reorder(xv);
    reorder(yv);
    reorder(zv);
    auto compute_accs_pots = [&]() {
        if (o.track_integrals) {
            t.accs_pots_u(acc_pot_its, mac_value, kwargs::split = o.split, kwargs::eps = eps);
        } else {
            t.accs_u(acc_its, mac_value, kwargs::split = o.split, kwargs::eps = eps);
        }
    };
    compute_accs_pots();
    double tot = 0;
    for (unsigned long s = 0; s < o.steps; ++s) {
        const auto t0 = clock::now();
        if (o.track_integrals) {
            const auto p = t.p_its_u();
            double com[3] = {0, 0, 0}, com_v[3] = {0, 0, 0}, tot_E = 0;
            for (std::size_t i = 0; i < nparts; ++i) {
                com[0] += p[0][i];
                com[1] += p[1][i];
                com[2] += p[2][i];
                com_v[0] += xv[i];
                com_v[1] += yv[i];
                com_v[2] += zv[i];
                const auto v2 = xv[i] * xv[i] + yv[i] * yv[i] + zv[i] * zv[i];
                tot_E += (F(1) / F(2)) * (F(1) / nparts) * v2 + pots[i];
            }
            std::cout << "Centre of mass: " << com[0] / nparts << ", " << com[1] / nparts << ", " << com[2] / nparts << '\n';
            std::cout << "Centre of mass velocity: " << com_v[0] / nparts << ", " << com_v[1] / nparts << ", "
                      << com_v[2] / nparts << '\n';
            std::cout << "Total energy: " << tot_E << '\n';
        }
        for (std::size_t i = 0; i < nparts; ++i) {
            kick_x_vel[i] = std::fma(acc_its[0][i], half_timestep, xv[i]);
            kick_y_vel[i] = std::fma(acc_its[1][i], half_timestep, yv[i]);
            kick_z_vel[i] = std::fma(acc_its[2][i], half_timestep, zv[i]);
        }
        t.update_particles_u([&](const auto &p_its) {
            const auto [x_it, y_it, z_it, m_it] = p_its;
            (void)m_it;
            for (std::size_t i = 0; i < nparts; ++i) {
                *(x_it + i) = std::fma(kick_x_vel[i], timestep, *(x_it + i));
                *(y_it + i) = std::fma(kick_y_vel[i], timestep, *(y_it + i));
                *(z_it + i) = std::fma(kick_z_vel[i], timestep, *(z_it + i));
            }
        });
        compute_accs_pots();
        {
            const auto &lp = t.last_perm();
            for (std::size_t i = 0; i < nparts; ++i) {
                xv[i] = std::fma(acc_its[0][i], half_timestep, kick_x_vel[lp[i]]);
                yv[i] = std::fma(acc_its[1][i], half_timestep, kick_y_vel[lp[i]]);
                zv[i] = std::fma(acc_its[2][i], half_timestep, kick_z_vel[lp[i]]);
            }
        }
        const double ms = std::chrono::duration<double, std::milli>(clock::now() - t0).count();
        std::cout << "step " << s << ": " << ms << " ms end to end (host functors, host vectors)\n";
        tot += ms;
    }
    std::cout << "Average time per step (host functors): " << tot / double(o.steps) << " ms\n";
}
} // namespace

int main(int argc, char **argv)
{
    try {
        const auto o = parse(argc, argv);
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
