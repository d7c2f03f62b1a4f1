// plummer.cpp — synthetic inputs of the reference's benchmarks, host side (part of librakau_b200.so so that
// bench.py and the C++ harnesses never need anything from oracle/).
//
// Follows benchmark/common.hpp:39-126 (get_plummer_sphere): std::mt19937, libstdc++
// uniform_real_distribution, masses U[0.1, 1.9), r = a / sqrt(u^(-2/3) - 1) with non-finite r rejected,
// sphere point picking with clamped longitude / colatitude, optional clipping to |coord| < size/2 - size/100.
//   mode 0: the sequential branch (default-seeded engine; all masses first, then the positions);
//   mode 1: the parallel branch made deterministic — fixed chunks, each chunk's engine seeded with its first
//           index as at common.hpp:66, masses interleaved with the position draws.

#include "../../include/rakau_b200.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <limits>
#include <random>
#include <thread>
#include <vector>

namespace
{

template <typename F>
struct sampler {
    F a, lim;
    std::uniform_real_distribution<F> udist{F(0), F(1)};
    sampler(F a_, F size) : a(a_), lim(size > F(0) ? (size / F(2) - size / F(100)) : std::numeric_limits<F>::infinity()) {}
    // Draws until a point inside the bounds comes out.
    template <typename Rng>
    void point(Rng &rng, F &x, F &y, F &z)
    {
        const F pi = static_cast<F>(3.141592653589793238462643383279502884L);
        for (;;) {
            F r;
            do {
                r = a / std::sqrt(std::pow(udist(rng), F(-2) / F(3)) - F(1));
            } while (!std::isfinite(r));
            const F u = udist(rng), v = udist(rng);
            const F lon = std::clamp(F(2) * pi * u, F(0), F(2) * pi);
            const F colat = std::acos(std::clamp(F(2) * v - F(1), F(-1), F(1)));
            x = r * std::cos(lon) * std::sin(colat);
            y = r * std::sin(lon) * std::sin(colat);
            z = r * std::cos(colat);
            if (x >= -lim && x < lim && y >= -lim && y < lim && z >= -lim && z < lim) {
                return;
            }
        }
    }
};

template <typename F>
void gen_sequential(size_t n, F a, F size, F *m, F *x, F *y, F *z)
{
    std::mt19937 rng;
    std::uniform_real_distribution<F> mdist(F(0.1), F(1.9));
    for (size_t i = 0; i < n; ++i) {
        m[i] = mdist(rng);
    }
    sampler<F> s(a, size);
    for (size_t i = 0; i < n; ++i) {
        s.point(rng, x[i], y[i], z[i]);
    }
}

template <typename F>
void gen_chunked(size_t n_total, size_t first, size_t count, F a, F size, size_t chunk, int nthreads, F *m, F *x, F *y,
                 F *z)
{
    const size_t c0 = first / chunk, c1 = (first + count + chunk - 1) / chunk;
    std::atomic<size_t> next{c0};
    auto worker = [&]() {
        for (;;) {
            const size_t c = next.fetch_add(1);
            if (c >= c1) {
                break;
            }
            const size_t b = c * chunk, e = std::min(n_total, std::min(first + count, b + chunk));
            std::mt19937 rng;
            rng.seed(static_cast<std::mt19937::result_type>(b));
            std::uniform_real_distribution<F> mdist(F(0.1), F(1.9));
            sampler<F> s(a, size);
            for (size_t i = b; i < e; ++i) {
                const size_t o = i - first;
                m[o] = mdist(rng);
                s.point(rng, x[o], y[o], z[o]);
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

} // namespace

// Initial conditions of benchmark/benchmark_leapfrog.cpp:49-102, 191-215: Plummer positions AND velocities
// (Aarseth-Henon-Wielen rejection sampling), G = M = 1, one default-seeded mt19937 stream, then clipping at
// 10 core radii. Returns the number of particles kept (the first *kept entries of every array).
template <typename F>
static size_t gen_leapfrog(size_t n, F a, F *x, F *y, F *z, F *vx, F *vy, F *vz)
{
    const F pi = static_cast<F>(3.141592653589793238462643383279502884L);
    const F G(1), M(1);
    std::mt19937 rng;
    std::uniform_real_distribution<F> dist1, dist2(F(-1), F(1)), dist3(F(0), F(2) * pi), dist4(F(0), F(1) / F(10));
    auto rej_sample = [&]() {
        F xx = F(0), yy = F(1) / F(10);
        while (yy > xx * xx * std::pow(F(1) - xx * xx, F(7) / F(2))) {
            xx = dist1(rng);
            yy = dist4(rng);
        }
        return xx;
    };
    size_t kept = 0;
    for (size_t i = 0; i < n; ++i) {
        const F r = a / std::sqrt(std::pow(dist1(rng), -F(2) / F(3)) - F(1)), theta = std::acos(dist2(rng)),
                phi = dist3(rng);
        const F px = r * std::sin(theta) * std::cos(phi), py = r * std::sin(theta) * std::sin(phi),
                pz = r * std::cos(theta);
        const F q = rej_sample();
        const F v = q * std::sqrt(F(2) * G * M / a) * std::pow(F(1) + r * r / (a * a), -F(1) / F(4));
        const F theta_v = std::acos(dist2(rng)), phi_v = dist3(rng);
        if (px * px + py * py + pz * pz < F(100) * a * a) {
            x[kept] = px;
            y[kept] = py;
            z[kept] = pz;
            vx[kept] = v * std::sin(theta_v) * std::cos(phi_v);
            vy[kept] = v * std::sin(theta_v) * std::sin(phi_v);
            vz[kept] = v * std::cos(theta_v);
            ++kept;
        }
    }
    return kept;
}

extern "C" int rk_plummer_leapfrog(int fp_bits, size_t n, double a, void *x, void *y, void *z, void *vx, void *vy,
                                   void *vz, size_t *kept)
{
    if ((fp_bits != 32 && fp_bits != 64) || !std::isfinite(a) || a <= 0 || !kept) {
        return RK_ERR_INVALID_ARGUMENT;
    }
    if (fp_bits == 32) {
        *kept = gen_leapfrog<float>(n, float(a), static_cast<float *>(x), static_cast<float *>(y),
                                    static_cast<float *>(z), static_cast<float *>(vx), static_cast<float *>(vy),
                                    static_cast<float *>(vz));
    } else {
        *kept = gen_leapfrog<double>(n, a, static_cast<double *>(x), static_cast<double *>(y), static_cast<double *>(z),
                                     static_cast<double *>(vx), static_cast<double *>(vy), static_cast<double *>(vz));
    }
    return RK_OK;
}

extern "C" int rk_plummer(int fp_bits, size_t n_total, size_t first, size_t count, double a, double size, int mode,
                          size_t chunk, int nthreads, void *m, void *x, void *y, void *z)
{
    if ((fp_bits != 32 && fp_bits != 64) || !std::isfinite(a) || a <= 0 || !std::isfinite(size) || size < 0
        || first + count > n_total) {
        return RK_ERR_INVALID_ARGUMENT;
    }
    if (mode == 0) {
        if (first != 0 || count != n_total) {
            return RK_ERR_INVALID_ARGUMENT;
        }
        if (fp_bits == 32) {
            gen_sequential<float>(count, float(a), float(size), static_cast<float *>(m), static_cast<float *>(x),
                                  static_cast<float *>(y), static_cast<float *>(z));
        } else {
            gen_sequential<double>(count, a, size, static_cast<double *>(m), static_cast<double *>(x),
                                   static_cast<double *>(y), static_cast<double *>(z));
        }
        return RK_OK;
    }
    if (!chunk || first % chunk) {
        return RK_ERR_INVALID_ARGUMENT;
    }
    if (fp_bits == 32) {
        gen_chunked<float>(n_total, first, count, float(a), float(size), chunk, nthreads, static_cast<float *>(m),
                           static_cast<float *>(x), static_cast<float *>(y), static_cast<float *>(z));
    } else {
        gen_chunked<double>(n_total, first, count, a, size, chunk, nthreads, static_cast<double *>(m),
                            static_cast<double *>(x), static_cast<double *>(y), static_cast<double *>(z));
    }
    return RK_OK;
}
