// Fixtures of the reference's test/test_utils.hpp: uniform particles [m | x | y | z] and a median helper.
#ifndef RAKAU_B200_TEST_UTILS_HPP
#define RAKAU_B200_TEST_UTILS_HPP

#include <algorithm>
#include <cstddef>
#include <random>
#include <vector>

namespace rakau_test
{
template <typename T>
inline T median(std::vector<T> &v)
{
    std::sort(v.begin(), v.end());
    const auto h = v.size() / 2u;
    return (v.size() % 2u) ? v[h] : (v[h - 1u] + v[h]) / T(2);
}

// Masses U[0, 1) first, then coordinates U[-size/2, size/2) (test/test_utils.hpp:41-59).
template <std::size_t D, typename F, typename Rng>
inline std::vector<F> get_uniform_particles(std::size_t n, F size, Rng &rng)
{
    std::vector<F> out(n * (D + 1u));
    std::uniform_real_distribution<F> mdist(F(0), F(1));
    std::generate(out.begin(), out.begin() + static_cast<std::ptrdiff_t>(n), [&]() { return mdist(rng); });
    std::uniform_real_distribution<F> rdist(-size / F(2), size / F(2));
    std::generate(out.begin() + static_cast<std::ptrdiff_t>(n), out.end(), [&]() { return rdist(rng); });
    return out;
}
} // namespace rakau_test
#endif
