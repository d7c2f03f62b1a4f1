// Minimal, from-scratch stand-in for the part of xsimd 7 that rakau uses (batch<T, N>, simd_type, load/store
// modes, fma/fnma/sqrt/hadd/any, the X86 instruction-set macros). It exists only so that the UNMODIFIED
// reference headers compile offline. Batches are GCC vector-extension types (bit-compatible with
// __m256/__m512, so the reference's raw rsqrt intrinsics, detail/simd.hpp:99,117, work unchanged).
#ifndef RAKAU_SHIM_XSIMD_HPP
#define RAKAU_SHIM_XSIMD_HPP

#include <cmath>
#include <cstddef>
#include <cstring>
#include <type_traits>

#include <immintrin.h>

#define XSIMD_X86_SSE2_VERSION 20
#define XSIMD_X86_AVX_VERSION 50
#define XSIMD_X86_FMA3_VERSION 51
#define XSIMD_X86_AVX2_VERSION 52
#define XSIMD_X86_AVX512_VERSION 60

#if defined(__AVX512F__) && defined(__AVX512DQ__)
#define XSIMD_X86_INSTR_SET XSIMD_X86_AVX512_VERSION
#define XSIMD_DEFAULT_ALIGNMENT 64
#define RAKAU_SHIM_VBYTES 64
#elif defined(__AVX2__) && defined(__FMA__)
#define XSIMD_X86_INSTR_SET XSIMD_X86_AVX2_VERSION
#define XSIMD_DEFAULT_ALIGNMENT 32
#define RAKAU_SHIM_VBYTES 32
#else
#error "the xsimd stand-in needs -mavx2 -mfma or -mavx512f -mavx512dq"
#endif

namespace xsimd
{

struct aligned_mode {
};
struct unaligned_mode {
};

// GCC ignores vector_size on dependent types, so the vector typedefs are spelled out per (type, width).
template <typename T, std::size_t N>
struct vec_of;
#define RAKAU_SHIM_VEC(T, N, M)                                                                                        \
    template <>                                                                                                        \
    struct vec_of<T, N> {                                                                                              \
        typedef T type __attribute__((vector_size(N * sizeof(T))));                                                    \
        typedef M mask_type __attribute__((vector_size(N * sizeof(T))));                                               \
        using mask_scalar = M;                                                                                         \
    };
RAKAU_SHIM_VEC(float, 8, int)
RAKAU_SHIM_VEC(double, 4, long long)
RAKAU_SHIM_VEC(float, 16, int)
RAKAU_SHIM_VEC(double, 8, long long)
#undef RAKAU_SHIM_VEC

template <typename T, std::size_t N>
class batch_bool
{
public:
    using mask_t = typename vec_of<T, N>::mask_scalar;
    using vec_t = typename vec_of<T, N>::mask_type;
    vec_t v;
};

template <typename T, std::size_t N>
class batch
{
public:
    using vec_t = typename vec_of<T, N>::type;
    static constexpr std::size_t size = N;
    using value_type = T;
    vec_t v;

    batch() = default;
    batch(T s) : v(vec_t{} + s) {}
    explicit batch(vec_t x) : v(x) {}
    batch(const T *p, aligned_mode) { std::memcpy(&v, __builtin_assume_aligned(p, N * sizeof(T)), sizeof(v)); }
    batch(const T *p, unaligned_mode) { std::memcpy(&v, p, sizeof(v)); }
    // interoperability with the raw intrinsic types
    template <typename U = T, std::enable_if_t<std::is_same_v<U, float> && N == 8, int> = 0>
    batch(__m256 x) : v((vec_t)x) {}
    template <typename U = T, std::enable_if_t<std::is_same_v<U, float> && N == 8, int> = 0>
    operator __m256() const { return (__m256)v; }
#if defined(__AVX512F__)
    template <typename U = T, std::enable_if_t<std::is_same_v<U, float> && N == 16, int> = 0>
    batch(__m512 x) : v((vec_t)x) {}
    template <typename U = T, std::enable_if_t<std::is_same_v<U, float> && N == 16, int> = 0>
    operator __m512() const { return (__m512)v; }
#endif
    void store_aligned(T *p) const { std::memcpy(__builtin_assume_aligned(p, N * sizeof(T)), &v, sizeof(v)); }
    void store_unaligned(T *p) const { std::memcpy(p, &v, sizeof(v)); }
    T operator[](std::size_t i) const { return v[i]; }

    batch &operator+=(const batch &o) { v += o.v; return *this; }
    batch &operator-=(const batch &o) { v -= o.v; return *this; }
    batch &operator*=(const batch &o) { v *= o.v; return *this; }
    batch &operator/=(const batch &o) { v /= o.v; return *this; }
    friend batch operator+(const batch &a, const batch &b) { return batch(a.v + b.v); }
    friend batch operator-(const batch &a, const batch &b) { return batch(a.v - b.v); }
    friend batch operator*(const batch &a, const batch &b) { return batch(a.v * b.v); }
    friend batch operator/(const batch &a, const batch &b) { return batch(a.v / b.v); }
    friend batch operator-(const batch &a) { return batch(-a.v); }
    friend batch_bool<T, N> operator>=(const batch &a, const batch &b) { return batch_bool<T, N>{a.v >= b.v}; }
    friend batch_bool<T, N> operator>(const batch &a, const batch &b) { return batch_bool<T, N>{a.v > b.v}; }
    friend batch_bool<T, N> operator<=(const batch &a, const batch &b) { return batch_bool<T, N>{a.v <= b.v}; }
    friend batch_bool<T, N> operator<(const batch &a, const batch &b) { return batch_bool<T, N>{a.v < b.v}; }
};

// Scalar "batches" (size 1) for types without SIMD support (long double).
template <typename T>
class batch<T, 1>
{
public:
    static constexpr std::size_t size = 1;
    using value_type = T;
    T v;
    batch() = default;
    batch(T s) : v(s) {}
};

template <typename T>
struct simd_traits {
    static constexpr std::size_t size = (std::is_same_v<T, float> || std::is_same_v<T, double>)
                                            ? RAKAU_SHIM_VBYTES / sizeof(T)
                                            : 1;
    using type = batch<T, size>;
};
template <typename T>
using simd_type = typename simd_traits<T>::type;

template <typename B>
struct revert_simd_traits {
    using type = typename B::value_type;
    static constexpr std::size_t size = B::size;
};

template <typename T, std::size_t N>
inline bool any(const batch_bool<T, N> &b)
{
    typename batch_bool<T, N>::mask_t acc = 0;
    for (std::size_t i = 0; i < N; ++i) {
        acc |= b.v[i];
    }
    return acc != 0;
}
template <typename T, std::size_t N>
inline bool all(const batch_bool<T, N> &b)
{
    typename batch_bool<T, N>::mask_t acc = -1;
    for (std::size_t i = 0; i < N; ++i) {
        acc &= b.v[i];
    }
    return acc != 0;
}

// Fused multiply-add on every lane (the FMA3 forms xsimd selects when FMA is available).
inline batch<float, 8> fma(const batch<float, 8> &x, const batch<float, 8> &y, const batch<float, 8> &z)
{
    return batch<float, 8>(_mm256_fmadd_ps((__m256)x.v, (__m256)y.v, (__m256)z.v));
}
inline batch<float, 8> fnma(const batch<float, 8> &x, const batch<float, 8> &y, const batch<float, 8> &z)
{
    return batch<float, 8>(_mm256_fnmadd_ps((__m256)x.v, (__m256)y.v, (__m256)z.v));
}
inline batch<double, 4> fma(const batch<double, 4> &x, const batch<double, 4> &y, const batch<double, 4> &z)
{
    return batch<double, 4>((batch<double, 4>::vec_t)_mm256_fmadd_pd((__m256d)x.v, (__m256d)y.v, (__m256d)z.v));
}
inline batch<double, 4> fnma(const batch<double, 4> &x, const batch<double, 4> &y, const batch<double, 4> &z)
{
    return batch<double, 4>((batch<double, 4>::vec_t)_mm256_fnmadd_pd((__m256d)x.v, (__m256d)y.v, (__m256d)z.v));
}
inline batch<float, 8> sqrt(const batch<float, 8> &x) { return batch<float, 8>(_mm256_sqrt_ps((__m256)x.v)); }
inline batch<double, 4> sqrt(const batch<double, 4> &x)
{
    return batch<double, 4>((batch<double, 4>::vec_t)_mm256_sqrt_pd((__m256d)x.v));
}
#if defined(__AVX512F__)
inline batch<float, 16> fma(const batch<float, 16> &x, const batch<float, 16> &y, const batch<float, 16> &z)
{
    return batch<float, 16>(_mm512_fmadd_ps((__m512)x.v, (__m512)y.v, (__m512)z.v));
}
inline batch<float, 16> fnma(const batch<float, 16> &x, const batch<float, 16> &y, const batch<float, 16> &z)
{
    return batch<float, 16>(_mm512_fnmadd_ps((__m512)x.v, (__m512)y.v, (__m512)z.v));
}
inline batch<double, 8> fma(const batch<double, 8> &x, const batch<double, 8> &y, const batch<double, 8> &z)
{
    return batch<double, 8>((batch<double, 8>::vec_t)_mm512_fmadd_pd((__m512d)x.v, (__m512d)y.v, (__m512d)z.v));
}
inline batch<double, 8> fnma(const batch<double, 8> &x, const batch<double, 8> &y, const batch<double, 8> &z)
{
    return batch<double, 8>((batch<double, 8>::vec_t)_mm512_fnmadd_pd((__m512d)x.v, (__m512d)y.v, (__m512d)z.v));
}
inline batch<float, 16> sqrt(const batch<float, 16> &x) { return batch<float, 16>(_mm512_sqrt_ps((__m512)x.v)); }
inline batch<double, 8> sqrt(const batch<double, 8> &x)
{
    return batch<double, 8>((batch<double, 8>::vec_t)_mm512_sqrt_pd((__m512d)x.v));
}
#endif

// Horizontal sum, pairwise from the outside in (the order of xsimd's hadd for AVX: fold halves, then pairs).
template <typename T, std::size_t N>
inline T hadd(const batch<T, N> &b)
{
    T tmp[N];
    b.store_unaligned(tmp);
    for (std::size_t w = N / 2; w >= 1; w /= 2) {
        for (std::size_t i = 0; i < w; ++i) {
            tmp[i] += tmp[i + w];
        }
    }
    return tmp[0];
}

} // namespace xsimd

#endif
