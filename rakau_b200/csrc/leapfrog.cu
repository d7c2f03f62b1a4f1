// leapfrog.cu — device-resident kick-drift-kick integrator around the tree (sm_100a; HBM-bound streaming kernels).
//
// Counterpart of the time loop of the reference's benchmark/benchmark_leapfrog.cpp:286-384, with the same operations in
// the same order: kicked velocities (349-356), update_particles_u functor (359-370: the drift), new accelerations,
// velocity update through last_perm (375-383), and the conserved quantities of track_integrals (292-347). Positions and
// velocities never leave the GPU: the drift writes the moved particles straight into the pre-sort array of the rebuild
// (it replaces the copy-in of rk_tree_update_positions), velocities are kept in the tree's internal order.
#include "common.cuh"

namespace rk
{

namespace
{

__device__ __forceinline__ u64 lf_abs_bits(float v) { return __float_as_uint(v) & 0x7fffffffu; }
__device__ __forceinline__ u64 lf_abs_bits(double v)
{
    return static_cast<u64>(__double_as_longlong(v)) & 0x7fffffffffffffffull;
}

// vec[i] = vec_in[perm[i]] for three arrays: the `reorder` helper, benchmark_leapfrog.cpp:252-267.
template <typename F>
__global__ void __launch_bounds__(256)
    lf_reorder_kernel(const F *__restrict__ ax, const F *__restrict__ ay, const F *__restrict__ az,
                      const u32 *__restrict__ perm, F *__restrict__ ox, F *__restrict__ oy, F *__restrict__ oz, size_t n)
{
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i < n) {
        const u32 j = perm[i];
        ox[i] = ax[j];
        oy[i] = ay[j];
        oz[i] = az[j];
    }
}

// kick: kv = fma(acc, dt/2, v) (349-356); drift: x = fma(kv, dt, x) (359-370). The moved particle goes to `pin`, the
// pre-sort array of the rebuild, together with the max |coordinate| the box deduction needs.
template <typename F>
__global__ void __launch_bounds__(256)
    lf_kick_drift_kernel(const F *__restrict__ ax, const F *__restrict__ ay, const F *__restrict__ az,
                         const F *__restrict__ vx, const F *__restrict__ vy, const F *__restrict__ vz,
                         const vec4<F> *__restrict__ pos, F half_dt, F dt, F *__restrict__ kx, F *__restrict__ ky,
                         F *__restrict__ kz, vec4<F> *__restrict__ pin, size_t n, u64 *__restrict__ absmax)
{
    u64 mx = 0;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const F k0 = rn_fma(ax[i], half_dt, vx[i]), k1 = rn_fma(ay[i], half_dt, vy[i]), k2 = rn_fma(az[i], half_dt, vz[i]);
        kx[i] = k0;
        ky[i] = k1;
        kz[i] = k2;
        vec4<F> p = pos[i];
        p.x = rn_fma(k0, dt, p.x);
        p.y = rn_fma(k1, dt, p.y);
        p.z = rn_fma(k2, dt, p.z);
        pin[i] = p;
        const u64 ba = lf_abs_bits(p.x), bb = lf_abs_bits(p.y), bc = lf_abs_bits(p.z);
        mx = ba > mx ? ba : mx;
        mx = bb > mx ? bb : mx;
        mx = bc > mx ? bc : mx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const u64 t = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = t > mx ? t : mx;
    }
    if ((threadIdx.x & 31) == 0 && mx) {
        atomicMax(absmax, mx);
    }
}

// v[i] = fma(acc[i], dt/2, kv[last_perm[i]]): 375-383 (the kicked velocities are still in the previous order).
template <typename F>
__global__ void __launch_bounds__(256)
    lf_kick_reindex_kernel(const F *__restrict__ ax, const F *__restrict__ ay, const F *__restrict__ az,
                           const F *__restrict__ kx, const F *__restrict__ ky, const F *__restrict__ kz,
                           const u32 *__restrict__ last_perm, F half_dt, F *__restrict__ vx, F *__restrict__ vy,
                           F *__restrict__ vz, size_t n)
{
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i < n) {
        const u32 j = last_perm[i];
        vx[i] = rn_fma(ax[i], half_dt, kx[j]);
        vy[i] = rn_fma(ay[i], half_dt, ky[j]);
        vz[i] = rn_fma(az[i], half_dt, kz[j]);
    }
}

// track_integrals, 292-347: sums of positions, velocities and of m v^2 / 2 + pot, in double. Two deterministic stages:
// per-CTA partial sums, then one CTA adds them in a fixed order.
constexpr int LF_RED_BLOCKS = 1184;
template <typename F>
__global__ void __launch_bounds__(256)
    lf_integrals_kernel(const vec4<F> *__restrict__ pos, const F *__restrict__ vx, const F *__restrict__ vy,
                        const F *__restrict__ vz, const F *__restrict__ pot, size_t n, double *__restrict__ partial)
{
    __shared__ double red[7][8];
    double s[7] = {0, 0, 0, 0, 0, 0, 0};
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const vec4<F> p = pos[i];
        const F a = vx[i], b = vy[i], c = vz[i];
        s[0] += p.x;
        s[1] += p.y;
        s[2] += p.z;
        s[3] += a;
        s[4] += b;
        s[5] += c;
        const F v2 = a * a + b * b + c * c;
        s[6] += double(F(0.5) * p.w * v2 + pot[i]); // (1/2) m v^2 + pots[i], 339-340
    }
#pragma unroll
    for (int q = 0; q < 7; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
        }
        if ((threadIdx.x & 31) == 0) {
            red[q][threadIdx.x >> 5] = s[q];
        }
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double t = 0;
        for (int w = 0; w < 8; ++w) {
            t += red[threadIdx.x][w];
        }
        partial[size_t(blockIdx.x) * 8 + threadIdx.x] = t;
    }
}
__global__ void lf_integrals_final_kernel(const double *__restrict__ partial, int nblocks, double *__restrict__ out)
{
    if (threadIdx.x < 7) {
        double t = 0;
        for (int b = 0; b < nblocks; ++b) {
            t += partial[size_t(b) * 8 + threadIdx.x];
        }
        out[threadIdx.x] = t;
    }
}

} // namespace

template <typename F>
void launch_lf_reorder(const F *const in[3], const u32 *perm, F *const out[3], size_t n, cudaStream_t st)
{
    if (n) {
        lf_reorder_kernel<F><<<div_up(n, 256), 256, 0, st>>>(in[0], in[1], in[2], perm, out[0], out[1], out[2], n); count_launch();
    }
}
template <typename F>
void launch_lf_kick_drift(const F *const acc[3], const F *const v[3], const vec4<F> *pos, F half_dt, F dt, F *const kv[3],
                          vec4<F> *pin, size_t n, u64 *absmax, cudaStream_t st)
{
    if (n) {
        lf_kick_drift_kernel<F><<<148 * 8, 256, 0, st>>>(acc[0], acc[1], acc[2], v[0], v[1], v[2], pos, half_dt, dt, kv[0],
                                                         kv[1], kv[2], pin, n, absmax); count_launch();
    }
}
template <typename F>
void launch_lf_kick_reindex(const F *const acc[3], const F *const kv[3], const u32 *last_perm, F half_dt, F *const v[3],
                            size_t n, cudaStream_t st)
{
    if (n) {
        lf_kick_reindex_kernel<F><<<div_up(n, 256), 256, 0, st>>>(acc[0], acc[1], acc[2], kv[0], kv[1], kv[2], last_perm,
                                                                  half_dt, v[0], v[1], v[2], n); count_launch();
    }
}
template <typename F>
void launch_lf_integrals(const vec4<F> *pos, const F *const v[3], const F *pot, size_t n, double *scratch /* >= 8 * 1185 */,
                         cudaStream_t st)
{
    if (n) {
        lf_integrals_kernel<F><<<LF_RED_BLOCKS, 256, 0, st>>>(pos, v[0], v[1], v[2], pot, n, scratch + 8); count_launch();
        lf_integrals_final_kernel<<<1, 32, 0, st>>>(scratch + 8, LF_RED_BLOCKS, scratch); count_launch();
    }
}
unsigned lf_scratch_doubles() { return 8u * (LF_RED_BLOCKS + 1); }

#define RK_LF_INST(F)                                                                                                   \
    template void launch_lf_reorder<F>(const F *const[3], const u32 *, F *const[3], size_t, cudaStream_t);              \
    template void launch_lf_kick_drift<F>(const F *const[3], const F *const[3], const vec4<F> *, F, F, F *const[3],     \
                                          vec4<F> *, size_t, u64 *, cudaStream_t);                                      \
    template void launch_lf_kick_reindex<F>(const F *const[3], const F *const[3], const u32 *, F, F *const[3], size_t,   \
                                            cudaStream_t);                                                              \
    template void launch_lf_integrals<F>(const vec4<F> *, const F *const[3], const F *, size_t, double *, cudaStream_t);
RK_LF_INST(float)
RK_LF_INST(double)

} // namespace rk
