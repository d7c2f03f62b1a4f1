// scan.cuh — small block/row scan helpers shared by sort.cu and build.cu.
#pragma once

#include "common.cuh"

namespace rk
{

__device__ __forceinline__ u32 lanemask_lt()
{
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Block-wide exclusive scan of one u32 per thread for a 256-thread CTA. Returns the exclusive prefix;
// *total receives the CTA sum. warp_sums needs >= 8 entries of shared memory. Contains two barriers.
__device__ __forceinline__ u32 block_exscan_256(u32 v, u32 *warp_sums, u32 *total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) {
            incl += t;
        }
    }
    if (lane == 31) {
        warp_sums[w] = incl;
    }
    __syncthreads();
    u32 wpre = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const u32 s = warp_sums[k];
        if (k < w) {
            wpre += s;
        }
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return wpre + incl - v;
}

// One 256-thread CTA per row: exclusive scan of rows[blockIdx.x][0..ncols) in place; the row sum goes to
// totals[blockIdx.x].
static __global__ void __launch_bounds__(256) row_scan_kernel(u32 *__restrict__ rows, u32 ncols, u32 *__restrict__ totals)
{
    __shared__ u32 ws[8];
    u32 *row = rows + size_t(blockIdx.x) * ncols;
    u32 carry = 0;
    for (u32 c = 0; c < ncols; c += 256) {
        const u32 i = c + threadIdx.x;
        const u32 v = i < ncols ? row[i] : 0u;
        u32 tot;
        const u32 ex = block_exscan_256(v, ws, &tot);
        if (i < ncols) {
            row[i] = carry + ex;
        }
        carry += tot;
    }
    if (threadIdx.x == 0) {
        totals[blockIdx.x] = carry;
    }
}

// The same with 1024 threads and four consecutive columns per thread (4096 columns per round): the per-level tile
// counts of the topology are 23 rows of N / 256 columns, which the 256-column rounds above scan in 61 dependent rounds
// at 4 M particles (42 us of a nearly idle GPU).
static __global__ void __launch_bounds__(1024) row_scan_wide_kernel(u32 *__restrict__ rows, u32 ncols, u32 *__restrict__ totals)
{
    __shared__ u32 ws[32];
    u32 *row = rows + size_t(blockIdx.x) * ncols;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 carry = 0;
    for (u32 c = 0; c < ncols; c += 4096) {
        const u32 i0 = c + 4u * threadIdx.x;
        u32 v[4], sum = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = i0 + j < ncols ? row[i0 + j] : 0u;
            sum += v[j];
        }
        u32 incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) {
                incl += t;
            }
        }
        if (lane == 31) {
            ws[w] = incl;
        }
        __syncthreads();
        u32 wpre = 0, tot = 0;
        {
            const u32 sv = ws[lane]; // 32 warps: one warp-level scan of the warp sums, done by every warp
            u32 si = sv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 t = __shfl_up_sync(0xffffffffu, si, o);
                if (lane >= o) {
                    si += t;
                }
            }
            wpre = __shfl_sync(0xffffffffu, si - sv, w);
            tot = __shfl_sync(0xffffffffu, si, 31);
        }
        __syncthreads();
        u32 run = carry + wpre + incl - sum;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i0 + j < ncols) {
                row[i0 + j] = run;
            }
            run += v[j];
        }
        carry += tot;
    }
    if (threadIdx.x == 0) {
        totals[blockIdx.x] = carry;
    }
}

} // namespace rk
