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

} // namespace rk
