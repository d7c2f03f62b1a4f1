// sort.cu — hand-rolled stable LSD radix sort of (63-bit Morton code, u32 index) pairs for sm_100a.
//
// Replaces the reference's indirect tbb::parallel_sort (tree.hpp:1266-1274). The reference sort is
// unstable; the stable order is one of its legal outcomes and is the canonical one here (SURVEY §7,
// hard part 2). No CUB/Thrust.
//
// Per 8-bit pass: tile_hist (per-tile digit counts, digit-major) -> row_scan (one CTA per digit scans its
// row over tiles) -> scatter (re-reads the tile, warp-level MATCH.ANY ranking, stable scatter).
// Passes whose digit is constant over all keys are skipped (found with one OR-reduction over key^key0):
// on a Plummer sphere the deduced box is ~1e3 core radii, so the top digits are nearly constant.
// All kernels are HBM-bound: 8 B key read (hist) + 12 B read + 12 B write (scatter) per element per pass.

#include "common.cuh"
#include "scan.cuh"

namespace rk
{

namespace
{

constexpr int SORT_THREADS = 256;
constexpr int SORT_IPT = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_IPT; // 2048 keys per CTA (small tiles: latency-bound kernel, occupancy matters)
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int WARP_SPAN = 32 * SORT_IPT; // 512 consecutive keys per warp

// OR over (key ^ key[0]): a digit with no bit set in the result is constant over the whole input.
__global__ void __launch_bounds__(256) key_or_kernel(const u64 *__restrict__ keys, size_t n, u64 *__restrict__ out)
{
    const u64 k0 = keys[0];
    u64 acc = 0;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        acc |= keys[i] ^ k0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc |= __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if ((threadIdx.x & 31) == 0 && acc) {
        atomicOr(out, acc);
    }
}

__global__ void __launch_bounds__(SORT_THREADS)
    tile_hist_kernel(const u64 *__restrict__ keys, size_t n, int shift, u32 ntiles, u32 *__restrict__ tilehist)
{
    __shared__ u32 hist[257];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < 257; i += SORT_THREADS) {
        hist[i] = 0;
    }
    __syncthreads();
    const size_t wbase = size_t(blockIdx.x) * SORT_TILE + size_t(w) * WARP_SPAN;
    u32 dg[SORT_IPT];
#pragma unroll
    for (int j = 0; j < SORT_IPT; ++j) { // all loads in flight before the first MATCH
        const size_t i = wbase + size_t(j) * 32 + lane;
        dg[j] = (i < n) ? static_cast<u32>((keys[i] >> shift) & 0xffu) : 256u;
    }
#pragma unroll
    for (int j = 0; j < SORT_IPT; ++j) {
        const u32 peers = __match_any_sync(0xffffffffu, dg[j]);
        if (lane == __ffs(peers) - 1) {
            atomicAdd(&hist[dg[j]], __popc(peers));
        }
    }
    __syncthreads();
    if (tid < 256) {
        tilehist[size_t(tid) * ntiles + blockIdx.x] = hist[tid];
    }
}

// Scatter with shared-memory staging: the tile is first put in digit order in shared memory (stable local
// rank), then written out with consecutive threads covering consecutive destinations, so each digit's run is
// one coalesced burst instead of 4096 scattered 8-byte stores (ncu: the direct-scatter version ran at 4x its
// algorithmic-bytes time).
__global__ void __launch_bounds__(SORT_THREADS, 4)
    scatter_kernel(const u64 *__restrict__ keys_in, const u32 *__restrict__ idx_in, u64 *__restrict__ keys_out,
                   u32 *__restrict__ idx_out, size_t n, int shift, u32 ntiles, const u32 *__restrict__ tilehist,
                   const u32 *__restrict__ totals)
{
    extern __shared__ __align__(16) unsigned char sort_smem[];
    u64 *skey = reinterpret_cast<u64 *>(sort_smem);                      // SORT_TILE keys in local digit order
    u32 *sidx = reinterpret_cast<u32 *>(skey + SORT_TILE);               // SORT_TILE payloads
    u32(*wcnt)[257] = reinterpret_cast<u32(*)[257]>(sidx + SORT_TILE);   // per-warp digit counts -> local bases
    u32 *gdelta = &wcnt[0][0] + SORT_WARPS * 257;                        // 256: global base - local start
    u32 *ws = gdelta + 256;                                              // 8
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < SORT_WARPS * 257; i += SORT_THREADS) {
        (&wcnt[0][0])[i] = 0;
    }
    u32 dummy;
    const u32 dbase = block_exscan_256(totals[tid], ws, &dummy); // global base of digit `tid`; syncs

    const size_t tile_base = size_t(blockIdx.x) * SORT_TILE;
    const size_t wbase = tile_base + size_t(w) * WARP_SPAN;
    u64 key[SORT_IPT];
    u32 val[SORT_IPT];
    unsigned short rank[SORT_IPT];
    const u32 lt = lanemask_lt();
    // issue every load of the tile up front: the kernel is latency-bound, not bandwidth-bound
#pragma unroll
    for (int j = 0; j < SORT_IPT; ++j) {
        const size_t i = wbase + size_t(j) * 32 + lane;
        const bool valid = i < n;
        key[j] = valid ? keys_in[i] : ~0ull;
        val[j] = (valid && idx_in) ? idx_in[i] : static_cast<u32>(i);
    }
#pragma unroll
    for (int j = 0; j < SORT_IPT; ++j) {
        const size_t i = wbase + size_t(j) * 32 + lane;
        const bool valid = i < n;
        const u32 d = valid ? static_cast<u32>((key[j] >> shift) & 0xffu) : 256u;
        const u32 peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        u32 old = 0;
        if (lane == leader) {
            old = wcnt[w][d];
            wcnt[w][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[j] = static_cast<unsigned short>(old + __popc(peers & lt));
        __syncwarp();
    }
    __syncthreads();
    // Per digit: total in the tile -> exclusive scan over digits = local start; per-warp counts -> local bases.
    {
        u32 tot = 0;
#pragma unroll
        for (int k = 0; k < SORT_WARPS; ++k) {
            tot += wcnt[k][tid];
        }
        u32 dummy2;
        const u32 lstart = block_exscan_256(tot, ws, &dummy2); // local start of digit `tid`
        u32 run = lstart;
#pragma unroll
        for (int k = 0; k < SORT_WARPS; ++k) {
            const u32 c = wcnt[k][tid];
            wcnt[k][tid] = run;
            run += c;
        }
        gdelta[tid] = dbase + tilehist[size_t(tid) * ntiles + blockIdx.x] - lstart;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SORT_IPT; ++j) {
        const size_t i = wbase + size_t(j) * 32 + lane;
        if (i < n) {
            const u32 d = static_cast<u32>((key[j] >> shift) & 0xffu);
            const u32 lpos = wcnt[w][d] + rank[j];
            skey[lpos] = key[j];
            sidx[lpos] = val[j];
        }
    }
    __syncthreads();
    const u32 nvalid = (tile_base + SORT_TILE <= n) ? u32(SORT_TILE) : static_cast<u32>(n - tile_base);
    for (u32 e = tid; e < nvalid; e += SORT_THREADS) {
        const u64 k = skey[e];
        const u32 pos = gdelta[(k >> shift) & 0xffu] + e;
        keys_out[pos] = k;
        idx_out[pos] = sidx[e];
    }
}

constexpr size_t SCATTER_SMEM = size_t(SORT_TILE) * 12 + size_t(SORT_WARPS) * 257 * 4 + 256 * 4 + 8 * 4;

__global__ void iota_kernel(u32 *p, size_t n)
{
    const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (i < n) {
        p[i] = static_cast<u32>(i);
    }
}

} // namespace

void launch_iota(u32 *p, size_t n, cudaStream_t st)
{
    if (n) {
        iota_kernel<<<div_up(n, 256), 256, 0, st>>>(p, n); count_launch();
    }
}

int radix_sort_pairs(u64 *keys_a, u64 *keys_b, u32 *idx_a, u32 *idx_b, size_t n, sort_scratch &sc, cudaStream_t st,
                     u64 **keys_out, u32 **idx_out)
{
    *keys_out = keys_a;
    *idx_out = idx_a;
    if (n == 0) {
        return 0;
    }
    const u32 ntiles = div_up(n, SORT_TILE);
    sc.ghist.reserve(256 + 8);
    sc.tilehist.reserve(size_t(256) * ntiles, 1.25);
    if (!sc.h_ghist) {
        RK_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void **>(&sc.h_ghist), 16 * sizeof(u32)));
    }
    u64 *d_or = reinterpret_cast<u64 *>(sc.ghist.p + 256); // 8-byte aligned: cudaMalloc base + 1024 B
    RK_CUDA_CHECK(cudaMemsetAsync(d_or, 0, sizeof(u64), st));
    key_or_kernel<<<592, 256, 0, st>>>(keys_a, n, d_or); count_launch();
    RK_CUDA_CHECK(cudaMemcpyAsync(sc.h_ghist, d_or, sizeof(u64), cudaMemcpyDeviceToHost, st));
    RK_CUDA_CHECK(cudaStreamSynchronize(st));
    u64 varying;
    static_assert(sizeof(u64) == 2 * sizeof(u32), "");
    varying = *reinterpret_cast<u64 *>(sc.h_ghist);

    RK_CUDA_CHECK(cudaFuncSetAttribute(scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(SCATTER_SMEM)));
    u64 *kin = keys_a, *kout = keys_b;
    u32 *iin = nullptr, *iout = idx_b; // first pass: implicit iota, write into idx_b
    u32 *ibufs[2] = {idx_a, idx_b};
    int passes = 0;
    for (int shift = 0; shift < 63; shift += 8) {
        if (((varying >> shift) & 0xffull) == 0) {
            continue;
        }
        tile_hist_kernel<<<ntiles, SORT_THREADS, 0, st>>>(kin, n, shift, ntiles, sc.tilehist.p); count_launch();
        row_scan_kernel<<<256, 256, 0, st>>>(sc.tilehist.p, ntiles, sc.ghist.p); count_launch();
        scatter_kernel<<<ntiles, SORT_THREADS, SCATTER_SMEM, st>>>(kin, iin, kout, iout, n, shift, ntiles,
                                                                   sc.tilehist.p, sc.ghist.p); count_launch();
        ++passes;
        u64 *tk = kin;
        kin = kout;
        kout = tk;
        iin = iout;
        iout = (iin == ibufs[0]) ? ibufs[1] : ibufs[0];
    }
    if (passes == 0) {
        // All keys equal: identity permutation.
        launch_iota(idx_a, n, st);
        *keys_out = keys_a;
        *idx_out = idx_a;
    } else {
        *keys_out = kin;
        *idx_out = iin;
    }
    RK_CUDA_CHECK(cudaGetLastError());
    return passes;
}

} // namespace rk
