// sort.cu — hand-rolled stable LSD radix sort of (63-bit Morton code, u32 index) pairs for sm_100a.
//
// Replaces the reference's indirect tbb::parallel_sort (tree.hpp:1266-1274). The reference sort is
// unstable; the stable order is one of its legal outcomes and is the canonical one here (SURVEY §7,
// hard part 2). No CUB/Thrust.
//
// One sweep per 8-bit digit (n < 2^30): hist8 reads the keys ONCE and builds all eight digit histograms; each pass is
// then a single kernel that ranks a tile (warp-level MATCH.ANY), obtains the number of equal digits in all earlier
// tiles by DECOUPLED LOOK-BACK over per-tile status words (tile ids are handed out by an atomic counter, so a
// predecessor is always resident or done), stages the tile in digit order in shared memory and writes each digit
// run as one coalesced burst. 8 B read once + (12 B read + 12 B write) per element per pass, 10 launches and no
// host read-back (round 1: 3 kernels per pass re-reading the keys, 25 launches, one blocking read-back of a digit
// mask: 0.59 ms of the 1.14 ms build at 4 M particles).
// n >= 2^30 (status words keep 30 bits of prefix) uses the three-kernel passes: tile_hist (per-tile digit counts,
// digit-major) -> row_scan (one CTA per digit scans its row over tiles) -> scatter.

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


// ---- one-sweep passes ---------------------------------------------------------------------------------------------
#ifndef RK_SORT_IPT
#define RK_SORT_IPT 16
#endif
constexpr int OS_IPT = RK_SORT_IPT;
constexpr int OS_TILE = SORT_THREADS * OS_IPT;
constexpr int OS_SPAN = 32 * OS_IPT; // consecutive keys per warp
#ifndef RK_SORT_LB
#define RK_SORT_LB 8
#endif
constexpr int OS_LB = RK_SORT_LB; // predecessors inspected per look-back round trip
constexpr u32 OS_AGG = 1u << 30, OS_PREFIX = 2u << 30, OS_VALUE = (1u << 30) - 1u;

// All eight digit histograms in one pass over the keys. A digit that is equal over the whole warp (the top digits of
// Morton codes are nearly constant) is counted by one lane; otherwise every lane adds to its bin.
__global__ void __launch_bounds__(512) hist8_kernel(const u64 *__restrict__ keys, size_t n, u32 *__restrict__ ghist)
{
    __shared__ u32 h[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) {
        (&h[0][0])[i] = 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t base = blockIdx.x * size_t(blockDim.x) + (threadIdx.x & ~31u); base < n; base += stride) {
        const size_t i = base + lane;
        const bool valid = i < n;
        const u64 k = valid ? keys[i] : 0ull;
        const u32 vmask = __ballot_sync(0xffffffffu, valid);
        const int leader = __ffs(vmask) - 1;
        const u64 k0 = __shfl_sync(0xffffffffu, k, leader);
        const u64 diff = __reduce_or_sync(0xffffffffu, valid ? static_cast<u32>((k ^ k0) >> 32) : 0u);
        const u32 diff_lo = __reduce_or_sync(0xffffffffu, valid ? static_cast<u32>(k ^ k0) : 0u);
        const u64 d64 = (diff << 32) | diff_lo;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            if (((d64 >> (8 * p)) & 0xffull) == 0ull) {
                if (lane == leader) {
                    atomicAdd(&h[p][(k >> (8 * p)) & 0xffu], __popc(vmask));
                }
            } else if (valid) {
                atomicAdd(&h[p][(k >> (8 * p)) & 0xffu], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) {
        const u32 v = (&h[0][0])[i];
        if (v) {
            atomicAdd(&ghist[i], v);
        }
    }
}

// One pass: stable partition of the (key, idx) pairs by the digit at `shift`. status: one word per (tile, digit),
// zeroed before the sort: bits 31:30 = 0 nothing yet, OS_AGG = count of this tile, OS_PREFIX = count of this and all
// earlier tiles.
#ifndef RK_SORT_MINB
#define RK_SORT_MINB 2
#endif
__global__ void __launch_bounds__(SORT_THREADS, RK_SORT_MINB)
    onesweep_kernel(const u64 *__restrict__ keys_in, const u32 *__restrict__ idx_in, u64 *__restrict__ keys_out,
                    u32 *__restrict__ idx_out, size_t n, int shift, const u32 *__restrict__ ghist_pass,
                    u32 *__restrict__ status, u32 *__restrict__ tile_counter)
{
    extern __shared__ __align__(16) unsigned char sort_smem[];
    u64 *skey = reinterpret_cast<u64 *>(sort_smem);                     // OS_TILE keys in local digit order
    u32 *sidx = reinterpret_cast<u32 *>(skey + OS_TILE);                // OS_TILE payloads
    u32(*wcnt)[257] = reinterpret_cast<u32(*)[257]>(sidx + OS_TILE);    // per-warp digit counts -> local bases
    u32 *gdelta = &wcnt[0][0] + SORT_WARPS * 257;                       // 256: global base - local start
    u32 *ws = gdelta + 256;                                             // 8 + the tile id
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) {
        ws[8] = atomicAdd(tile_counter, 1u); // tile ids in launch order: a predecessor is resident or finished
    }
    for (int i = tid; i < SORT_WARPS * 257; i += SORT_THREADS) {
        (&wcnt[0][0])[i] = 0;
    }
    u32 dummy;
    const u32 dbase = block_exscan_256(ghist_pass[tid], ws, &dummy); // global base of digit `tid`; syncs
    const u32 tile = ws[8];

    const size_t tile_base = size_t(tile) * OS_TILE;
    const size_t wbase = tile_base + size_t(w) * OS_SPAN;
    u64 key[OS_IPT];
    u32 val[OS_IPT];
    unsigned short rank[OS_IPT];
    const u32 lt = lanemask_lt();
#pragma unroll
    for (int j = 0; j < OS_IPT; ++j) { // every load of the tile in flight before the first MATCH
        const size_t i = wbase + size_t(j) * 32 + lane;
        const bool valid = i < n;
        key[j] = valid ? keys_in[i] : ~0ull;
        val[j] = (valid && idx_in) ? idx_in[i] : static_cast<u32>(i);
    }
#pragma unroll
    for (int j = 0; j < OS_IPT; ++j) {
        const size_t i = wbase + size_t(j) * 32 + lane;
        const u32 d = (i < n) ? static_cast<u32>((key[j] >> shift) & 0xffu) : 256u;
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
    {
        u32 tot = 0;
#pragma unroll
        for (int k = 0; k < SORT_WARPS; ++k) {
            tot += wcnt[k][tid];
        }
        // publish this tile's count of digit `tid`, then sum the earlier tiles' counts by looking back
        volatile u32 *st = status + size_t(tile) * 256 + tid;
        u32 excl = 0;
        if (tile == 0) {
            *st = tot | OS_PREFIX;
        } else {
            *st = tot | OS_AGG;
            // Look back OS_LB tiles per round trip: the status words of the nearest OS_LB predecessors are loaded
            // together (independent L2 accesses), then consumed nearest first; only a word that is not published yet
            // is polled. One dependent load per predecessor made the pass latency-bound (most of the ~300 resident
            // tiles have only published their own count when a tile looks back).
            const volatile u32 *pv = st - 256;
            u32 remaining = tile, polls = 0;
            bool found = false;
            while (!found) {
                u32 v[OS_LB];
#pragma unroll
                for (int j = 0; j < OS_LB; ++j) {
                    v[j] = static_cast<u32>(j) < remaining ? pv[-256 * j] : OS_PREFIX; // before tile 0: prefix 0
                }
#pragma unroll
                for (int j = 0; j < OS_LB; ++j) {
                    if (!found) {
                        while ((v[j] & ~OS_VALUE) == 0u) {
                            v[j] = pv[-256 * j];
                            if (++polls == (1u << 26)) {
                                __trap(); // a predecessor never published: fail loudly instead of hanging the device
                            }
                        }
                        excl += v[j] & OS_VALUE;
                        found = (v[j] & OS_PREFIX) != 0u;
                    }
                }
                pv -= 256 * OS_LB;
                remaining = remaining > u32(OS_LB) ? remaining - OS_LB : 0u;
            }
            *st = (excl + tot) | OS_PREFIX;
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
        gdelta[tid] = dbase + excl - lstart;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < OS_IPT; ++j) {
        const size_t i = wbase + size_t(j) * 32 + lane;
        if (i < n) {
            const u32 d = static_cast<u32>((key[j] >> shift) & 0xffu);
            const u32 lpos = wcnt[w][d] + rank[j];
            skey[lpos] = key[j];
            sidx[lpos] = val[j];
        }
    }
    __syncthreads();
    const u32 nvalid = (tile_base + OS_TILE <= n) ? u32(OS_TILE) : static_cast<u32>(n - tile_base);
    for (u32 e = tid; e < nvalid; e += SORT_THREADS) {
        const u64 k = skey[e];
        const u32 pos = gdelta[(k >> shift) & 0xffu] + e;
        keys_out[pos] = k;
        idx_out[pos] = sidx[e];
    }
}

constexpr size_t ONESWEEP_SMEM = size_t(OS_TILE) * 12 + size_t(SORT_WARPS) * 257 * 4 + 256 * 4 + 16 * 4;

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

// Stable partition of the indices 0 .. n-1 by an 8-bit bucket id held in the low byte of keys_in: ONE one-sweep pass.
// idx_out receives the permutation, keys_out the permuted ids; counts_dev (256 u32, device) the bucket sizes.
void radix_partition_pass(const u64 *keys_in, u64 *keys_out, u32 *idx_out, size_t n, sort_scratch &sc, cudaStream_t st,
                          const u32 **counts_dev)
{
    const u32 nt = div_up(n, OS_TILE);
    sc.ghist.reserve(8 * 256 + 8);
    sc.tilehist.reserve(size_t(8) * 256 * (nt ? nt : 1), 1.25);
    RK_CUDA_CHECK(cudaMemsetAsync(sc.ghist.p, 0, (8 * 256 + 8) * sizeof(u32), st));
    *counts_dev = sc.ghist.p;
    if (!n || n >= (size_t(1) << 30)) {
        if (n) {
            throw cuda_error(3, "radix_partition_pass: too many elements");
        }
        return;
    }
    RK_CUDA_CHECK(cudaMemsetAsync(sc.tilehist.p, 0, size_t(256) * nt * sizeof(u32), st));
    hist8_kernel<<<148 * 2, 512, 0, st>>>(keys_in, n, sc.ghist.p); count_launch();
    RK_CUDA_CHECK(cudaFuncSetAttribute(onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(ONESWEEP_SMEM)));
    onesweep_kernel<<<nt, SORT_THREADS, ONESWEEP_SMEM, st>>>(keys_in, nullptr, keys_out, idx_out, n, 0, sc.ghist.p,
                                                            sc.tilehist.p, sc.ghist.p + 8 * 256); count_launch();
    RK_CUDA_CHECK(cudaGetLastError());
}

int radix_sort_pairs(u64 *keys_a, u64 *keys_b, u32 *idx_a, u32 *idx_b, size_t n, sort_scratch &sc, cudaStream_t st,
                     u64 **keys_out, u32 **idx_out)
{
    *keys_out = keys_a;
    *idx_out = idx_a;
    if (n == 0) {
        return 0;
    }
    if (n < (size_t(1) << 30)) {
        // one sweep per digit, no host read-back
        const u32 nt = div_up(n, OS_TILE);
        sc.ghist.reserve(8 * 256 + 8);
        sc.tilehist.reserve(size_t(8) * 256 * nt, 1.25);
        RK_CUDA_CHECK(cudaMemsetAsync(sc.ghist.p, 0, (8 * 256 + 8) * sizeof(u32), st));
        RK_CUDA_CHECK(cudaMemsetAsync(sc.tilehist.p, 0, size_t(8) * 256 * nt * sizeof(u32), st));
        hist8_kernel<<<148 * 2, 512, 0, st>>>(keys_a, n, sc.ghist.p); count_launch();
        RK_CUDA_CHECK(cudaFuncSetAttribute(onesweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(ONESWEEP_SMEM)));
        u64 *kin = keys_a, *kout = keys_b;
        u32 *iin = nullptr, *iout = idx_b; // first pass: implicit iota, written into idx_b
        for (int pass = 0; pass < 8; ++pass) {
            onesweep_kernel<<<nt, SORT_THREADS, ONESWEEP_SMEM, st>>>(kin, iin, kout, iout, n, 8 * pass,
                                                                    sc.ghist.p + 256 * pass,
                                                                    sc.tilehist.p + size_t(pass) * 256 * nt,
                                                                    sc.ghist.p + 8 * 256 + pass); count_launch();
            u64 *tk = kin;
            kin = kout;
            kout = tk;
            iin = iout;
            iout = (iin == idx_a) ? idx_b : idx_a;
        }
        *keys_out = kin; // 8 swaps: back in keys_a / idx_a
        *idx_out = iin;
        RK_CUDA_CHECK(cudaGetLastError());
        return 8;
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
