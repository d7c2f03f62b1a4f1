// Micro-benchmark: packed FP32 (FFMA2/FADD2/FMUL2, fma.rn.f32x2) against scalar FFMA on sm_100a.
// Question: does a packed instruction free issue slots (one issue, two FMAs) and what is its pipe throughput?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ffma2 tools/microbench/ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b, unsigned m)
{
    float x[8];
    u64 p[8];
    unsigned q[8];
    for (int i = 0; i < 8; ++i) {
        x[i] = threadIdx.x * 1e-3f + i;
        p[i] = (u64(__float_as_uint(x[i])) << 32) | __float_as_uint(x[i] + 0.5f);
        q[i] = threadIdx.x + i;
    }
    const u64 a2 = (u64(__float_as_uint(a)) << 32) | __float_as_uint(a), b2 = (u64(__float_as_uint(b)) << 32) | __float_as_uint(b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 2) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
            }
            if (MODE == 1 || MODE == 3) {
                p[i] = fma2(p[i], a2, b2);
            }
            if (MODE == 2 || MODE == 3) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(m), "r"(it)); // ALU pipe
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) {
        s += x[i] + __uint_as_float(unsigned(p[i])) + __uint_as_float(unsigned(p[i] >> 32)) + q[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char *name, int fma_per_inst, float *d)
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 20000, grid = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(d, 100, 1.0001f, 0.5f, 3);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(d, iters, 1.0001f, 0.5f, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double inst = double(grid) * 256 / 32 * iters * 8; // warp-level FP instructions
    printf("%-28s %8.3f ms  %7.2f TFLOP/s  %6.3f FP warp-inst/clk/SM (at 1.92 GHz)\n", name, ms,
           inst * 32 * fma_per_inst * 2 / ms / 1e9, inst / (ms * 1e-3 * 1.92e9) / sms);
}
int main()
{
    float *d;
    cudaMalloc(&d, 148 * 8 * 256 * 4 * 4);
    run<0>("FFMA", 1, d);
    run<1>("FFMA2", 2, d);
    run<2>("FFMA + LOP3 (1:1)", 1, d);
    run<3>("FFMA2 + LOP3 (1:1)", 2, d);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
