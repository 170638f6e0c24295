// Microbenchmark: is MUFU.EX2 on bf16 / f16 operands cheaper than on f32?  nvcc -arch=sm_100a -o mufu_bf16_bench mufu_bf16_bench.cu
// (ex2.approx.ftz.bf16x2 / ex2.approx.f16x2 compile to two MUFU.EX2.BF16 / .F16 per packed register.)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, uint32_t seed, int iters) {
    uint32_t x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = seed + i * 0x00010001u + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[i]));
            if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[i]));
            if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    const int iters = 2000;
    const char* nm[] = {"ex2.f32 x16 (16 exps)", "ex2.bf16x2 x16 (32 exps)", "ex2.f16x2 x16 (32 exps)"};
    for (int mode = 0; mode < 3; ++mode)
        for (int warps : {4, 8}) {
            if (mode == 0) k<0><<<1, warps * 32>>>(out, cyc, 0x3f003f00u, iters);
            if (mode == 1) k<1><<<1, warps * 32>>>(out, cyc, 0x3f003f00u, iters);
            if (mode == 2) k<2><<<1, warps * 32>>>(out, cyc, 0x38003800u, iters);
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            const int exps = mode == 0 ? 16 : 32;
            printf("%-26s warps/SM %d: %.1f cycles per iteration -> %.2f cycles per exponential per warp (SMSP share)\n", nm[mode], warps,
                   (double)c / iters, (double)c / iters / exps / (warps / 4));
        }
    return 0;
}
