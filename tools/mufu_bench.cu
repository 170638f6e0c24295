// Microbenchmark: MUFU.EX2 / F2FP issue cost per warp-instruction on one SM.  nvcc -arch=sm_100a -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, float a, int iters) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = a + i * 0.01f + threadIdx.x * 1e-4f;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0 || MODE == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                uint32_t r;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x[i]), "f"(x[i + 1]));
                acc ^= r;
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    const int iters = 2000;
    for (int mode = 0; mode < 3; ++mode)
        for (int warps : {1, 4, 8, 16}) {
            if (mode == 0) k<0><<<1, warps * 32>>>(out, cyc, 0.5f, iters);
            if (mode == 1) k<1><<<1, warps * 32>>>(out, cyc, 0.5f, iters);
            if (mode == 2) k<2><<<1, warps * 32>>>(out, cyc, 0.5f, iters);
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            const char* nm[] = {"MUFU.EX2 x16", "FFMA x16", "MUFU x16 + F2FP x8"};
            printf("%-20s warps/SM %2d: %.2f cycles per loop iteration (16 ops/thread) -> %.2f cyc/warp-instr/SMSP-share\n", nm[mode], warps,
                   (double)c / iters, (double)c / iters / 16.0);
        }
    return 0;
}
