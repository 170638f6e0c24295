// Microbenchmark: the half-row softmax inner loop (56 scores per thread: FFMA, MUFU.EX2, PRMT pack) with 1..4 warps per
// SM sub-partition, optional polynomial share on the FMA pipe.  One CTA on one SM; cycles per pass per warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_mt_bench exp_mt_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t packt(float lo, float hi) { return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632); }
__device__ __forceinline__ float ex2p(float x) {
    x = fmaxf(x, -126.0f);
    const float t = x + 12582912.0f;
    const float f = x - (t - 12582912.0f);
    float p = fmaf(f, 0.0551716685f, 0.242611125f);
    p = fmaf(f, p, 0.693260968f);
    p = fmaf(f, p, 0.999928057f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}
template <int NE, int POLY8, bool WITH_MAX>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, long long* cyc, const float* in, float sl2, float moff, int iters) {
    float s[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) s[i] = in[(threadIdx.x * NE + i) & 4095];
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        float mo = moff;
        if (WITH_MAX) {
            float m0 = s[0], m1 = s[1], m2 = s[2], m3 = s[3];
#pragma unroll
            for (int i = 4; i < NE; i += 4) { m0 = fmaxf(m0, s[i]); m1 = fmaxf(m1, s[i + 1]); m2 = fmaxf(m2, s[i + 2]); m3 = fmaxf(m3, s[i + 3]); }
            mo += fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * 1e-9f;
        }
        uint32_t pk[NE / 2];
#pragma unroll
        for (int i = 0; i < NE / 2; ++i) {
            const float a0 = fmaf(s[2 * i], sl2, -mo), a1 = fmaf(s[2 * i + 1], sl2, -mo);
            const bool poly = (i & 7) < POLY8;
            pk[i] = packt(poly ? ex2p(a0) : ex2a(a0), poly ? ex2p(a1) : ex2a(a1));
        }
#pragma unroll
        for (int i = 0; i < NE / 2; ++i) acc ^= pk[i];
        s[it & (NE - 1) & 7] += __uint_as_float(acc & 1u);   // loop-carried dependence so nothing is hoisted
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NE, int POLY8, bool WITH_MAX> void run(const char* name, uint32_t* out, long long* cyc, float* in) {
    const int iters = 500;
    for (int wps : {1, 2, 3, 4}) {
        k<NE, POLY8, WITH_MAX><<<1, wps * 128>>>(out, cyc, in, 0.17f, 3.0f, iters);
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-30s warps/SMSP %d: %7.1f cycles per pass (all warps together); XU floor %d; XU utilisation %.2f  (%s)\n", name, wps,
               (double)c / iters, wps * (NE - 2 * POLY8 * (NE / 16)) * 8, wps * (NE - 2.0 * POLY8 * (NE / 16)) * 8 / ((double)c / iters),
               cudaGetErrorString(cudaGetLastError()));
    }
}
int main() {
    uint32_t* out; float* in; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024); cudaMalloc(&in, 4096 * 4);
    cudaMemset(in, 0, 4096 * 4);
    run<64, 0, false>("64 elems, MUFU", out, cyc, in);
    run<64, 0, true>("64 elems, max + MUFU", out, cyc, in);
    run<64, 1, true>("64 elems, max + poly 1/8", out, cyc, in);
    run<64, 2, true>("64 elems, max + poly 2/8", out, cyc, in);
    run<64, 3, true>("64 elems, max + poly 3/8", out, cyc, in);
    run<64, 4, true>("64 elems, max + poly 4/8", out, cyc, in);
    return 0;
}
