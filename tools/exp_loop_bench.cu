// Microbenchmark of the attention softmax inner loop (one thread = one row of 128 scores) on a single SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exp_loop_bench exp_loop_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t packt(float lo, float hi) { return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632); }
__device__ __forceinline__ uint32_t packr(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

// MODE bits: 1 = FFMA, 2 = MUFU, 4 = pack (trunc), 8 = pack (cvt.rn), 16 = STS, 32 = max pass (FMNMX) first
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, const float* in, float sl2, float moff, int iters) {
    extern __shared__ uint8_t smem[];
    const int r = threadIdx.x & 127;
    const uint32_t pbase = (uint32_t)__cvta_generic_to_shared(smem) + (threadIdx.x >> 7) * 32768 + r * 128;
    const uint32_t rsw = r & 7;
    float s[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) s[i] = in[(threadIdx.x * 128 + i) & 4095];
    float accum = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        float mx = 0.f;
        if (MODE & 32) {
            float m0 = s[0], m1 = s[1], m2 = s[2], m3 = s[3];
#pragma unroll
            for (int i = 4; i < 128; i += 4) { m0 = fmaxf(m0, s[i]); m1 = fmaxf(m1, s[i + 1]); m2 = fmaxf(m2, s[i + 2]); m3 = fmaxf(m3, s[i + 3]); }
            mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * 1e-9f;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float pe[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float x = s[c * 32 + i];
                if (MODE & 1) x = fmaf(x, sl2, -moff - mx);
                if (MODE & 2) x = ex2a(x);
                pe[i] = x;
            }
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t a0, a1, a2, a3;
                if (MODE & 4) { a0 = packt(pe[ch*8], pe[ch*8+1]); a1 = packt(pe[ch*8+2], pe[ch*8+3]); a2 = packt(pe[ch*8+4], pe[ch*8+5]); a3 = packt(pe[ch*8+6], pe[ch*8+7]); }
                else if (MODE & 8) { a0 = packr(pe[ch*8], pe[ch*8+1]); a1 = packr(pe[ch*8+2], pe[ch*8+3]); a2 = packr(pe[ch*8+4], pe[ch*8+5]); a3 = packr(pe[ch*8+6], pe[ch*8+7]); }
                else { a0 = __float_as_uint(pe[ch*8]); a1 = __float_as_uint(pe[ch*8+2]); a2 = __float_as_uint(pe[ch*8+4]); a3 = __float_as_uint(pe[ch*8+6]); }
                if (MODE & 16) {
                    const uint32_t addr = pbase + (c >> 1) * 16384 + ((((c & 1) * 4 + ch) ^ rsw) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a0), "r"(a1), "r"(a2), "r"(a3) : "memory");
                } else {
                    accum += __uint_as_float(a0 ^ a1 ^ a2 ^ a3);
                }
            }
        }
        // keep the inputs "fresh" so the loop is not hoisted: rotate registers cheaply
        const float t = s[0];
#pragma unroll
        for (int i = 0; i < 127; ++i) s[i] = s[i + 1];
        s[127] = t;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = accum + s[5];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, float* out, long long* cyc, float* in) {
    const int iters = 200;
    for (int warps : {4, 8}) {
        cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768);
        k<MODE><<<1, warps * 32, 4 * 32768>>>(out, cyc, in, 0.17f, 3.0f, iters);
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-34s warps/SM %2d: %7.0f cycles per 128-element row pass  (err %s)\n", name, warps, (double)c / iters, cudaGetErrorString(cudaGetLastError()));
    }
}
int main() {
    float *out, *in; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024); cudaMalloc(&in, 4096 * 4);
    cudaMemset(in, 0, 4096 * 4);
    run<2>("MUFU only", out, cyc, in);
    run<3>("FFMA+MUFU", out, cyc, in);
    run<3 + 4>("FFMA+MUFU+PRMT", out, cyc, in);
    run<3 + 8>("FFMA+MUFU+F2FP", out, cyc, in);
    run<3 + 4 + 16>("FFMA+MUFU+PRMT+STS", out, cyc, in);
    run<3 + 8 + 16>("FFMA+MUFU+F2FP+STS", out, cyc, in);
    run<3 + 4 + 16 + 32>("max+FFMA+MUFU+PRMT+STS", out, cyc, in);
    return 0;
}
