// HBM-bound row-wise kernels of the Next-DiT block: everything between the GEMMs.
// One warp owns one token row (D <= 4096), 16-byte vector loads, shuffle-only reductions.
// Rounding points follow the reference under torch.autocast(bf16):
//   RMSNorm           lumina_next_t2i/models/components.py:40,53-54  (fp32 normalise -> bf16 -> * weight -> bf16)
//   modulate          lumina_next_t2i/models/model.py:28-29          (x * (1+scale), both bf16)
//   gated residual    model.py:597-610                               (x + tanh(gate) * norm2(...))
//   q/k LayerNorm     model.py:361-362 (fp32 out), RoPE :255-282 (fp32) then .to(bf16) :371
//   final layer       model.py:657-662
//   patchify/embed    model.py:770-788; unpatchify :743-768; CFG combine :901-913
//   timestep embed    model.py:64-87; caption pool :847-851
#include <math.h>

#include <stdlib.h>

#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace ndit {

constexpr int ROW_WARPS = 8;      // rows (warps) per block
constexpr int MAX_VEC = 16;       // per-lane 8-element vectors -> D <= 4096

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void load8(const bf16* p, float* f) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8(bf16* p, const float* f) {
    uint4 u;
    u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
}

// ---------------------------------------------------------------------------------------------
// Gated residual update + pre-norm + modulate.  NV = ceil(D / 256) vectors per lane.
//   if (o)  X = bf16(X + bf16(tanh_g * bf16(bf16(rmsnorm(o)) * w_post)))
//   u = bf16(bf16(bf16(rmsnorm(X)) * w_pre) * onepls)
// Rows stay packed (bf16) in registers between the passes so two 256-thread CTAs fit per SM.
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 u;
    u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
    return u;
}

__device__ __forceinline__ bf162 as_bf162(uint32_t u) { return *reinterpret_cast<bf162*>(&u); }
__device__ __forceinline__ uint32_t as_u32(bf162 h) { return *reinterpret_cast<uint32_t*>(&h); }

// (Tried and dropped: a persistent variant that streams the X and o rows through shared memory with 1-D bulk copies, one row
// ahead of the arithmetic, 8 warps per SM: 124.7 vs 111.4 ms of row-wise time per latent - slower than this register-resident
// kernel at 16 warps per SM.)
// bf16 x bf16 -> bf16 and bf16 + bf16 -> bf16 are done with the native packed instructions (HMUL2.BF16 / HADD2.BF16:
// exact product or sum, one rounding) - the same result PyTorch produces by computing in fp32 and rounding.
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32, 2)
resid_rms_mod_kernel(bf16* __restrict__ X, const bf16* __restrict__ o, const bf16* __restrict__ w_post,
                     const bf16* __restrict__ tanh_g, const bf16* __restrict__ w_pre,
                     const bf16* __restrict__ onepls, const bf16* __restrict__ shift, bf16* __restrict__ u, int M,
                     int rows_per_batch, int D, int mod_stride, float eps) {
    pdl_trigger();
    pdl_wait();
    const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    const int nvec = D >> 3;
    const int b = row / rows_per_batch;
    const size_t off = static_cast<size_t>(row) * D;
    uint4 xv[NV];
    float ss2 = 0.f;      // sum of squares of the (updated) residual row
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) xv[i] = *reinterpret_cast<const uint4*>(X + off + v * 8);
    }
    if (o != nullptr) {
        uint4 ov[NV];
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + i * 32;
            if (v < nvec) ov[i] = *reinterpret_cast<const uint4*>(o + off + v * 8);
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + i * 32;
            if (v < nvec) {
                const uint32_t w4[4] = {ov[i].x, ov[i].y, ov[i].z, ov[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack_bf16(w4[j]);
                    ss = fmaf(f.x, f.x, ss);
                    ss = fmaf(f.y, f.y, ss);
                }
            }
        }
        // w_post == nullptr: no post-norm, the branch output is gated as it is (Flag-DiT, lumina_t2i model.py:596-609)
        const float rinv = w_post != nullptr ? rsqrtf(warp_sum(ss) / D + eps) : 0.f;
        const bf16* tg = tanh_g + static_cast<size_t>(b) * mod_stride;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + i * 32;
            if (v < nvec) {
                const uint4 wq = w_post != nullptr ? *reinterpret_cast<const uint4*>(w_post + v * 8) : make_uint4(0, 0, 0, 0);
                const uint4 gq = *reinterpret_cast<const uint4*>(tg + v * 8);
                const uint32_t o4[4] = {ov[i].x, ov[i].y, ov[i].z, ov[i].w};
                const uint32_t w4[4] = {wq.x, wq.y, wq.z, wq.w};
                const uint32_t g4[4] = {gq.x, gq.y, gq.z, gq.w};
                uint32_t x4[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    bf162 r2 = as_bf162(o4[j]);
                    if (w_post != nullptr) {
                        const float2 of = unpack_bf16(o4[j]);
                        const bf162 n2 = __floats2bfloat162_rn(of.x * rinv, of.y * rinv);
                        r2 = __hmul2_rn(n2, as_bf162(w4[j]));
                    }
                    const bf162 p2 = __hmul2_rn(as_bf162(g4[j]), r2);
                    const bf162 x2 = __hadd2_rn(as_bf162(x4[j]), p2);
                    x4[j] = as_u32(x2);
                    const float2 xf = __bfloat1622float2(x2);
                    ss2 = fmaf(xf.x, xf.x, ss2);
                    ss2 = fmaf(xf.y, xf.y, ss2);
                }
                xv[i] = make_uint4(x4[0], x4[1], x4[2], x4[3]);
                *reinterpret_cast<uint4*>(X + off + v * 8) = xv[i];
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + i * 32;
            if (v < nvec) {
                const uint32_t x4[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack_bf16(x4[j]);
                    ss2 = fmaf(f.x, f.x, ss2);
                    ss2 = fmaf(f.y, f.y, ss2);
                }
            }
        }
    }
    if (u == nullptr) return;
    const float rinv = rsqrtf(warp_sum(ss2) / D + eps);
    const bf16* op = onepls + static_cast<size_t>(b) * mod_stride;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            const uint4 wq = *reinterpret_cast<const uint4*>(w_pre + v * 8);
            const uint4 sq = *reinterpret_cast<const uint4*>(op + v * 8);
            const uint4 hq = shift != nullptr ? *reinterpret_cast<const uint4*>(shift + static_cast<size_t>(b) * mod_stride + v * 8)
                                              : make_uint4(0, 0, 0, 0);
            const uint32_t w4[4] = {wq.x, wq.y, wq.z, wq.w};
            const uint32_t s4[4] = {sq.x, sq.y, sq.z, sq.w};
            const uint32_t h4[4] = {hq.x, hq.y, hq.z, hq.w};
            const uint32_t x4[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
            uint32_t r4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 xf = unpack_bf16(x4[j]);
                const bf162 n2 = __floats2bfloat162_rn(xf.x * rinv, xf.y * rinv);
                bf162 m2 = __hmul2_rn(__hmul2_rn(n2, as_bf162(w4[j])), as_bf162(s4[j]));
                if (shift != nullptr) m2 = __hadd2_rn(m2, as_bf162(h4[j]));      // modulate with shift (lumina_t2i model.py:28-29)
                r4[j] = as_u32(m2);
            }
            *reinterpret_cast<uint4*>(u + off + v * 8) = make_uint4(r4[0], r4[1], r4[2], r4[3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Same op with ONE ROW PER 4 WARPS (128-thread block, <= 3 vectors per thread): ~50 registers instead of 108, so 8 blocks
// = 32 warps per SM instead of 16 - the register-resident kernel above is latency-bound at 23 % occupancy
// (profiles/r01_ncu_resid_rms_mod_full_final.txt).  Two block-level reductions (warp shuffle + 4 partials in shared memory).
template <int NV>
__global__ void __launch_bounds__(128, 8)
resid_rms_mod4_kernel(bf16* __restrict__ X, const bf16* __restrict__ o, const bf16* __restrict__ w_post,
                      const bf16* __restrict__ tanh_g, const bf16* __restrict__ w_pre, const bf16* __restrict__ onepls,
                      const bf16* __restrict__ shift, bf16* __restrict__ u, int M, int rows_per_batch, int D, int mod_stride,
                      float eps) {
    __shared__ float red1[4], red2[4];
    pdl_trigger();
    pdl_wait();
    const int row = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nvec = D >> 3;
    const int b = row / rows_per_batch;
    const size_t off = static_cast<size_t>(row) * D;
    uint4 xv[NV];
    float ss2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = tid + i * 128;
        if (v < nvec) xv[i] = *reinterpret_cast<const uint4*>(X + off + v * 8);
    }
    if (o != nullptr) {
        uint4 ov[NV];
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = tid + i * 128;
            if (v < nvec) ov[i] = *reinterpret_cast<const uint4*>(o + off + v * 8);
        }
        float rinv = 0.f;
        if (w_post != nullptr) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int v = tid + i * 128;
                if (v < nvec) {
                    const uint32_t w4[4] = {ov[i].x, ov[i].y, ov[i].z, ov[i].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = unpack_bf16(w4[j]);
                        ss = fmaf(f.x, f.x, ss);
                        ss = fmaf(f.y, f.y, ss);
                    }
                }
            }
            ss = warp_sum(ss);
            if (lane == 0) red1[warp] = ss;
            __syncthreads();
            rinv = rsqrtf((red1[0] + red1[1] + red1[2] + red1[3]) / D + eps);
        }
        const bf16* tg = tanh_g + static_cast<size_t>(b) * mod_stride;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = tid + i * 128;
            if (v < nvec) {
                const uint4 wq = w_post != nullptr ? *reinterpret_cast<const uint4*>(w_post + v * 8) : make_uint4(0, 0, 0, 0);
                const uint4 gq = *reinterpret_cast<const uint4*>(tg + v * 8);
                const uint32_t o4[4] = {ov[i].x, ov[i].y, ov[i].z, ov[i].w};
                const uint32_t w4[4] = {wq.x, wq.y, wq.z, wq.w};
                const uint32_t g4[4] = {gq.x, gq.y, gq.z, gq.w};
                uint32_t x4[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    bf162 r2 = as_bf162(o4[j]);
                    if (w_post != nullptr) {
                        const float2 of = unpack_bf16(o4[j]);
                        const bf162 n2 = __floats2bfloat162_rn(of.x * rinv, of.y * rinv);
                        r2 = __hmul2_rn(n2, as_bf162(w4[j]));
                    }
                    const bf162 p2 = __hmul2_rn(as_bf162(g4[j]), r2);
                    const bf162 x2 = __hadd2_rn(as_bf162(x4[j]), p2);
                    x4[j] = as_u32(x2);
                    const float2 xf = __bfloat1622float2(x2);
                    ss2 = fmaf(xf.x, xf.x, ss2);
                    ss2 = fmaf(xf.y, xf.y, ss2);
                }
                xv[i] = make_uint4(x4[0], x4[1], x4[2], x4[3]);
                *reinterpret_cast<uint4*>(X + off + v * 8) = xv[i];
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = tid + i * 128;
            if (v < nvec) {
                const uint32_t x4[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack_bf16(x4[j]);
                    ss2 = fmaf(f.x, f.x, ss2);
                    ss2 = fmaf(f.y, f.y, ss2);
                }
            }
        }
    }
    if (u == nullptr) return;
    ss2 = warp_sum(ss2);
    if (lane == 0) red2[warp] = ss2;
    __syncthreads();
    const float rinv2 = rsqrtf((red2[0] + red2[1] + red2[2] + red2[3]) / D + eps);
    const bf16* op = onepls + static_cast<size_t>(b) * mod_stride;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = tid + i * 128;
        if (v < nvec) {
            const uint4 wq = *reinterpret_cast<const uint4*>(w_pre + v * 8);
            const uint4 sq = *reinterpret_cast<const uint4*>(op + v * 8);
            const uint4 hq = shift != nullptr ? *reinterpret_cast<const uint4*>(shift + static_cast<size_t>(b) * mod_stride + v * 8)
                                              : make_uint4(0, 0, 0, 0);
            const uint32_t w4[4] = {wq.x, wq.y, wq.z, wq.w};
            const uint32_t s4[4] = {sq.x, sq.y, sq.z, sq.w};
            const uint32_t h4[4] = {hq.x, hq.y, hq.z, hq.w};
            const uint32_t x4[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
            uint32_t r4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 xf = unpack_bf16(x4[j]);
                const bf162 n2 = __floats2bfloat162_rn(xf.x * rinv2, xf.y * rinv2);
                bf162 m2 = __hmul2_rn(__hmul2_rn(n2, as_bf162(w4[j])), as_bf162(s4[j]));
                if (shift != nullptr) m2 = __hadd2_rn(m2, as_bf162(h4[j]));
                r4[j] = as_u32(m2);
            }
            *reinterpret_cast<uint4*>(u + off + v * 8) = make_uint4(r4[0], r4[1], r4[2], r4[3]);
        }
    }
}

cudaError_t resid_rms_mod(bf16* X, const bf16* o, const bf16* w_post, const bf16* tanh_g, const bf16* w_pre,
                          const bf16* onepls, const bf16* shift, bf16* u, int M, int rows_per_batch, int D, int mod_stride,
                          float eps, cudaStream_t s) {
    if (D % 8 != 0 || D > MAX_VEC * 256 || mod_stride % 8 != 0) return cudaErrorInvalidValue;
    const int nv = (D / 8 + 31) / 32;
    // One row per 128-thread block (32 warps per SM).  Measured inside the step (round 2, config 2): row-wise class 112.7 ->
    // 100.4 ms per solve, 883.5 -> 876.8 ms total (profiles/r02_bench_resid4_ab.json).  NDIT_RESID4=0 selects the one-row-per-warp
    // kernel; outputs differ from it in 1 ulp on ~1e-5 of the elements (reduction order).
    static const int resid4_env = getenv("NDIT_RESID4") ? atoi(getenv("NDIT_RESID4")) : 1;
    if (resid4_env && M >= 1024 && D / 8 <= 3 * 128) {
        if (D / 8 <= 2 * 128) return launch_k(resid_rms_mod4_kernel<2>, dim3(M), dim3(128), 0, s, X, o, w_post, tanh_g, w_pre, onepls, shift, u, M,
                                              rows_per_batch, D, mod_stride, eps);
        return launch_k(resid_rms_mod4_kernel<3>, dim3(M), dim3(128), 0, s, X, o, w_post, tanh_g, w_pre, onepls, shift, u, M, rows_per_batch,
                        D, mod_stride, eps);
    }
    const dim3 grid((M + ROW_WARPS - 1) / ROW_WARPS), block(ROW_WARPS * 32);
#define LAUNCH(NVV)                                                                                             \
    return launch_k(resid_rms_mod_kernel<NVV>, grid, block, 0, s, X, o, w_post, tanh_g, w_pre, onepls, shift, u, M, \
                    rows_per_batch, D, mod_stride, eps)
    if (nv <= 3) LAUNCH(3);
    else if (nv <= 9) LAUNCH(9);
    else if (nv <= 12) LAUNCH(12);
    else LAUNCH(MAX_VEC);
#undef LAUNCH
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Final layer, first half: last gated residual + xn = bf16(LayerNorm(no affine, eps 1e-6)(x) * (1+scale) (+ shift)).
// (model.py:657-662; the Linear D -> patch*patch*C_out + bias that follows is a tcgen05 GEMM with a bias epilogue: the first
// version of this kernel did the projection itself and re-read the 147 KB weight for every row - 204 us for one launch.)
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
final_norm_kernel(const bf16* __restrict__ X, const bf16* __restrict__ o, const bf16* __restrict__ w_post,
                  const bf16* __restrict__ tanh_g, const bf16* __restrict__ onepls, const bf16* __restrict__ shift,
                  bf16* __restrict__ xn, int M, int rows_per_batch, int D, int mod_stride, float eps_rms, bf16* x_out) {
    const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    const int nvec = D >> 3;
    const int b = row / rows_per_batch;
    const size_t off = static_cast<size_t>(row) * D;
    float x[NV][8];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) load8(X + off + v * 8, x[i]);
        else {
#pragma unroll
            for (int e = 0; e < 8; ++e) x[i][e] = 0.f;
        }
    }
    if (o != nullptr) {
        float ov[NV][8];
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + i * 32;
            if (v < nvec) {
                load8(o + off + v * 8, ov[i]);
#pragma unroll
                for (int e = 0; e < 8; ++e) ss += ov[i][e] * ov[i][e];
            }
        }
        const float rinv = rsqrtf(warp_sum(ss) / D + eps_rms);
        const bf16* tg = tanh_g + static_cast<size_t>(b) * mod_stride;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + i * 32;
            if (v < nvec) {
                float w[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, g[8];
                if (w_post != nullptr) load8(w_post + v * 8, w);
                load8(tg + v * 8, g);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float n = w_post != nullptr ? bf16_round(bf16_round(ov[i][e] * rinv) * w[e]) : ov[i][e];
                    x[i][e] = bf16_round(x[i][e] + bf16_round(g[e] * n));
                }
                if (x_out != nullptr) store8(x_out + off + v * 8, x[i]);     // debug tap of the last block's residual stream
            }
        }
    }
    // LayerNorm, no affine, eps 1e-6, fp32 (two-pass)
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) sum += x[i][e];
    const float mean = warp_sum(sum) / D;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float d = x[i][e] - mean;
                sq += d * d;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / D + 1e-6f);
    const bf16* op = onepls + static_cast<size_t>(b) * mod_stride;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            float sc[8], sh[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, r[8];
            load8(op + v * 8, sc);
            if (shift != nullptr) load8(shift + static_cast<size_t>(b) * mod_stride + v * 8, sh);
#pragma unroll
            for (int e = 0; e < 8; ++e) r[e] = (x[i][e] - mean) * rstd * sc[e] + sh[e];
            store8(xn + off + v * 8, r);
        }
    }
}

cudaError_t final_norm(const bf16* X, const bf16* o, const bf16* w_post, const bf16* tanh_g, const bf16* onepls,
                       const bf16* shift, bf16* xn, int M, int rows_per_batch, int D, int mod_stride, float eps, cudaStream_t s, bf16* x_out) {
    if (D % 8 != 0 || D > MAX_VEC * 256) return cudaErrorInvalidValue;
    const int nv = (D / 8 + 31) / 32;
    const dim3 grid((M + ROW_WARPS - 1) / ROW_WARPS), block(ROW_WARPS * 32);
#define LAUNCH(NVV)                                                                                              \
    final_norm_kernel<NVV><<<grid, block, 0, s>>>(X, o, w_post, tanh_g, onepls, shift, xn, M, rows_per_batch, D, \
                                                  mod_stride, eps, x_out)
    if (nv <= 3) LAUNCH(3);
    else if (nv <= 9) LAUNCH(9);
    else if (nv <= 12) LAUNCH(12);
    else LAUNCH(MAX_VEC);
#undef LAUNCH
    return cudaGetLastError();
}

// y_embedder lookup (models.py:216-225): out[b, :] = table[label[b], :]  (fp32 storage of bf16 values)
__global__ void gather_label_rows_kernel(const bf16* __restrict__ table, const long long* __restrict__ labels,
                                         float* __restrict__ out, int n_rows, int width) {
    const int b = blockIdx.x;
    long long lab = labels[b];
    if (lab < 0) lab = 0;
    if (lab >= n_rows) lab = n_rows - 1;
    for (int c = threadIdx.x; c < width; c += blockDim.x) out[static_cast<size_t>(b) * width + c] = __bfloat162float(table[lab * width + c]);
}

cudaError_t gather_label_rows(const bf16* table, const long long* labels, float* out, int B, int n_rows, int width, cudaStream_t s) {
    gather_label_rows_kernel<<<B, 256, 0, s>>>(table, labels, out, n_rows, width);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Patchify + x_embedder.  One block per token, thread d-strided over D.
__global__ void patch_embed_kernel(const bf16* __restrict__ x, const bf16* __restrict__ Wx, const bf16* __restrict__ bx,
                                   const bf16* __restrict__ eol, bf16* __restrict__ X, int n_unique, int C, int Hh, int Ww, int D) {
    const int Wp = Ww >> 1, Hp = Hh >> 1;
    const int Wt = Wp + (eol != nullptr ? 1 : 0);   // Flag-DiT: a learned [eol] token closes every row of patches
    const int tok = blockIdx.x;               // over B * Hp * Wt
    const int N = Hp * Wt;
    const int b = tok / N, t = tok % N;
    const int bi = b % n_unique;              // second half of the batch re-uses the first (model.py:901-902)
    const int i = t / Wt, j = t % Wt;
    if (j == Wp) {                            // lumina_t2i model.py:779-785
        for (int d = threadIdx.x; d < D; d += blockDim.x) X[static_cast<size_t>(tok) * D + d] = eol[d];
        return;
    }
    __shared__ float patch[64];
    const int K = C * 4;
    if (threadIdx.x < K) {
        const int c = threadIdx.x >> 2, ph = (threadIdx.x >> 1) & 1, pw = threadIdx.x & 1;
        patch[threadIdx.x] = __bfloat162float(x[((static_cast<size_t>(bi) * C + c) * Hh + 2 * i + ph) * Ww + 2 * j + pw]);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float acc = 0.f;
        for (int k = 0; k < K; k += 8) {
            float w[8];
            load8(Wx + static_cast<size_t>(d) * K + k, w);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc += patch[k + e] * w[e];
        }
        X[static_cast<size_t>(tok) * D + d] = __float2bfloat16_rn(acc + __bfloat162float(bx[d]));
    }
}

cudaError_t patch_embed(const bf16* x, const bf16* Wx, const bf16* bx, const bf16* eol, bf16* X, int B, int n_unique, int C,
                        int Hh, int Ww, int D, cudaStream_t s) {
    if (C * 4 > 64 || (C * 4) % 8 != 0 || (Hh & 1) || (Ww & 1)) return cudaErrorInvalidValue;
    patch_embed_kernel<<<B * (Hh / 2) * (Ww / 2 + (eol != nullptr ? 1 : 0)), 256, 0, s>>>(x, Wx, bx, eol, X, n_unique, C, Hh, Ww, D);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Conditioning inputs: sinusoidal timestep features and LayerNorm'd masked-mean caption.
__global__ void cond_prepare_kernel(float t, const float* __restrict__ t_rows, const bf16* __restrict__ cap,
                                    const uint8_t* __restrict__ mask, const bf16* __restrict__ ln_w, const bf16* __restrict__ ln_b,
                                    float* __restrict__ tf, float* __restrict__ pool, int T, int C, int do_caption) {
    const int b = blockIdx.x;
    if (t_rows != nullptr) t = t_rows[b];    // NextDiT.forward: one timestep per row
    // model.py:64-87: freqs = exp(-ln(1e4) * i / 128); [cos | sin]; cast to the weight dtype
    for (int i = threadIdx.x; i < 128; i += blockDim.x) {
        const float f = expf(-logf(10000.0f) * static_cast<float>(i) / 128.0f);
        const float a = t * f;
        tf[b * 256 + i] = bf16_round(cosf(a));
        tf[b * 256 + 128 + i] = bf16_round(sinf(a));
    }
    if (!do_caption) return;
    extern __shared__ float sh[];            // C pooled values + reduction scratch
    float* pv = sh;
    __shared__ float red[32];
    float cnt = 0.f;
    for (int k = 0; k < T; ++k) cnt += mask[b * T + k] ? 1.f : 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc = 0.f;
        for (int k = 0; k < T; ++k)
            if (mask[b * T + k]) acc += __bfloat162float(cap[(static_cast<size_t>(b) * T + k) * C + c]);
        pv[c] = bf16_round(acc / cnt);       // .to(cap_feats) (model.py:849)
    }
    __syncthreads();
    // LayerNorm(C), affine, fp32 (two-pass)
    float s = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) s += pv[c];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tot += red[w];
    const float mean = tot / C;
    __syncthreads();
    float q = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float d = pv[c] - mean;
        q += d * d;
    }
    q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
    __syncthreads();
    tot = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tot += red[w];
    const float rstd = rsqrtf(tot / C + 1e-5f);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float y = (pv[c] - mean) * rstd * __bfloat162float(ln_w[c]) + __bfloat162float(ln_b[c]);
        pool[static_cast<size_t>(b) * C + c] = bf16_round(y);   // Linear casts its fp32 input to bf16
    }
}

cudaError_t cond_prepare(float t, const float* t_rows, const bf16* cap, const uint8_t* mask, const bf16* ln_w, const bf16* ln_b, float* tf,
                         float* pool, int B, int T, int C, int do_caption, cudaStream_t s) {
    cond_prepare_kernel<<<B, 256, do_caption ? C * sizeof(float) : 0, s>>>(t, t_rows, cap, mask, ln_w, ln_b, tf, pool, T, C,
                                                                         do_caption);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Small-batch Linear (B <= 4 rows share every weight row): HBM-bound on the weight matrix.  The input rows are
// staged once per block in shared memory; each warp streams GEMV_RPW weight rows at a time with 16-byte loads.
constexpr int GEMV_MAXB = 4;
constexpr int GEMV_RPW = 4;                   // output rows per warp (independent loads in flight)
template <int NB>
__global__ void __launch_bounds__(256)
gemv_rows_kernel(const float* __restrict__ in, const bf16* __restrict__ W, const bf16* __restrict__ bias,
                 const float* __restrict__ addend, float* __restrict__ out, bf16* __restrict__ out_b, int O, int K,
                 int in_silu, int post, int adaln_D, int adaln_blocks, int adaln_kind) {
    extern __shared__ float xin[];            // [NB][K]
    for (int i = threadIdx.x; i < NB * K; i += blockDim.x) {
        float v = in[i];
        if (in_silu) v = bf16_round(silu_f(v));
        xin[i] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int o0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * GEMV_RPW;
    if (o0 >= O) return;
    float acc[GEMV_RPW][NB];
#pragma unroll
    for (int r = 0; r < GEMV_RPW; ++r)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[r][b] = 0.f;
    for (int k = lane * 8; k < K; k += 256) {
        uint4 wq[GEMV_RPW];
#pragma unroll
        for (int r = 0; r < GEMV_RPW; ++r) {
            const int o = o0 + r < O ? o0 + r : O - 1;
            wq[r] = *reinterpret_cast<const uint4*>(W + static_cast<size_t>(o) * K + k);
        }
        float x[NB][8];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const float4 a = *reinterpret_cast<const float4*>(xin + b * K + k);
            const float4 c = *reinterpret_cast<const float4*>(xin + b * K + k + 4);
            x[b][0] = a.x; x[b][1] = a.y; x[b][2] = a.z; x[b][3] = a.w; x[b][4] = c.x; x[b][5] = c.y; x[b][6] = c.z; x[b][7] = c.w;
        }
#pragma unroll
        for (int r = 0; r < GEMV_RPW; ++r) {
            float w[8];
            unpack8(wq[r], w);
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[r][b] = fmaf(x[b][e], w[e], acc[r][b]);
        }
    }
#pragma unroll
    for (int r = 0; r < GEMV_RPW; ++r)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[r][b] = warp_sum(acc[r][b]);
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < GEMV_RPW; ++r) {
            const int o = o0 + r;
            if (o >= O) break;
            const float bv = bias ? __bfloat162float(bias[o]) : 0.f;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                float y = bf16_round(acc[r][b] + bv);
                if (addend) y = bf16_round(y + addend[static_cast<size_t>(b) * O + o]);
                if (post == POST_SILU) y = bf16_round(silu_f(y));
                else if (post == POST_ADALN) {
                    // ADALN_NEXT  per layer [scale_msa | gate_msa | scale_mlp | gate_mlp] (model.py:595), final [scale]
                    // ADALN_CLASS same layers, final [shift | scale] (Next-DiT-ImageNet models.py:829-833)
                    // ADALN_FLAG  per layer [shift | scale | gate] x2, plain gates (lumina_t2i model.py:596-609), final [shift | scale]
                    // stored: scale -> bf16(1+scale); Next-DiT gate -> bf16(tanh(gate)); shift and Flag-DiT gate as they are
                    const int chunk = o / adaln_D;
                    int kind;                 // 0 shift / plain gate (raw), 1 scale, 2 tanh gate
                    // adaln_blocks = number of per-layer chunks in front of the final layer's
                    if (adaln_kind == ADALN_FLAG) {
                        const int per = chunk < adaln_blocks ? chunk % 3 : (chunk - adaln_blocks);   // final: 0 shift, 1 scale
                        kind = per == 1 ? 1 : 0;
                    } else if (chunk < adaln_blocks) {
                        kind = (chunk & 1) ? 2 : 1;               // (scale, gate) pairs: attention, FFN, (second FFN of the MoE "both" block)
                    } else {
                        kind = (adaln_kind == ADALN_CLASS && chunk == adaln_blocks) ? 0 : 1;
                    }
                    if (kind == 1) y = bf16_round(1.0f + y);
                    else if (kind == 2) y = bf16_round(tanhf(y));
                }
                if (out_b) out_b[static_cast<size_t>(b) * O + o] = __float2bfloat16_rn(y);   // y is bf16-representable
                else out[static_cast<size_t>(b) * O + o] = y;
            }
        }
    }
}

cudaError_t gemv_rows(const float* in, const bf16* W, const bf16* bias, const float* addend, float* out, bf16* out_b,
                      int B, int O, int K, int in_silu, int post, int adaln_D, int adaln_blocks, int adaln_kind, cudaStream_t s) {
    if (B < 1 || B > GEMV_MAXB || K % 8 != 0 || static_cast<size_t>(B) * K * sizeof(float) > 64 * 1024) return cudaErrorInvalidValue;
    const int grid = (O + 8 * GEMV_RPW - 1) / (8 * GEMV_RPW);
    const size_t sh = static_cast<size_t>(B) * K * sizeof(float);
    if (sh > 48 * 1024) {
        static PerDeviceFlag flags;
        bool& configured = flags.here();
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(gemv_rows_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(gemv_rows_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            if (e != cudaSuccess) return e;
            configured = true;
        }
    }
#define LAUNCH(NBB) \
    gemv_rows_kernel<NBB><<<grid, 256, sh, s>>>(in, W, bias, addend, out, out_b, O, K, in_silu, post, adaln_D, adaln_blocks, adaln_kind)
    switch (B) {
        case 1: LAUNCH(1); break;
        case 2: LAUNCH(2); break;
        case 3: LAUNCH(3); break;
        default: LAUNCH(4); break;
    }
#undef LAUNCH
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// RoPE table: tab[token][m] = (cos, sin)(pos * w_{m/2}), m even -> row index, m odd -> column index
// (model.py:951-961).  w_i = (theta)^(-4i/hd) / linear_factor computed like the reference in fp32.
__global__ void rope_table_kernel(float2* __restrict__ tab, int Hp, int Wp, int hd, float theta, float linear_factor, int one_d) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = hd >> 1;
    if (idx >= Hp * Wp * half) return;
    const int m = idx % half, tok = idx / half;
    if (one_d) {
        // Flag-DiT (lumina_t2i model.py:925-960): 1-D table over the token index (eol tokens included),
        // freqs 1/theta^(2m/hd), positions divided by the rope scaling factor BEFORE the outer product
        const float freq = 1.0f / powf(theta, static_cast<float>(2 * m) / static_cast<float>(hd));
        const float a = (static_cast<float>(tok) / linear_factor) * freq;
        float sn, cs;
        sincosf(a, &sn, &cs);
        tab[idx] = make_float2(cs, sn);
        return;
    }
    const int i = tok / Wp, j = tok % Wp;
    const int fi = m >> 1;
    const float freq = 1.0f / powf(theta, static_cast<float>(4 * fi) / static_cast<float>(hd)) / linear_factor;
    const float pos = static_cast<float>((m & 1) ? j : i);
    const float a = pos * freq;
    float sn, cs;
    sincosf(a, &sn, &cs);
    tab[idx] = make_float2(cs, sn);
}

__global__ void broadcast_row_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, int rows, int vecs) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < rows * vecs) dst[idx] = src[idx % vecs];
}

cudaError_t broadcast_row(void* dst, const void* src, int rows, int row_bytes, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    if (row_bytes % 16 != 0) return cudaErrorInvalidValue;
    const int vecs = row_bytes / 16, n = rows * vecs;
    broadcast_row_kernel<<<(n + 255) / 256, 256, 0, s>>>(static_cast<uint4*>(dst), static_cast<const uint4*>(src), rows, vecs);
    return cudaGetLastError();
}

cudaError_t rope_table(float2* tab, int Hp, int Wp, int hd, float theta, float linear_factor, int one_d, cudaStream_t s) {
    const int n = Hp * Wp * (hd / 2);
    rope_table_kernel<<<(n + 255) / 256, 256, 0, s>>>(tab, Hp, Wp, hd, theta, linear_factor, one_d);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// q/k LayerNorm over ALL heads jointly + 2-D RoPE, in place on the fused qkv GEMM output.
// One warp per token row.  seg 0 = q (width H*hd), seg 1 = k (width Hkv*hd).
template <int NV>
__device__ __forceinline__ void ln_segment_load(const bf16* __restrict__ p, int width, int lane, uint4* raw) {
    const int nvec = width >> 3;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) raw[i] = *reinterpret_cast<const uint4*>(p + v * 8);
    }
}

template <int NV>
__device__ __forceinline__ void ln_rope_segment(bf16* __restrict__ p, int width, const bf16* __restrict__ w,
                                                const bf16* __restrict__ bsh, const float2* __restrict__ rp, int hd,
                                                int lane, const uint4* raw) {
    const int nvec = width >> 3;
    float x[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            unpack8(raw[i], x[i]);
#pragma unroll
            for (int e = 0; e < 8; ++e) sum += x[i][e];
        }
    }
    const float mean = warp_sum(sum) / width;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float d = x[i][e] - mean;
                sq += d * d;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / width + 1e-5f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            float g[8], bb[8], r[8];
            load8(w + v * 8, g);
            load8(bsh + v * 8, bb);
#pragma unroll
            for (int e = 0; e < 8; ++e) x[i][e] = (x[i][e] - mean) * rstd * g[e] + bb[e];
            if (rp != nullptr) {
                const int m0 = ((v * 8) % hd) >> 1;   // hd % 8 == 0: a vector never straddles heads
#pragma unroll
                for (int pr = 0; pr < 4; ++pr) {
                    const float2 cs = rp[m0 + pr];
                    const float a = x[i][2 * pr], bq = x[i][2 * pr + 1];
                    r[2 * pr] = a * cs.x - bq * cs.y;
                    r[2 * pr + 1] = a * cs.y + bq * cs.x;
                }
                store8(p + v * 8, r);
            } else {
                store8(p + v * 8, x[i]);
            }
        }
    }
}

template <int NVQ, int NVK>
__global__ void __launch_bounds__(ROW_WARPS * 32)
ln_rope_qk_kernel(bf16* __restrict__ qkv, int ld, const bf16* __restrict__ qw, const bf16* __restrict__ qb,
                  const bf16* __restrict__ kw, const bf16* __restrict__ kb, const float2* __restrict__ rope, int M,
                  int N_tokens, int H, int Hkv, int hd) {
    pdl_trigger();
    pdl_wait();
    const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    bf16* p = qkv + static_cast<size_t>(row) * ld;
    const float2* rp = rope + static_cast<size_t>(row % N_tokens) * (hd >> 1);
    uint4 rq[NVQ], rk[NVK];          // both segments' loads are in flight before any math starts
    ln_segment_load<NVQ>(p, H * hd, lane, rq);
    ln_segment_load<NVK>(p + H * hd, Hkv * hd, lane, rk);
    ln_rope_segment<NVQ>(p, H * hd, qw, qb, rp, hd, lane, rq);
    ln_rope_segment<NVK>(p + H * hd, Hkv * hd, kw, kb, rp, hd, lane, rk);
}

// Same op with ONE ROW PER 4 WARPS (128-thread block): q and k of one token are one contiguous run of (H + Hkv) * hd elements
// = nq + nk 16-byte vectors (2B GQA: 288 + 72 = 360 -> 3 vectors per thread, ~50 registers), so 8+ blocks = 32+ warps are
// resident per SM instead of the 16 of the one-row-per-warp kernel above, which ncu shows latency-bound at 0.38 of the HBM
// peak (profiles/r01_ncu_ln_rope_qk_full_final.txt).  Both LayerNorms (q over H*hd, k over Hkv*hd) are reduced together:
// two block-level reductions of a (q, k) pair each (mean, then centred second moment - the same two-pass form as above).
template <int NV>
__global__ void __launch_bounds__(128, NV <= 3 ? 8 : 4)
ln_rope_qk4_kernel(bf16* __restrict__ qkv, int ld, const bf16* __restrict__ qw, const bf16* __restrict__ qb,
                   const bf16* __restrict__ kw, const bf16* __restrict__ kb, const float2* __restrict__ rope, int N_tokens, int nq,
                   int nk, int hd) {
    __shared__ float red[2][4][2];
    pdl_trigger();
    pdl_wait();
    const int row = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntot = nq + nk;
    bf16* p = qkv + static_cast<size_t>(row) * ld;
    const float2* rp = rope + static_cast<size_t>(row % N_tokens) * (hd >> 1);
    uint4 raw[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = tid + i * 128;
        if (v < ntot) raw[i] = *reinterpret_cast<const uint4*>(p + v * 8);
    }
    float x[NV][8];
    float s_q = 0.f, s_k = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = tid + i * 128;
        if (v < ntot) {
            unpack8(raw[i], x[i]);
            float a = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) a += x[i][e];
            if (v < nq) s_q += a; else s_k += a;
        }
    }
    s_q = warp_sum(s_q); s_k = warp_sum(s_k);
    if (lane == 0) { red[0][warp][0] = s_q; red[0][warp][1] = s_k; }
    __syncthreads();
    const float mean_q = (red[0][0][0] + red[0][1][0] + red[0][2][0] + red[0][3][0]) / static_cast<float>(nq * 8);
    const float mean_k = (red[0][0][1] + red[0][1][1] + red[0][2][1] + red[0][3][1]) / static_cast<float>(nk * 8);
    float v_q = 0.f, v_k = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = tid + i * 128;
        if (v < ntot) {
            const float mu = v < nq ? mean_q : mean_k;
            float a = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float d = x[i][e] - mu;
                a += d * d;
            }
            if (v < nq) v_q += a; else v_k += a;
        }
    }
    v_q = warp_sum(v_q); v_k = warp_sum(v_k);
    if (lane == 0) { red[1][warp][0] = v_q; red[1][warp][1] = v_k; }
    __syncthreads();
    const float rstd_q = rsqrtf((red[1][0][0] + red[1][1][0] + red[1][2][0] + red[1][3][0]) / static_cast<float>(nq * 8) + 1e-5f);
    const float rstd_k = rsqrtf((red[1][0][1] + red[1][1][1] + red[1][2][1] + red[1][3][1]) / static_cast<float>(nk * 8) + 1e-5f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = tid + i * 128;
        if (v < ntot) {
            const bool isq = v < nq;
            const int vv = isq ? v : v - nq;                 // vector index inside its segment
            const float mu = isq ? mean_q : mean_k, rs = isq ? rstd_q : rstd_k;
            float g[8], bb[8], r[8];
            load8((isq ? qw : kw) + vv * 8, g);
            load8((isq ? qb : kb) + vv * 8, bb);
#pragma unroll
            for (int e = 0; e < 8; ++e) x[i][e] = (x[i][e] - mu) * rs * g[e] + bb[e];
            const int m0 = ((vv * 8) % hd) >> 1;             // hd % 8 == 0: a vector never straddles heads
#pragma unroll
            for (int pr = 0; pr < 4; ++pr) {
                const float2 cs = rp[m0 + pr];
                const float a = x[i][2 * pr], bq = x[i][2 * pr + 1];
                r[2 * pr] = a * cs.x - bq * cs.y;
                r[2 * pr + 1] = a * cs.y + bq * cs.x;
            }
            store8(p + v * 8, r);
        }
    }
}

cudaError_t ln_rope_qk(bf16* qkv, int ld, const bf16* qw, const bf16* qb, const bf16* kw, const bf16* kb,
                       const float2* rope, int M, int N_tokens, int H, int Hkv, int hd, cudaStream_t s) {
    if (hd % 8 != 0 || H * hd > 12 * 256 || Hkv * hd > 12 * 256 || ld % 8 != 0) return cudaErrorInvalidValue;
    // one row per 128-thread block (NDIT_LNROPE4=0 selects the one-row-per-warp kernel below)
    static const int lnrope4_env = getenv("NDIT_LNROPE4") ? atoi(getenv("NDIT_LNROPE4")) : 1;
    const int nq = H * hd / 8, nk = Hkv * hd / 8;
    if (lnrope4_env && M >= 1024 && nq + nk <= 6 * 128) {
        const int nv = (nq + nk + 127) / 128;
#define LAUNCH4(NVV) return launch_k(ln_rope_qk4_kernel<NVV>, dim3(M), dim3(128), 0, s, qkv, ld, qw, qb, kw, kb, rope, N_tokens, nq, nk, hd)
        if (nv <= 2) LAUNCH4(2);
        if (nv <= 3) LAUNCH4(3);
        if (nv <= 4) LAUNCH4(4);
        LAUNCH4(6);
#undef LAUNCH4
    }
    const dim3 grid((M + ROW_WARPS - 1) / ROW_WARPS), block(ROW_WARPS * 32);
    if (H * hd > 9 * 256)
        return launch_k(ln_rope_qk_kernel<12, 12>, grid, block, 0, s, qkv, ld, qw, qb, kw, kb, rope, M, N_tokens, H, Hkv, hd);
    else if (Hkv * hd <= 3 * 256)
        return launch_k(ln_rope_qk_kernel<9, 3>, grid, block, 0, s, qkv, ld, qw, qb, kw, kb, rope, M, N_tokens, H, Hkv, hd);
    else
        return launch_k(ln_rope_qk_kernel<9, 9>, grid, block, 0, s, qkv, ld, qw, qb, kw, kb, rope, M, N_tokens, H, Hkv, hd);
    return cudaGetLastError();
}

// qk_norm=False (model.py:219-220: q_norm = k_norm = Identity): only the rotary embedding, q|k <- bf16(rope(float(q|k))), in place.
// One thread per 8-element vector of the contiguous q|k run of a token (nqk vectors per row).
__global__ void __launch_bounds__(256)
rope_qk_kernel(bf16* __restrict__ qkv, int ld, const float2* __restrict__ rope, int M, int N_tokens, int nq, int nqk, int hd) {
    pdl_trigger();
    pdl_wait();
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<size_t>(M) * nqk) return;
    const int row = static_cast<int>(idx / nqk), v = static_cast<int>(idx % nqk);
    const int vv = v < nq ? v : v - nq;                  // vector index inside its segment (q or k); hd % 8 == 0
    bf16* p = qkv + static_cast<size_t>(row) * ld + v * 8;
    const float2* rp = rope + static_cast<size_t>(row % N_tokens) * (hd >> 1) + (((vv * 8) % hd) >> 1);
    float x[8], r[8];
    load8(p, x);
#pragma unroll
    for (int pr = 0; pr < 4; ++pr) {
        const float2 cs = rp[pr];
        r[2 * pr] = x[2 * pr] * cs.x - x[2 * pr + 1] * cs.y;
        r[2 * pr + 1] = x[2 * pr] * cs.y + x[2 * pr + 1] * cs.x;
    }
    store8(p, r);
}

cudaError_t rope_qk(bf16* qkv, int ld, const float2* rope, int M, int N_tokens, int H, int Hkv, int hd, cudaStream_t s) {
    if (hd % 8 != 0 || ld % 8 != 0) return cudaErrorInvalidValue;
    const int nq = H * hd / 8, nqk = (H + Hkv) * hd / 8;
    const size_t total = static_cast<size_t>(M) * nqk;
    return launch_k(rope_qk_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, qkv, ld, rope, M, N_tokens, nq, nqk, hd);
}

// LayerNorm(width)+affine in place on rows (ky_norm, model.py:421), batched over layers.
__global__ void __launch_bounds__(ROW_WARPS * 32)
ln_rows_kernel(bf16* __restrict__ x, int ld, size_t lsx, const bf16* __restrict__ w, const bf16* __restrict__ b,
               size_t lsw, int M, int width) {
    const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const int l = blockIdx.y;
    uint4 raw[12];
    bf16* xp = x + l * lsx + static_cast<size_t>(row) * ld;
    ln_segment_load<12>(xp, width, threadIdx.x & 31, raw);
    ln_rope_segment<12>(xp, width, w + l * lsw, b + l * lsw, nullptr, 8, threadIdx.x & 31, raw);
}

cudaError_t ln_rows(bf16* x, int ld, size_t layer_stride_x, const bf16* w, const bf16* b, size_t layer_stride_w, int M,
                    int width, int layers, cudaStream_t s) {
    if (width % 8 != 0 || width > 12 * 256) return cudaErrorInvalidValue;
    const dim3 grid((M + ROW_WARPS - 1) / ROW_WARPS, layers), block(ROW_WARPS * 32);
    ln_rows_kernel<<<grid, block, 0, s>>>(x, ld, layer_stride_x, w, b, layer_stride_w, M, width);
    return cudaGetLastError();
}

// attention_y_norm for every layer (model.py:571,602): out[l][row] = RMS(y[row]; w[l])
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
rms_rows_layers_kernel(const bf16* __restrict__ y, const bf16* __restrict__ w, bf16* __restrict__ out, int M, int C,
                       float eps) {
    const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const int l = blockIdx.y, lane = threadIdx.x & 31, nvec = C >> 3;
    float x[NV][8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            load8(y + static_cast<size_t>(row) * C + v * 8, x[i]);
#pragma unroll
            for (int e = 0; e < 8; ++e) ss += x[i][e] * x[i][e];
        }
    }
    const float rinv = rsqrtf(warp_sum(ss) / C + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            float g[8], r[8];
            load8(w + static_cast<size_t>(l) * C + v * 8, g);
#pragma unroll
            for (int e = 0; e < 8; ++e) r[e] = bf16_round(x[i][e] * rinv) * g[e];
            store8(out + (static_cast<size_t>(l) * M + row) * C + v * 8, r);
        }
    }
}

cudaError_t rms_rows_layers(const bf16* y, const bf16* w, bf16* out, int M, int C, int layers, float eps, cudaStream_t s) {
    if (C % 8 != 0 || C > MAX_VEC * 256) return cudaErrorInvalidValue;
    const dim3 grid((M + ROW_WARPS - 1) / ROW_WARPS, layers), block(ROW_WARPS * 32);
    const int nv = (C / 8 + 31) / 32;
    if (nv <= 8) rms_rows_layers_kernel<8><<<grid, block, 0, s>>>(y, w, out, M, C, eps);
    else rms_rows_layers_kernel<MAX_VEC><<<grid, block, 0, s>>>(y, w, out, M, C, eps);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// v -> v^T (head-dim major) so the P.V product reads a K-major B operand.  32x32 smem tiles.
// Each (batch, kv-head) group owns `grows` (>= hd) rows of the destination; row `hd` is the all-ones row that makes
// the tensor core produce the softmax row sum (fill_ones_row), rows above it stay zero.
__global__ void __launch_bounds__(256)
transpose_v_kernel(const bf16* __restrict__ src, int ld, int col0, size_t sls, bf16* __restrict__ dst, int ld_dst,
                   size_t dls, int N, int G, int hd, int grows) {
    // one block = 64 tokens of one (batch, kv head): 16-byte loads along head_dim, token PAIRS packed into 32-bit words
    // in shared memory (pitch 33 words), 128-byte coalesced stores along the token axis.
    __shared__ uint32_t tile[128 * 33];
    pdl_trigger();
    pdl_wait();
    const int l = blockIdx.z;
    const int bg = blockIdx.y;                 // b * G + g
    const int b = bg / G, g = bg % G;
    const int n0 = blockIdx.x * 64;
    const int nvec = hd >> 3;                  // 16-byte vectors per token row
    const bf16* sp = src + l * sls + static_cast<size_t>(b) * N * ld + col0 + g * hd;
    for (int idx = threadIdx.x; idx < 32 * nvec; idx += blockDim.x) {
        const int p = idx / nvec, dv = idx % nvec;
        const int t0 = n0 + 2 * p;
        uint4 a = make_uint4(0, 0, 0, 0), c = a;
        if (t0 < N) a = *reinterpret_cast<const uint4*>(sp + static_cast<size_t>(t0) * ld + dv * 8);
        if (t0 + 1 < N) c = *reinterpret_cast<const uint4*>(sp + static_cast<size_t>(t0 + 1) * ld + dv * 8);
        const uint32_t av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            tile[(dv * 8 + 2 * j) * 33 + p] = (av[j] & 0xffffu) | (cv[j] << 16);
            tile[(dv * 8 + 2 * j + 1) * 33 + p] = (av[j] >> 16) | (cv[j] & 0xffff0000u);
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int n = n0 + 2 * lane;
    if (n < N) {    // token pair (n, n+1); for odd N the last pair spills one zero into the row padding
        for (int d = threadIdx.x >> 5; d < hd; d += blockDim.x >> 5)
            *reinterpret_cast<uint32_t*>(dst + l * dls + (static_cast<size_t>(bg) * grows + d) * ld_dst + n) = tile[d * 33 + lane];
    }
}

cudaError_t transpose_v(const bf16* src, int ld, int col0, size_t src_layer_stride, bf16* dst, int ld_dst,
                        size_t dst_layer_stride, int B, int N, int G, int hd, int grows, int layers, cudaStream_t s) {
    if (hd % 8 != 0 || hd > 128 || (ld_dst & 1) || ld_dst < ((N + 1) & ~1) || (ld & 7) || (col0 & 7)) return cudaErrorInvalidValue;
    const dim3 grid((N + 63) / 64, B * G, layers);
    return launch_k(transpose_v_kernel, grid, dim3(256), 0, s, src, ld, col0, src_layer_stride, dst, ld_dst, dst_layer_stride, N, G, hd,
                    grows);
}

__global__ void fill_ones_row_kernel(bf16* __restrict__ dst, int ld_dst, size_t dls, int n_cols, int hd, int grows) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_cols) return;
    dst[blockIdx.z * dls + (static_cast<size_t>(blockIdx.y) * grows + hd) * ld_dst + n] = __float2bfloat16(1.0f);
}

cudaError_t fill_ones_row(bf16* dst, int ld_dst, size_t dst_layer_stride, int groups, int n_cols, int hd, int grows, int layers,
                          cudaStream_t s) {
    const dim3 grid((n_cols + 255) / 256, groups, layers);
    fill_ones_row_kernel<<<grid, 256, 0, s>>>(dst, ld_dst, dst_layer_stride, n_cols, hd, grows);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// unpatchify (token feature order (ph,pw,c_out), model.py:753-754) + keep first C of 2C channels (:859-861)
// + CFG on channels 0..2 only (:904-913).
__global__ void unpatchify_cfg_kernel(const bf16* __restrict__ tok, bf16* __restrict__ v_out, int n, int C, int Hh,
                                      int Ww, int O, float cfg_scale, int eol) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over n * C * Hh * Ww
    const int total = n * C * Hh * Ww;
    if (idx >= total) return;
    const int x = idx % Ww, y = (idx / Ww) % Hh, c = (idx / (Ww * Hh)) % C, s = idx / (Ww * Hh * C);
    const int Wt = (Ww >> 1) + eol, N = (Hh >> 1) * Wt;      // eol = 1: the [eol] column is dropped (lumina_t2i model.py:745-755)
    const int t = (y >> 1) * Wt + (x >> 1);
    const int Cout = O / 4;
    const int f = ((y & 1) * 2 + (x & 1)) * Cout + c;
    const float cond = __bfloat162float(tok[(static_cast<size_t>(s) * N + t) * O + f]);
    const float unc = __bfloat162float(tok[(static_cast<size_t>(s + n) * N + t) * O + f]);
    const size_t plane = static_cast<size_t>(C) * Hh * Ww;
    const size_t o = static_cast<size_t>(c) * Hh * Ww + static_cast<size_t>(y) * Ww + x;
    if (c < 3) {
        const float g = bf16_round(unc + bf16_round(cfg_scale * bf16_round(cond - unc)));
        v_out[s * plane + o] = __float2bfloat16_rn(g);
        v_out[(s + n) * plane + o] = __float2bfloat16_rn(g);
    } else {
        v_out[s * plane + o] = __float2bfloat16_rn(cond);
        v_out[(s + n) * plane + o] = __float2bfloat16_rn(unc);
    }
}

// unpatchify without guidance (NextDiT.forward, model.py:858-863): every row keeps the first C of its 2C output channels
__global__ void unpatchify_plain_kernel(const bf16* __restrict__ tok, bf16* __restrict__ v_out, int n, int C, int Hh, int Ww, int O,
                                        int eol) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over n * C * Hh * Ww
    if (idx >= n * C * Hh * Ww) return;
    const int x = idx % Ww, y = (idx / Ww) % Hh, c = (idx / (Ww * Hh)) % C, s = idx / (Ww * Hh * C);
    const int Wt = (Ww >> 1) + eol, N = (Hh >> 1) * Wt;
    const int t = (y >> 1) * Wt + (x >> 1);
    const int f = ((y & 1) * 2 + (x & 1)) * (O / 4) + c;
    v_out[idx] = tok[(static_cast<size_t>(s) * N + t) * O + f];
}

cudaError_t unpatchify_plain(const bf16* tok, bf16* v_out, int n, int C, int Hh, int Ww, int O, int eol, cudaStream_t s) {
    const int total = n * C * Hh * Ww;
    unpatchify_plain_kernel<<<(total + 255) / 256, 256, 0, s>>>(tok, v_out, n, C, Hh, Ww, O, eol);
    return cudaGetLastError();
}

cudaError_t unpatchify_cfg(const bf16* tok, bf16* v_out, int n, int C, int Hh, int Ww, int O, float cfg_scale, int eol,
                           cudaStream_t s) {
    const int total = n * C * Hh * Ww;
    unpatchify_cfg_kernel<<<(total + 255) / 256, 256, 0, s>>>(tok, v_out, n, C, Hh, Ww, O, cfg_scale, eol);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Mixture-of-experts helpers (Next-DiT-MoE models.py:459-477, models2.py:459-506).
// Token gate: logits = bf16(u . Wg^T) (E <= 8), top-2 (ties: lower expert index first), weights = bf16(softmax over the two
// selected logits, fp32); wtok[row][e] = weight of expert e for this token, 0 if not selected.
__global__ void __launch_bounds__(ROW_WARPS * 32)
moe_space_gate_kernel(const bf16* __restrict__ u, const bf16* __restrict__ Wg, bf16* __restrict__ wtok, int M, int D, int E) {
    const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = lane * 8; k < D; k += 256) {
        float x[8];
        load8(u + static_cast<size_t>(row) * D + k, x);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (e < E) {
                float w[8];
                load8(Wg + static_cast<size_t>(e) * D + k, w);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[e] = fmaf(x[j], w[j], acc[e]);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = bf16_round(warp_sum(acc[e]));
    if (lane == 0) {
        int i0 = 0, i1 = -1;
        for (int e = 1; e < E; ++e) if (acc[e] > acc[i0]) i0 = e;
        for (int e = 0; e < E; ++e) if (e != i0 && (i1 < 0 || acc[e] > acc[i1])) i1 = e;
        const float e1 = expf(acc[i1] - acc[i0]);           // softmax over (l0, l1), l0 >= l1
        const float w0 = bf16_round(1.0f / (1.0f + e1)), w1 = bf16_round(e1 / (1.0f + e1));
        for (int e = 0; e < E; ++e)
            wtok[static_cast<size_t>(row) * E + e] = __float2bfloat16_rn(e == i0 ? w0 : (e == i1 ? w1 : 0.f));
    }
}

cudaError_t moe_space_gate(const bf16* u, const bf16* Wg, bf16* wtok, int M, int D, int E, cudaStream_t s) {
    if (E < 2 || E > 8 || D % 8 != 0) return cudaErrorInvalidValue;
    moe_space_gate_kernel<<<(M + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, s>>>(u, Wg, wtok, M, D, E);
    return cudaGetLastError();
}

// out = sum over experts in index order, accumulated in bf16 like ``results[idx] += w * expert(x[idx])`` on a zero tensor:
// r = bf16(r + bf16(w_e * o_e)) for every selected expert.  wtok != nullptr: per-token weights [M, E] (0 = not selected);
// else the E buffers are the selected experts of a time-gated layer with the uniform weights uw[0..E).
__global__ void moe_combine_kernel(const bf16* __restrict__ oe, size_t estride, int E, const bf16* __restrict__ wtok,
                                   const float* __restrict__ uw, bf16* __restrict__ out, size_t count8, int D) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;     // one 8-element vector
    if (i >= count8) return;
    const size_t row = (i * 8) / D;
    float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int e = 0; e < E; ++e) {
        const float w = wtok != nullptr ? __bfloat162float(wtok[row * E + e]) : uw[e];
        if (wtok != nullptr && w == 0.f) continue;
        float o[8];
        load8(oe + e * estride + i * 8, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = bf16_round(r[j] + bf16_round(w * o[j]));
    }
    store8(out + i * 8, r);
}

cudaError_t moe_combine(const bf16* oe, size_t estride, int E, const bf16* wtok, const float* uniform_w, bf16* out, int M, int D,
                        cudaStream_t s) {
    if (E < 1 || E > 8 || D % 8 != 0) return cudaErrorInvalidValue;
    if (wtok == nullptr && uniform_w == nullptr) return cudaErrorInvalidValue;
    const size_t count8 = static_cast<size_t>(M) * D / 8;
    moe_combine_kernel<<<static_cast<unsigned>((count8 + 255) / 256), 256, 0, s>>>(oe, estride, E, wtok, uniform_w, out, count8, D);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Token-routed experts as a grouped GEMM (MoeLayer.forward, Next-DiT-MoE models1.py:459-477: ``results[idx] += w * expert(x[idx])``
// over the tokens that selected the expert).  The gate kernel left wtok[M][E] (non-zero = selected, exactly two per token).
//   route_count : cnt[e] = tokens that selected expert e
//   route_scan  : segment of expert e in the gathered buffers = rows [off[e], off[e] + cnt[e]), padded to 256 rows (GEMM tile);
//                 cntp[e] = padded count (the GEMM's device-side row window), cursor[e] = 0
//   route_gather: every token copies its row of u into the segments of its two experts (slot order inside a segment is whatever
//                 the atomics give - each gathered row is computed on its own, so the result does not depend on it)
//   combine     : out[token] = bf16(bf16(0 + bf16(w_a o_a)) + bf16(w_b o_b)), a < b the two selected experts - the same
//                 arithmetic, in the same (expert index) order, as the dense moe_combine
__global__ void moe_route_count_kernel(const bf16* __restrict__ wtok, int* __restrict__ cnt, int M, int E) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    for (int e = 0; e < E; ++e)
        if (__bfloat162float(wtok[static_cast<size_t>(row) * E + e]) != 0.f) atomicAdd(cnt + e, 1);
}
__global__ void moe_route_scan_kernel(int* __restrict__ cnt, int* __restrict__ cntp, int* __restrict__ off, int* __restrict__ cursor, int E) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int o = 0;
    for (int e = 0; e < E; ++e) {
        const int c = cnt[e], cp = (c + 255) / 256 * 256;
        off[e] = o; cntp[e] = cp; cursor[e] = 0;
        o += cp;
        cnt[e] = 0;                       // ready for the next layer
    }
    off[E] = o;
}
__global__ void __launch_bounds__(ROW_WARPS * 32)
moe_route_gather_kernel(const bf16* __restrict__ u, const bf16* __restrict__ wtok, const int* __restrict__ off, int* __restrict__ cursor,
                        int* __restrict__ pos, bf16* __restrict__ u_perm, int M, int D, int E) {
    const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    int p[2] = {-1, -1};
    if (lane == 0) {
        int k = 0;
        for (int e = 0; e < E && k < 2; ++e)
            if (__bfloat162float(wtok[static_cast<size_t>(row) * E + e]) != 0.f) p[k++] = off[e] + atomicAdd(cursor + e, 1);
        pos[2 * row] = p[0];
        pos[2 * row + 1] = p[1];
    }
    p[0] = __shfl_sync(0xffffffffu, p[0], 0);
    p[1] = __shfl_sync(0xffffffffu, p[1], 0);
    const uint4* src = reinterpret_cast<const uint4*>(u + static_cast<size_t>(row) * D);
    for (int v = lane; v < D / 8; v += 32) {
        const uint4 x = src[v];
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (p[k] >= 0) reinterpret_cast<uint4*>(u_perm + static_cast<size_t>(p[k]) * D)[v] = x;
    }
}
__global__ void moe_combine_routed_kernel(const bf16* __restrict__ o_perm, const int* __restrict__ pos, const bf16* __restrict__ wtok,
                                          bf16* __restrict__ out, size_t count8, int D, int E) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;     // one 8-element vector
    if (i >= count8) return;
    const size_t row = (i * 8) / D;
    const int col = static_cast<int>(i * 8 - row * D);
    float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int k = 0;
    for (int e = 0; e < E && k < 2; ++e) {
        const float w = __bfloat162float(wtok[row * E + e]);
        if (w == 0.f) continue;
        const int p = pos[2 * row + k];
        ++k;
        float o[8];
        load8(o_perm + static_cast<size_t>(p) * D + col, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = bf16_round(r[j] + bf16_round(w * o[j]));
    }
    store8(out + i * 8, r);
}

cudaError_t moe_route(const bf16* u, const bf16* wtok, int* cnt, int* cntp, int* off, int* cursor, int* pos, bf16* u_perm, int M, int D, int E,
                      cudaStream_t s) {
    if (E < 2 || E > 8 || D % 8 != 0) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(cnt, 0, sizeof(int) * E, s);     // also after a forward that failed half way
    if (e != cudaSuccess) return e;
    moe_route_count_kernel<<<(M + 255) / 256, 256, 0, s>>>(wtok, cnt, M, E);
    moe_route_scan_kernel<<<1, 32, 0, s>>>(cnt, cntp, off, cursor, E);
    moe_route_gather_kernel<<<(M + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, s>>>(u, wtok, off, cursor, pos, u_perm, M, D, E);
    return cudaGetLastError();
}

cudaError_t moe_combine_routed(const bf16* o_perm, const int* pos, const bf16* wtok, bf16* out, int M, int D, int E, cudaStream_t s) {
    if (E < 2 || E > 8 || D % 8 != 0) return cudaErrorInvalidValue;
    const size_t count8 = static_cast<size_t>(M) * D / 8;
    moe_combine_routed_kernel<<<static_cast<unsigned>((count8 + 255) / 256), 256, 0, s>>>(o_perm, pos, wtok, out, count8, D, E);
    return cudaGetLastError();
}

// One thread per layer: top-2 of E gate logits (ties: lower expert index first), softmax over the two, bf16-rounded weights,
// stored in ascending expert order (the reference accumulates ``results += w * expert(x)`` in expert-index order).
__global__ void moe_time_select_kernel(const float* __restrict__ logits, int L, int E, int* __restrict__ sel, float* __restrict__ w) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const float* lg = logits + static_cast<size_t>(l) * E;
    int i0 = 0, i1 = -1;
    for (int e = 1; e < E; ++e) if (lg[e] > lg[i0]) i0 = e;
    for (int e = 0; e < E; ++e) if (e != i0 && (i1 < 0 || lg[e] > lg[i1])) i1 = e;
    const float e1 = expf(lg[i1] - lg[i0]);
    const float w0 = bf16_round(1.0f / (1.0f + e1)), w1 = bf16_round(e1 / (1.0f + e1));
    if (i0 < i1) { sel[2 * l] = i0; sel[2 * l + 1] = i1; w[2 * l] = w0; w[2 * l + 1] = w1; }
    else { sel[2 * l] = i1; sel[2 * l + 1] = i0; w[2 * l] = w1; w[2 * l + 1] = w0; }
}

cudaError_t moe_time_select(const float* logits, int L, int E, int* sel, float* w, cudaStream_t s) {
    if (E < 2 || E > 8 || L < 1) return cudaErrorInvalidValue;
    moe_time_select_kernel<<<(L + 63) / 64, 64, 0, s>>>(logits, L, E, sel, w);
    return cudaGetLastError();
}

// y_out = bf16(y_in + bf16(dt * v))   (torchdiffeq fixed-grid update in the state dtype)
__global__ void axpy_bf16_kernel(bf16* __restrict__ y_out, const bf16* __restrict__ y_in, const bf16* __restrict__ v,
                                 float dt, size_t count) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float d = bf16_round(dt * __bfloat162float(v[i]));
    y_out[i] = __float2bfloat16_rn(__bfloat162float(y_in[i]) + d);
}

// torchdiffeq's fixed-grid rk4 (rk4_alt_step_func, the 3/8 rule) on a bf16 state, every tensor op rounded to bf16 like
// PyTorch does (dt is a 0-dim tensor and is rounded to bf16 by type promotion, Python scalars enter in fp32):
//   stage 1: out = y + (dt * k1) * (1/3)
//   stage 2: out = y + dt * (k2 - k1 * (1/3))
//   stage 3: out = y + dt * (k1 - k2 + k3)
//   stage 4: out = y + (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
__global__ void rk4_stage_kernel(int stage, bf16* __restrict__ out, const bf16* __restrict__ y, const bf16* __restrict__ k1,
                                 const bf16* __restrict__ k2, const bf16* __restrict__ k3, const bf16* __restrict__ k4, float dt,
                                 size_t count) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float third = 1.0f / 3.0f;          // float(1 / 3) as PyTorch passes the Python scalar
    const float yv = __bfloat162float(y[i]), a = __bfloat162float(k1[i]);
    float d;
    if (stage == 1) {
        d = bf16_round(bf16_round(dt * a) * third);
    } else if (stage == 2) {
        d = bf16_round(dt * bf16_round(__bfloat162float(k2[i]) - bf16_round(a * third)));
    } else if (stage == 3) {
        d = bf16_round(dt * bf16_round(bf16_round(a - __bfloat162float(k2[i])) + __bfloat162float(k3[i])));
    } else {
        const float s23 = bf16_round(__bfloat162float(k2[i]) + __bfloat162float(k3[i]));
        const float acc = bf16_round(bf16_round(a + bf16_round(3.0f * s23)) + __bfloat162float(k4[i]));
        d = bf16_round(bf16_round(acc * dt) * 0.125f);
    }
    out[i] = __float2bfloat16_rn(yv + d);
}

cudaError_t rk4_stage(int stage, bf16* out, const bf16* y, const bf16* k1, const bf16* k2, const bf16* k3, const bf16* k4, float dt,
                      size_t count, cudaStream_t s) {
    if (stage < 1 || stage > 4) return cudaErrorInvalidValue;
    rk4_stage_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, s>>>(stage, out, y, k1, k2, k3, k4, dt, count);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// transport.Sampler.sample_sde on a bf16 state (transport.py:285-344, integrators.py:5-76; velocity model, Linear path).  Every
// PyTorch tensor op of the reference rounds to bf16; the t-dependent scalars (computed by the caller with the same ops on a
// one-element bf16 tensor) arrive as floats that hold bf16 values: ratio = alpha_t / d_alpha_t, var, diffusion, sqrt(2 diffusion).
//   score      = ((ratio * v) - x) / var                         (path.py get_score_from_velocity)
//   sde_drift  = v + diffusion * score
// mode 0  Euler-Maruyama:  out = (x + drift * dt) + sqrt2d * (w * sqrt_dt)
// mode 1  Heun, noise:     out = x + sqrt2d * (w * sqrt_dt)                                  (xhat)
// mode 2  Heun, stage 1:   k = drift(x = xhat, v);  out = xhat + dt * k                       (k1 -> kout, xp -> out)
// mode 3  Heun, stage 2:   k2 = drift(x = xp, v);   out = xhat + (0.5 dt) * (k1 + k2)         (xp in `x`, xhat in `a`, k1 in `kin`)
struct SdeCoef { float ratio, var, diffusion, sqrt2d, dt, sqrt_dt, half_dt; };
__global__ void sde_step_kernel(int mode, bf16* __restrict__ out, bf16* __restrict__ kout, const bf16* __restrict__ x,
                                const bf16* __restrict__ v, const bf16* __restrict__ w, const bf16* __restrict__ a,
                                const bf16* __restrict__ kin, SdeCoef c, size_t count) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float xv = __bfloat162float(x[i]);
    if (mode == 1) {
        const float n = bf16_round(c.sqrt2d * bf16_round(__bfloat162float(w[i]) * c.sqrt_dt));
        out[i] = __float2bfloat16_rn(xv + n);
        return;
    }
    const float vv = __bfloat162float(v[i]);
    const float sc = bf16_round(bf16_round(bf16_round(c.ratio * vv) - xv) / c.var);
    const float drift = bf16_round(vv + bf16_round(c.diffusion * sc));
    if (mode == 0) {
        const float x1 = bf16_round(xv + bf16_round(drift * c.dt));
        const float n = bf16_round(c.sqrt2d * bf16_round(__bfloat162float(w[i]) * c.sqrt_dt));
        out[i] = __float2bfloat16_rn(x1 + n);
    } else if (mode == 2) {
        kout[i] = __float2bfloat16_rn(drift);
        out[i] = __float2bfloat16_rn(xv + bf16_round(c.dt * drift));
    } else {
        const float ks = bf16_round(__bfloat162float(kin[i]) + drift);
        out[i] = __float2bfloat16_rn(__bfloat162float(a[i]) + bf16_round(c.half_dt * ks));
    }
}

cudaError_t sde_step(int mode, bf16* out, bf16* kout, const bf16* x, const bf16* v, const bf16* w, const bf16* a, const bf16* kin,
                     float ratio, float var, float diffusion, float sqrt2d, float dt, float sqrt_dt, float half_dt, size_t count,
                     cudaStream_t s) {
    if (mode < 0 || mode > 3) return cudaErrorInvalidValue;
    const SdeCoef c{ratio, var, diffusion, sqrt2d, dt, sqrt_dt, half_dt};
    sde_step_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, s>>>(mode, out, kout, x, v, w, a, kin, c, count);
    return cudaGetLastError();
}

cudaError_t axpy_bf16(bf16* y_out, const bf16* y_in, const bf16* v, float dt, size_t count, cudaStream_t s) {
    axpy_bf16_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, s>>>(y_out, y_in, v, dt, count);
    return cudaGetLastError();
}

}  // namespace ndit
