// VAE-decode end of the Lumina-Next-T2I sampling path (include/ndit_vae.h): diffusers' AutoencoderKL.decode behind
//   samples = vae.decode(samples / factor).sample                    (lumina_next_t2i/sample.py:238, under autocast(bf16), :173)
// restated from the published diffusers implementation (models/autoencoders/autoencoder_kl.py AutoencoderKL.decode / _decode,
// models/autoencoders/vae.py Decoder.forward, models/unets/unet_2d_blocks.py UNetMidBlock2D / UpDecoderBlock2D,
// models/resnet.py ResnetBlock2D, models/upsampling.py Upsample2D, models/attention_processor.py Attention); diffusers is not in
// /root/reference nor in the authoring image, so this end is "parity unpinned" (checked against oracle/vae_oracle.py only).
//
// Data layout in HBM: every activation is NHWC bf16 with a one-pixel zero border, [batch][H + 2][W + 2][C], i.e. a row-major
// [batch * (H+2) * (W+2), C] matrix.  With the border in memory a 3x3 convolution with padding 1 is nine accumulating GEMMs over
// the SAME matrix shifted by (dy * (W+2) + dx) rows: the tcgen05 CTA-pair GEMM of gemm_tcgen05.cu runs it with K = 9 * Cin, its
// TMA producer only offsets the row coordinate of the A tile per tap (GemmExt), W packed [Cout][tap][Cin].  The epilogue adds the
// bias, rounds to bf16 (the autocast conv output), adds the ResnetBlock2D / Attention residual and stores zeros on border rows so
// the next convolution and the next GroupNorm statistics see the padding they expect.  About 10.3 TFLOP per 1024 x 1024 image
// (conv 9.7, mid-block attention 0.6), tensor-core bound; GroupNorm is a two-pass row-wise job (statistics, then normalise + SiLU
// + bf16 into the buffer the next convolution reads), HBM bound.
// Mid-block attention (one head of head_dim C = 512 over all H*W latent pixels): q, k as GEMMs; V^T = Wv . Xn^T directly as a GEMM
// with the weight as the A operand; scores S = Q K^T in fp32 for a chunk of query rows sized to stay in L2 (64 MB), a row softmax
// that masks the border pixels, O = P V^T^T + bv (rows of P sum to one, so the value bias moves behind the product), out
// projection with the residual in the epilogue.
// Rounding points follow the reference's autocast: conv / linear outputs bf16 (fp32 accumulate), GroupNorm + SiLU fp32 with one
// bf16 rounding at the next conv's input, softmax fp32 with bf16 probabilities, residual sums bf16.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/ndit.h"
#include "../../include/ndit_vae.h"
#include "kernels.h"
#include "ptx.cuh"

using namespace ndit;

namespace {

constexpr int GN_PPC = 512;        // pixels per block of the statistics pass (at most)
constexpr float GN_EPS = 1e-6f;    // Decoder / ResnetBlock2D / Attention all build their GroupNorm with eps = 1e-6

__device__ __forceinline__ void ld8(const bf16* p, float* f) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void st8(bf16* p, const float* f) {
    uint4 u;
    u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
}

// ---------------------------------------------------------------------------------------------- weight packing
// dst[(o * taps + t) * cin + i] = bf16(src[(o * cin + i) * taps + t])     ([Cout, Cin, kh, kw] -> [Cout][tap][Cin])
__global__ void vae_pack_conv_kernel(bf16* __restrict__ dst, const void* __restrict__ src, int src_f32, int cout, int cin, int taps) {
    const size_t total = static_cast<size_t>(cout) * cin * taps;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % cin);
        const size_t r = i / cin;
        const int t = static_cast<int>(r % taps);
        const size_t o = r / taps;
        const size_t si = (o * cin + ci) * taps + t;
        dst[i] = src_f32 ? __float2bfloat16_rn(static_cast<const float*>(src)[si]) : static_cast<const bf16*>(src)[si];
    }
}
// dst[(t * cin + i) * cout + o] = bf16(src[(o * cin + i) * taps + t])    (conv_in: [K][Cout], consecutive threads read consecutive channels)
__global__ void vae_pack_conv_t_kernel(bf16* __restrict__ dst, const void* __restrict__ src, int src_f32, int cout, int cin, int taps) {
    const size_t total = static_cast<size_t>(cout) * cin * taps;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t o = i % cout, r = i / cout;
        const size_t ci = r % cin, t = r / cin;
        const size_t si = (o * cin + ci) * taps + t;
        dst[i] = src_f32 ? __float2bfloat16_rn(static_cast<const float*>(src)[si]) : static_cast<const bf16*>(src)[si];
    }
}
__global__ void vae_copy_vec_kernel(void* __restrict__ dst, int dst_f32, const void* __restrict__ src, int src_f32, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = src_f32 ? static_cast<const float*>(src)[i] : __bfloat162float(static_cast<const bf16*>(src)[i]);
    if (dst_f32) static_cast<float*>(dst)[i] = v;
    else static_cast<bf16*>(dst)[i] = __float2bfloat16_rn(v);
}

// ---------------------------------------------------------------------------------------------- conv_in (+ post_quant_conv)
// AutoencoderKL._decode: z = post_quant_conv(z) (1x1, Cz -> Cz);  Decoder.forward: sample = conv_in(z) (3x3, Cz -> Cout, padding 1).
// One block per padded output pixel, one thread per 8 output channels; Cz * 9 inputs are tiny, so this is a CUDA-core kernel.
__global__ void vae_conv_in_kernel(const bf16* __restrict__ z, const bf16* __restrict__ pqw, const bf16* __restrict__ pqb,
                                   const bf16* __restrict__ w, const bf16* __restrict__ bias, bf16* __restrict__ out, int Hh, int Ww,
                                   int Cz, int Cout) {
    __shared__ float pq[9 * 16];
    const int Wp = Ww + 2, Hp = Hh + 2;
    const int p = blockIdx.x;
    const int b = p / (Hp * Wp), rr = p - b * (Hp * Wp), y = rr / Wp, x = rr - y * Wp;
    bf16* orow = out + static_cast<size_t>(p) * Cout;
    if (y == 0 || y == Hp - 1 || x == 0 || x == Wp - 1) {
        const float zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int o0 = threadIdx.x * 8; o0 < Cout; o0 += blockDim.x * 8) st8(orow + o0, zero);
        return;
    }
    for (int t = threadIdx.x; t < 9 * Cz; t += blockDim.x) {
        const int tap = t / Cz, o = t - tap * Cz;
        const int iy = y - 1 + tap / 3 - 1, ix = x - 1 + tap % 3 - 1;
        float v = 0.f;                                          // conv_in pads post_quant_conv's OUTPUT with zeros
        if (iy >= 0 && iy < Hh && ix >= 0 && ix < Ww) {
            float acc = 0.f;
            for (int c = 0; c < Cz; ++c)
                acc += __bfloat162float(z[((static_cast<size_t>(b) * Cz + c) * Hh + iy) * Ww + ix]) * __bfloat162float(pqw[o * Cz + c]);
            v = bf16_round(acc + __bfloat162float(pqb[o]));
        }
        pq[t] = v;
    }
    __syncthreads();
    const int K = 9 * Cz;           // w is [K][Cout]
    for (int o0 = threadIdx.x * 8; o0 < Cout; o0 += blockDim.x * 8) {
        float acc[8];
        ld8(bias + o0, acc);
        for (int k = 0; k < K; ++k) {
            float wv[8];
            ld8(w + static_cast<size_t>(k) * Cout + o0, wv);
            const float a = pq[k];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(a, wv[e], acc[e]);
        }
        st8(orow + o0, acc);
    }
}

// ---------------------------------------------------------------------------------------------- GroupNorm
// pass 1: per (image, chunk of pixels, 4-channel slice) partial sum / sum of squares (border pixels are zero and add nothing)
__global__ void __launch_bounds__(256) vae_gn_stats_kernel(const bf16* __restrict__ x, float4* __restrict__ partial, int P, int C, int chunks,
                                                            int ppc) {
    __shared__ float4 sm[256];
    const int tpp = C >> 3, ppb = 256 / tpp;
    const int c8 = threadIdx.x % tpp, pl = threadIdx.x / tpp;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int p0 = chunk * ppc, p1 = min(P, p0 + ppc);
    const bf16* xb = x + static_cast<size_t>(b) * P * C + c8 * 8;
    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
    auto add = [&](const uint4& u) {
        const float2 a = unpack_bf16(u.x), bq = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
        s0 += (a.x + a.y) + (bq.x + bq.y);
        q0 += (a.x * a.x + a.y * a.y) + (bq.x * bq.x + bq.y * bq.y);
        s1 += (c.x + c.y) + (d.x + d.y);
        q1 += (c.x * c.x + c.y * c.y) + (d.x * d.x + d.y * d.y);
    };
    int p = p0 + pl;
    for (; p + 3 * ppb < p1; p += 4 * ppb) {          // four independent 16-byte loads in flight per thread
        const uint4 u0 = *reinterpret_cast<const uint4*>(xb + static_cast<size_t>(p) * C);
        const uint4 u1 = *reinterpret_cast<const uint4*>(xb + static_cast<size_t>(p + ppb) * C);
        const uint4 u2 = *reinterpret_cast<const uint4*>(xb + static_cast<size_t>(p + 2 * ppb) * C);
        const uint4 u3 = *reinterpret_cast<const uint4*>(xb + static_cast<size_t>(p + 3 * ppb) * C);
        add(u0); add(u1); add(u2); add(u3);
    }
    for (; p < p1; p += ppb) add(*reinterpret_cast<const uint4*>(xb + static_cast<size_t>(p) * C));
    sm[threadIdx.x] = make_float4(s0, q0, s1, q1);
    __syncthreads();
    if (threadIdx.x < tpp) {
        float4 a = sm[threadIdx.x];
        for (int i = 1; i < ppb; ++i) {
            const float4 t = sm[i * tpp + threadIdx.x];
            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        partial[(static_cast<size_t>(b) * chunks + chunk) * tpp + threadIdx.x] = a;
    }
}
// pass 1b: one block per (group, image): double accumulation over the chunk partials, then the per-channel affine of the group
//   ab[b][c] = (rstd * gamma[c], beta[c] - mean * rstd * gamma[c])
__global__ void __launch_bounds__(256) vae_gn_finalize_kernel(const float4* __restrict__ partial, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float2* __restrict__ ab, int C, int chunks, int G,
                                                              double inv_count, float eps) {
    __shared__ double rs[8], rq[8];
    __shared__ float2 st;
    const int g = blockIdx.x, b = blockIdx.y;
    const int tpp = C >> 3, cpg = C / G, hv0 = g * cpg / 4, nhv = cpg / 4;
    double s = 0.0, q = 0.0;
    for (int chunk = threadIdx.x; chunk < chunks; chunk += 256) {
        const float4* row = partial + (static_cast<size_t>(b) * chunks + chunk) * tpp;
        for (int i = 0; i < nhv; ++i) {
            const int hv = hv0 + i;
            const float4 t = row[hv >> 1];
            if (hv & 1) { s += t.z; q += t.w; } else { s += t.x; q += t.y; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tq = 0.0;
        for (int i = 0; i < 8; ++i) { ts += rs[i]; tq += rq[i]; }
        const double mean = ts * inv_count;
        double var = tq * inv_count - mean * mean;
        if (var < 0.0) var = 0.0;
        st = make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))));
    }
    __syncthreads();
    if (threadIdx.x < cpg) {
        const int c = g * cpg + threadIdx.x;
        const float a = st.y * gamma[c];
        ab[static_cast<size_t>(b) * C + c] = make_float2(a, beta[c] - st.x * a);
    }
}
// pass 2: y = bf16(act(x * a[c] + b[c])), act = SiLU or identity, zeros on the border.  C / 8 divides the block size (C is 128,
// 256 or 512), so a thread keeps the same 8 channels for all of its GN_VPT vectors: their affine lives in registers and consecutive
// vectors of a thread are 256 / (C / 8) pixels apart.
constexpr int GN_VPT = 4;
__global__ void __launch_bounds__(256, 3) vae_gn_apply_kernel(const bf16* __restrict__ x, const float2* __restrict__ ab_g, bf16* __restrict__ y,
                                                              int P, int C, int Hp, int Wp, int silu) {
    const int b = blockIdx.y;
    const unsigned tpp = C >> 3, ppi = 256u / tpp;                    // pixels covered by one pass of the block
    const unsigned c8 = threadIdx.x % tpp;
    const unsigned p0 = blockIdx.x * (ppi * GN_VPT) + threadIdx.x / tpp;
    float2 ab[8];
    {
        const float4* src = reinterpret_cast<const float4*>(ab_g + static_cast<size_t>(b) * C + c8 * 8);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float4 t = src[e];
            ab[2 * e] = make_float2(t.x, t.y);
            ab[2 * e + 1] = make_float2(t.z, t.w);
        }
    }
    const bf16* xb = x + static_cast<size_t>(b) * P * C + c8 * 8;
    bf16* yb = y + static_cast<size_t>(b) * P * C + c8 * 8;
    uint4 u[GN_VPT];
#pragma unroll
    for (int i = 0; i < GN_VPT; ++i) {
        const unsigned p = p0 + i * ppi;
        u[i] = p < static_cast<unsigned>(P) ? *reinterpret_cast<const uint4*>(xb + static_cast<size_t>(p) * C) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < GN_VPT; ++i) {
        const unsigned p = p0 + i * ppi;
        if (p >= static_cast<unsigned>(P)) break;
        const unsigned py = p / static_cast<unsigned>(Wp), px = p - py * static_cast<unsigned>(Wp);
        float v[8];
        if (py == 0 || py == static_cast<unsigned>(Hp - 1) || px == 0 || px == static_cast<unsigned>(Wp - 1)) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = 0.f;
        } else {
            const float2 a = unpack_bf16(u[i].x), bq = unpack_bf16(u[i].y), c = unpack_bf16(u[i].z), d = unpack_bf16(u[i].w);
            v[0] = a.x; v[1] = a.y; v[2] = bq.x; v[3] = bq.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float n = fmaf(v[e], ab[e].x, ab[e].y);
                // ex2 + rcp.  (h + h * tanh.approx(h), h = n / 2, saves a MUFU and 15 % of this kernel but loses relative precision on the
                // negative tail - the sum cancels - and moved the 1024 x 1024 decode from 1.72e-2 to 1.96e-2 of the fp32 result: not taken.)
                v[e] = silu ? __fdividef(n, 1.0f + __expf(-n)) : n;
            }
        }
        st8(yb + static_cast<size_t>(p) * C, v);
    }
}

// ---------------------------------------------------------------------------------------------- Upsample2D (nearest, x2)
// F.interpolate(x, scale_factor=2.0, mode="nearest") into the padded buffer of the doubled resolution
__global__ void __launch_bounds__(256) vae_upsample_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int B, int Hh, int Ww, int C) {
    const int tpp = C >> 3, Hp2 = 2 * Hh + 2, Wp2 = 2 * Ww + 2, Wp = Ww + 2;
    const long long total = static_cast<long long>(B) * Hp2 * Wp2 * tpp;
    const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= total) return;
    const int c8 = static_cast<int>(idx % tpp);
    long long p = idx / tpp;
    const int X = static_cast<int>(p % Wp2);
    p /= Wp2;
    const int Y = static_cast<int>(p % Hp2), b = static_cast<int>(p / Hp2);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (Y > 0 && Y < Hp2 - 1 && X > 0 && X < Wp2 - 1) {
        const int sy = (Y - 1) / 2 + 1, sx = (X - 1) / 2 + 1;
        v = *reinterpret_cast<const uint4*>(src + ((static_cast<size_t>(b) * (Hh + 2) + sy) * Wp + sx) * C + c8 * 8);
    }
    *reinterpret_cast<uint4*>(dst + static_cast<size_t>(idx) * 8) = v;
}

// ---------------------------------------------------------------------------------------------- attention softmax
// P[r, :] = bf16(softmax(scale * S[r, :])) over the columns whose colmask is 1 (interior pixels); other columns get 0
__global__ void __launch_bounds__(256) vae_softmax_kernel(const float* __restrict__ S, bf16* __restrict__ Pm, const uint8_t* __restrict__ colmask,
                                                          int ld, float scale) {
    __shared__ float red[8];
    __shared__ float bcast;
    const float* s = S + static_cast<size_t>(blockIdx.x) * ld;
    bf16* pr = Pm + static_cast<size_t>(blockIdx.x) * ld;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mx = -INFINITY;
    for (int j = threadIdx.x * 4; j < ld; j += 1024) {
        const float4 v = *reinterpret_cast<const float4*>(s + j);
        const uchar4 m = *reinterpret_cast<const uchar4*>(colmask + j);
        if (m.x) mx = fmaxf(mx, v.x);
        if (m.y) mx = fmaxf(mx, v.y);
        if (m.z) mx = fmaxf(mx, v.z);
        if (m.w) mx = fmaxf(mx, v.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = red[0];
        for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
        bcast = m;
    }
    __syncthreads();
    mx = bcast;
    const float k2 = scale * 1.4426950408889634f, off = mx * k2;
    float sum = 0.f;
    for (int j = threadIdx.x * 4; j < ld; j += 1024) {
        const float4 v = *reinterpret_cast<const float4*>(s + j);
        const uchar4 m = *reinterpret_cast<const uchar4*>(colmask + j);
        if (m.x) sum += exp2f(fmaf(v.x, k2, -off));
        if (m.y) sum += exp2f(fmaf(v.y, k2, -off));
        if (m.z) sum += exp2f(fmaf(v.z, k2, -off));
        if (m.w) sum += exp2f(fmaf(v.w, k2, -off));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        bcast = 1.0f / t;
    }
    __syncthreads();
    const float inv = bcast;
    for (int j = threadIdx.x * 4; j < ld; j += 1024) {
        const float4 v = *reinterpret_cast<const float4*>(s + j);
        const uchar4 m = *reinterpret_cast<const uchar4*>(colmask + j);
        const float p0 = m.x ? exp2f(fmaf(v.x, k2, -off)) * inv : 0.f, p1 = m.y ? exp2f(fmaf(v.y, k2, -off)) * inv : 0.f;
        const float p2 = m.z ? exp2f(fmaf(v.z, k2, -off)) * inv : 0.f, p3 = m.w ? exp2f(fmaf(v.w, k2, -off)) * inv : 0.f;
        uint2 o;
        o.x = pack_bf16(p0, p1);
        o.y = pack_bf16(p2, p3);
        *reinterpret_cast<uint2*>(pr + j) = o;
    }
}
__global__ void vae_colmask_kernel(uint8_t* __restrict__ m, int Hp, int Wp, int Tpad) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Tpad) return;
    const int y = j / Wp, x = j - y * Wp;
    m[j] = (j < Hp * Wp && y > 0 && y < Hp - 1 && x > 0 && x < Wp - 1) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------- conv_out
// Decoder.forward tail: conv_out(silu(conv_norm_out(x))) (3x3, C -> CO <= 4, padding 1) on the normalised padded buffer; output NCHW.
// CUDA cores (N = 3 is no tensor-core shape).  A half-warp owns 8 consecutive output pixels of a row: its 16 lanes split the
// channels (lane l: channels 8l..8l+7 of every 128, one coalesced 256-byte read per pixel), each lane keeps 8 x CO partial sums,
// one 16-byte shared-memory weight read ([tap][e][C/8] float4, conflict free) feeds 8 x CO FMAs, and the partial sums meet in a
// butterfly at the end.  Each block walks CONV_OUT_ROWS image rows so the weight staging is amortised.
constexpr int CONV_OUT_ROWS = 8;
template <int CO>
__global__ void __launch_bounds__(128) vae_conv_out_kernel(const bf16* __restrict__ xn, const bf16* __restrict__ w, const bf16* __restrict__ bias,
                                                           bf16* __restrict__ out, int Hh, int Ww, int C, int Co) {
    extern __shared__ float4 wsm[];     // [9][8][C / 8]
    const int tpp = C >> 3;
    for (int i = threadIdx.x; i < 9 * C; i += 128) {
        const int tap = i / C, c = i - tap * C;
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (int o = 0; o < Co; ++o) t[o] = __bfloat162float(w[(static_cast<size_t>(o) * 9 + tap) * C + c]);
        wsm[(tap * 8 + (c & 7)) * tpp + (c >> 3)] = make_float4(t[0], t[1], t[2], t[3]);
    }
    __syncthreads();
    const int hw = threadIdx.x >> 4, l = threadIdx.x & 15, b = blockIdx.z;
    const int xq = blockIdx.x * 64 + hw * 8;
    const bool live = xq < Ww;                     // Ww % 8 == 0
    const int x0 = live ? xq : Ww - 8;
    const int Wp = Ww + 2, Hp = Hh + 2;
    float bo[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) bo[o] = o < Co ? __bfloat162float(bias[o]) : 0.f;
    for (int yy = 0; yy < CONV_OUT_ROWS; ++yy) {
        const int y = blockIdx.y * CONV_OUT_ROWS + yy;
        if (y >= Hh) break;
        float acc[8][CO];
#pragma unroll
        for (int px = 0; px < 8; ++px)
#pragma unroll
            for (int o = 0; o < CO; ++o) acc[px][o] = 0.f;
        for (int c8 = l; c8 < tpp; c8 += 16) {
#pragma unroll 1
            for (int r = 0; r < 3; ++r) {
                const bf16* src = xn + ((static_cast<size_t>(b) * Hp + y + r) * Wp + x0) * C + c8 * 8;
                float v[10][8];
#pragma unroll
                for (int j = 0; j < 10; ++j) ld8(src + static_cast<size_t>(j) * C, v[j]);
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float4 ww = wsm[((r * 3 + dx) * 8 + e) * tpp + c8];
                        const float wq[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
                        for (int px = 0; px < 8; ++px)
#pragma unroll
                            for (int o = 0; o < CO; ++o) acc[px][o] = fmaf(v[px + dx][e], wq[o], acc[px][o]);
                    }
                }
            }
        }
#pragma unroll
        for (int px = 0; px < 8; ++px)
#pragma unroll
            for (int o = 0; o < CO; ++o) {
                float a = acc[px][o];
                a += __shfl_xor_sync(0xffffffffu, a, 8);
                a += __shfl_xor_sync(0xffffffffu, a, 4);
                a += __shfl_xor_sync(0xffffffffu, a, 2);
                a += __shfl_xor_sync(0xffffffffu, a, 1);
                acc[px][o] = a + bo[o];
            }
        if (live && l < Co) {
            float r8[8];
#pragma unroll
            for (int px = 0; px < 8; ++px) {
                float t = acc[px][0];
#pragma unroll
                for (int o = 1; o < CO; ++o) t = l == o ? acc[px][o] : t;
                r8[px] = t;
            }
            st8(out + ((static_cast<size_t>(b) * Co + l) * Hh + y) * Ww + x0, r8);
        }
    }
}

// ---------------------------------------------------------------------------------------------- engine
struct Slot {
    int kind;          // 0: conv / linear weight -> bf16 [cout][taps][cin]; 1: bias -> bf16 [n]; 2: GroupNorm vector -> f32 [n]
    bool transposed = false;   // kind 0 packed [taps][cin][cout] instead (conv_in)
    int cout, cin, taps;
    void* ptr;
    bool loaded;
};
struct Conv { const bf16* w; const bf16* b; int cin, cout, taps; };
struct Norm { const float* g; const float* b; int c; };
struct Res { Norm n1, n2; Conv c1, c2, sc; bool has_sc; };
struct Attn { Norm gn; Conv q, k, v, o; };

}  // namespace

struct nvae_engine {
    nvae_config cfg;
    int num_sms = 0, G = 32;
    int conv_share = 2;                // 3x3 convolutions with one haloed A tile per (dy, k-block): 0 never, 1 for Cout = 128, 2 always (NVAE_CONV_SHARE);
                                       // measured, 1024 x 1024 decode: 13.1 / 12.5 / 11.4 ms
    int ch[4];                         // decoder order: reversed block_out_channels
    std::map<std::string, Slot> slots;
    std::vector<void*> w_allocs, ws_allocs;
    bool finalized = false;
    // network
    Conv pq, conv_in, conv_out;
    Norm norm_out;
    Res mid_res[2];
    Attn mid_attn;
    std::vector<Res> up_res[4];
    Conv up_conv[3];
    // workspace for the current (batch, lat_h, lat_w)
    int wsB = 0, wsH = 0, wsW = 0;
    bf16 *X = nullptr, *Y = nullptr, *Nn = nullptr, *Hb = nullptr, *Sc = nullptr;
    bf16 *Q = nullptr, *Kb = nullptr, *O = nullptr, *Vt = nullptr, *Pm = nullptr;
    float* S = nullptr;
    uint8_t* colmask = nullptr;
    float4* partial = nullptr;
    float2* ab = nullptr;
    int Tpad = 0, chunk = 0;
    std::vector<GemmPlan> plans;
    size_t plan_i = 0;
    bool building = false;
    char err[512] = {0};
    int fail(int code, const char* fmt, ...) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, sizeof(err), fmt, ap);
        va_end(ap);
        return code;
    }
};

static thread_local char g_vae_err[512] = "";

#define VCK(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return h->fail(NDIT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

extern "C" const char* nvae_last_error(nvae_handle h) { return h ? h->err : g_vae_err; }

namespace {

int add_slot(nvae_engine* h, const std::string& key, int kind, int cout, int cin, int taps) {
    const size_t n = static_cast<size_t>(cout) * (kind == 0 ? static_cast<size_t>(cin) * taps : 1);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, n * (kind == 2 ? 4 : 2) + 256);
    if (e != cudaSuccess) return h->fail(NDIT_ERR_NOMEM, "cudaMalloc(%s): %s", key.c_str(), cudaGetErrorString(e));
    h->w_allocs.push_back(p);
    Slot sl;
    sl.kind = kind; sl.cout = cout; sl.cin = cin; sl.taps = taps; sl.ptr = p; sl.loaded = false;
    h->slots[key] = sl;
    return 0;
}
int add_conv(nvae_engine* h, const std::string& pre, int cin, int cout, int taps, Conv* c) {
    if (int e = add_slot(h, pre + ".weight", 0, cout, cin, taps)) return e;
    if (int e = add_slot(h, pre + ".bias", 1, cout, 0, 0)) return e;
    *c = Conv{static_cast<const bf16*>(h->slots[pre + ".weight"].ptr), static_cast<const bf16*>(h->slots[pre + ".bias"].ptr), cin, cout, taps};
    return 0;
}
int add_norm(nvae_engine* h, const std::string& pre, int c, Norm* n) {
    if (int e = add_slot(h, pre + ".weight", 2, c, 0, 0)) return e;
    if (int e = add_slot(h, pre + ".bias", 2, c, 0, 0)) return e;
    *n = Norm{static_cast<const float*>(h->slots[pre + ".weight"].ptr), static_cast<const float*>(h->slots[pre + ".bias"].ptr), c};
    return 0;
}
int add_res(nvae_engine* h, const std::string& pre, int cin, int cout, Res* r) {
    int e = 0;
    e = e ? e : add_norm(h, pre + ".norm1", cin, &r->n1);
    e = e ? e : add_conv(h, pre + ".conv1", cin, cout, 9, &r->c1);
    e = e ? e : add_norm(h, pre + ".norm2", cout, &r->n2);
    e = e ? e : add_conv(h, pre + ".conv2", cout, cout, 9, &r->c2);
    r->has_sc = cin != cout;                  // ResnetBlock2D: use_in_shortcut = in_channels != out_channels -> 1x1 conv_shortcut
    if (r->has_sc) e = e ? e : add_conv(h, pre + ".conv_shortcut", cin, cout, 1, &r->sc);
    return e;
}

template <typename T>
int ws_alloc(nvae_engine* h, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
    if (e != cudaSuccess) return h->fail(NDIT_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
    cudaMemset(q, 0, count * sizeof(T) + 256);
    h->ws_allocs.push_back(q);
    *p = static_cast<T*>(q);
    return 0;
}

// One GEMM of the walk: built (tensor maps encoded) on the first pass over a new shape, launched on the later ones.
int vae_gemm(nvae_engine* h, cudaStream_t s, const bf16* A, int a_rows, int a_cols, const bf16* W, int w_rows, int K, bf16* C, int ldc, int M,
             int N, const GemmExt& ext) {
    if (h->building) {
        GemmPlan p;
        memset(&p, 0, sizeof(p));
        p.C = C; p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.epi = EPI_STORE; p.num_sms = h->num_sms; p.pair = 1; p.use_ext = 1; p.ext = ext;
        p.bn = (N % 256 == 0 || N > 256) ? 256 : 128;
        if (N % p.bn != 0 && (p.bn != 256 || (N % 256) % 32 != 0)) return h->fail(NDIT_ERR_INVALID, "GEMM N = %d not tileable", N);
        if (K % 64 != 0) return h->fail(NDIT_ERR_INVALID, "GEMM K = %d not a multiple of 64", K);
        if (make_tmap_2d(&p.tmA, A, a_rows, a_cols, a_cols, ext.share ? 136 : 128, 64, 128)) return h->fail(NDIT_ERR_CUDA, "%s", tmap_last_error());
        if (make_tmap_2d(&p.tmB, W, w_rows, K, K, p.bn / 2, 64, 128)) return h->fail(NDIT_ERR_CUDA, "%s", tmap_last_error());
        h->plans.push_back(p);
        return 0;
    }
    if (h->plan_i >= h->plans.size()) return h->fail(NDIT_ERR_STATE, "plan list out of step");
    VCK(gemm_bf16_tn(h->plans[h->plan_i++], s));
    return 0;
}

struct Geo { int B, H, W, Hp, Wp, P, M; };      // P = padded pixels per image, M = rows of the activation matrix
Geo geo(int B, int H, int W) { return Geo{B, H, W, H + 2, W + 2, (H + 2) * (W + 2), B * (H + 2) * (W + 2)}; }

// 3x3 (taps 9) or 1x1 convolution of the padded NHWC matrix `in` [M, cin] -> `out` [M, cout]
int vae_conv(nvae_engine* h, cudaStream_t s, const Geo& g, const Conv& c, const bf16* in, bf16* out, const bf16* resid, bool mask_border) {
    GemmExt ext;
    memset(&ext, 0, sizeof(ext));
    ext.kpt = c.taps == 9 ? c.cin / 64 : 0;
    ext.share = (c.taps == 9 && (h->conv_share == 2 || (h->conv_share == 1 && c.cout == 128))) ? 1 : 0;
    ext.wp = g.Wp;
    ext.hp = mask_border ? g.Hp : 0;
    ext.bias = c.b;
    ext.resid = resid;
    ext.ldr = c.cout;
    return vae_gemm(h, s, in, g.M, c.cin, c.w, c.cout, c.taps * c.cin, out, c.cout, g.M, c.cout, ext);
}

int vae_gn(nvae_engine* h, cudaStream_t s, const Geo& g, const Norm& n, const bf16* in, bf16* out, int silu) {
    if (h->building) return 0;
    // pixels per block of the statistics pass: GN_PPC for the large stages, fewer when that would leave SMs idle (>= ~4 blocks per SM)
    int ppc = GN_PPC;
    while (ppc > 32 && (g.P + ppc - 1) / ppc < 4 * h->num_sms) ppc >>= 1;      // independent of the batch: one image decodes the same alone or in a batch
    const int chunks = (g.P + ppc - 1) / ppc;
    vae_gn_stats_kernel<<<dim3(chunks, g.B), 256, 0, s>>>(in, h->partial, g.P, n.c, chunks, ppc);
    VCK(cudaGetLastError());
    const double inv_count = 1.0 / (static_cast<double>(g.H) * g.W * (n.c / h->G));
    vae_gn_finalize_kernel<<<dim3(h->G, g.B), 256, 0, s>>>(h->partial, n.g, n.b, h->ab, n.c, chunks, h->G, inv_count, GN_EPS);
    VCK(cudaGetLastError());
    const int ppb = (256 / (n.c / 8)) * GN_VPT;          // pixels per block
    vae_gn_apply_kernel<<<dim3((g.P + ppb - 1) / ppb, g.B), 256, 0, s>>>(in, h->ab, out, g.P, n.c, g.Hp, g.Wp, silu);
    VCK(cudaGetLastError());
    return 0;
}

// ResnetBlock2D.forward (temb = None): x + conv2(silu(norm2(conv1(silu(norm1(x)))))), x through conv_shortcut when channels change
int vae_res(nvae_engine* h, cudaStream_t s, const Geo& g, const Res& r) {
    if (int e = vae_gn(h, s, g, r.n1, h->X, h->Nn, 1)) return e;
    if (int e = vae_conv(h, s, g, r.c1, h->Nn, h->Hb, nullptr, true)) return e;
    if (int e = vae_gn(h, s, g, r.n2, h->Hb, h->Nn, 1)) return e;
    const bf16* resid = h->X;
    if (r.has_sc) {
        if (int e = vae_conv(h, s, g, r.sc, h->X, h->Sc, nullptr, false)) return e;
        resid = h->Sc;
    }
    if (int e = vae_conv(h, s, g, r.c2, h->Nn, h->Y, resid, true)) return e;
    std::swap(h->X, h->Y);
    return 0;
}

// Attention.forward of the mid block (heads = 1, residual_connection = True): x + to_out(softmax(q k^T / sqrt(C)) v) on norm(x)
int vae_attn(nvae_engine* h, cudaStream_t s, const Geo& g, const Attn& a) {
    const int C = a.gn.c, T = g.P, Tpad = h->Tpad;
    if (int e = vae_gn(h, s, g, a.gn, h->X, h->Nn, 0)) return e;
    GemmExt lin;
    memset(&lin, 0, sizeof(lin));
    lin.bias = a.q.b;
    if (int e = vae_gemm(h, s, h->Nn, g.M, C, a.q.w, C, C, h->Q, C, g.M, C, lin)) return e;
    lin.bias = a.k.b;
    if (int e = vae_gemm(h, s, h->Nn, g.M, C, a.k.w, C, C, h->Kb, C, g.M, C, lin)) return e;
    const float scale = 1.0f / sqrtf(static_cast<float>(C));
    for (int b = 0; b < g.B; ++b) {
        const bf16* xn_b = h->Nn + static_cast<size_t>(b) * T * C;
        GemmExt none;
        memset(&none, 0, sizeof(none));
        // V^T[c, j] = sum_k Wv[c, k] xn[j, k]: the weight is the A operand, the tokens the W operand (rows >= T zero-filled by TMA)
        if (int e = vae_gemm(h, s, a.v.w, C, C, xn_b, T, C, h->Vt, Tpad, C, Tpad, none)) return e;
        for (int r0 = 0; r0 < T; r0 += h->chunk) {
            const int m = std::min(h->chunk, T - r0);
            const size_t row = static_cast<size_t>(b) * T + r0;
            GemmExt sc;
            memset(&sc, 0, sizeof(sc));
            sc.out_f32 = h->S;
            if (int e = vae_gemm(h, s, h->Q + row * C, m, C, h->Kb + static_cast<size_t>(b) * T * C, T, C, nullptr, Tpad, m, Tpad, sc)) return e;
            if (!h->building) {
                vae_softmax_kernel<<<m, 256, 0, s>>>(h->S, h->Pm, h->colmask, Tpad, scale);
                VCK(cudaGetLastError());
            }
            GemmExt pv;
            memset(&pv, 0, sizeof(pv));
            pv.bias = a.v.b;                 // softmax rows sum to one: P (V + 1 bv^T) = P V + bv
            if (int e = vae_gemm(h, s, h->Pm, m, Tpad, h->Vt, C, Tpad, h->O + row * C, C, m, C, pv)) return e;
        }
    }
    GemmExt out;
    memset(&out, 0, sizeof(out));
    out.bias = a.o.b;
    out.resid = h->X;
    out.ldr = C;
    out.wp = g.Wp;
    out.hp = g.Hp;
    if (int e = vae_gemm(h, s, h->O, g.M, C, a.o.w, C, C, h->Y, C, g.M, C, out)) return e;
    std::swap(h->X, h->Y);
    return 0;
}

// AutoencoderKL._decode + Decoder.forward; building = true only encodes the tensor maps in walk order
int vae_walk(nvae_engine* h, cudaStream_t s, const bf16* z, bf16* out, int B, int lh, int lw) {
    h->plan_i = 0;
    bf16* x0 = h->X;
    bf16* y0 = h->Y;
    Geo g = geo(B, lh, lw);
    if (!h->building) {
        vae_conv_in_kernel<<<g.M, std::max(32, h->ch[0] / 8), 0, s>>>(z, h->pq.w, h->pq.b, h->conv_in.w, h->conv_in.b, h->X, lh, lw,
                                                                       h->cfg.latent_channels, h->ch[0]);
        VCK(cudaGetLastError());
    }
    int e = 0;
    e = e ? e : vae_res(h, s, g, h->mid_res[0]);
    e = e ? e : vae_attn(h, s, g, h->mid_attn);
    e = e ? e : vae_res(h, s, g, h->mid_res[1]);
    for (int i = 0; i < 4 && !e; ++i) {
        for (const Res& r : h->up_res[i]) e = e ? e : vae_res(h, s, g, r);
        if (i < 3 && !e) {
            if (!h->building) {
                const long long vecs = static_cast<long long>(B) * (2 * g.H + 2) * (2 * g.W + 2) * (h->ch[i] / 8);
                vae_upsample_kernel<<<static_cast<unsigned>((vecs + 255) / 256), 256, 0, s>>>(h->X, h->Nn, B, g.H, g.W, h->ch[i]);
                VCK(cudaGetLastError());
            }
            g = geo(B, 2 * g.H, 2 * g.W);
            e = vae_conv(h, s, g, h->up_conv[i], h->Nn, h->Y, nullptr, true);
            std::swap(h->X, h->Y);
        }
    }
    if (!e) e = vae_gn(h, s, g, h->norm_out, h->X, h->Nn, 1);
    if (!e && !h->building) {
        const int C = h->ch[3];
        const dim3 grid((g.W + 63) / 64, (g.H + CONV_OUT_ROWS - 1) / CONV_OUT_ROWS, B);
        const size_t sm = 9 * static_cast<size_t>(C) * sizeof(float4);
        if (h->cfg.out_channels == 3) vae_conv_out_kernel<3><<<grid, 128, sm, s>>>(h->Nn, h->conv_out.w, h->conv_out.b, out, g.H, g.W, C, 3);
        else vae_conv_out_kernel<4><<<grid, 128, sm, s>>>(h->Nn, h->conv_out.w, h->conv_out.b, out, g.H, g.W, C, h->cfg.out_channels);
        cudaError_t ce = cudaGetLastError();
        if (ce != cudaSuccess) e = h->fail(NDIT_ERR_CUDA, "conv_out: %s", cudaGetErrorString(ce));
    }
    h->X = x0;          // the same ping-pong order on every pass (the plans hold the pointers)
    h->Y = y0;
    return e;
}

void free_ws(nvae_engine* h) {
    for (void* p : h->ws_allocs) cudaFree(p);
    h->ws_allocs.clear();
    h->plans.clear();
    h->wsB = h->wsH = h->wsW = 0;
}

int prepare(nvae_engine* h, int B, int lh, int lw, cudaStream_t s) {
    if (B == h->wsB && lh == h->wsH && lw == h->wsW) return 0;
    VCK(cudaStreamSynchronize(s));
    free_ws(h);
    // largest activation matrix of the walk: up block i at resolution 2^i with max(cin, cout) channels, its upsample output one
    // resolution up with cout channels
    size_t max_act = 0, max_p = 0;
    int H = lh, W = lw;
    for (int i = 0; i < 4; ++i) {
        const int cin = i == 0 ? h->ch[0] : h->ch[i - 1], cout = h->ch[i];
        const size_t P = static_cast<size_t>(H + 2) * (W + 2);
        max_act = std::max(max_act, static_cast<size_t>(B) * P * std::max(cin, cout));
        max_p = std::max(max_p, P);
        if (i < 3) {
            H *= 2; W *= 2;
            max_act = std::max(max_act, static_cast<size_t>(B) * (H + 2) * (W + 2) * cout);
        }
    }
    const int C0 = h->ch[0], T = (lh + 2) * (lw + 2);
    h->Tpad = (T + 63) / 64 * 64;
    // query rows per pass of the mid-block attention: the P V product has only C / 256 column tiles, so a pass needs about
    // 256 * (SM pairs) / (C / 256) rows to occupy every SM pair; the image is split evenly into passes of at most that many rows
    const int want = 256 * ((h->num_sms / 2 + C0 / 256 - 1) / (C0 / 256));
    const int passes = (T + want - 1) / want;
    long long chunk = ((T + passes - 1) / passes + 255) / 256 * 256;
    if (chunk > T) chunk = T;
    h->chunk = static_cast<int>(chunk);
    int e = 0;
    e = e ? e : ws_alloc(h, &h->X, max_act);
    e = e ? e : ws_alloc(h, &h->Y, max_act);
    e = e ? e : ws_alloc(h, &h->Nn, max_act);
    e = e ? e : ws_alloc(h, &h->Hb, max_act);
    e = e ? e : ws_alloc(h, &h->Sc, max_act);
    const size_t MT = static_cast<size_t>(B) * T * C0;
    e = e ? e : ws_alloc(h, &h->Q, MT);
    e = e ? e : ws_alloc(h, &h->Kb, MT);
    e = e ? e : ws_alloc(h, &h->O, MT);
    e = e ? e : ws_alloc(h, &h->Vt, static_cast<size_t>(C0) * h->Tpad);
    e = e ? e : ws_alloc(h, &h->S, static_cast<size_t>(h->chunk) * h->Tpad);
    e = e ? e : ws_alloc(h, &h->Pm, static_cast<size_t>(h->chunk) * h->Tpad);
    e = e ? e : ws_alloc(h, &h->colmask, static_cast<size_t>(h->Tpad));
    // most blocks any statistics pass launches per image (see vae_gn: at most GN_PPC pixels per block, halved while SMs would idle)
    const size_t chunks = std::max<size_t>((max_p + GN_PPC - 1) / GN_PPC, 8 * static_cast<size_t>(h->num_sms) + 2);
    e = e ? e : ws_alloc(h, &h->partial, static_cast<size_t>(B) * chunks * 64);
    e = e ? e : ws_alloc(h, &h->ab, static_cast<size_t>(B) * 512);
    if (e) { free_ws(h); return e; }
    vae_colmask_kernel<<<(h->Tpad + 255) / 256, 256, 0, s>>>(h->colmask, lh + 2, lw + 2, h->Tpad);
    VCK(cudaGetLastError());
    h->building = true;
    e = vae_walk(h, s, nullptr, nullptr, B, lh, lw);
    h->building = false;
    if (e) { free_ws(h); return e; }
    h->wsB = B; h->wsH = lh; h->wsW = lw;
    return 0;
}

}  // namespace

extern "C" int nvae_create(const nvae_config* c, nvae_handle* out) {
    if (!c || !out) return NDIT_ERR_INVALID;
    auto bad = [&](const char* m) { snprintf(g_vae_err, sizeof(g_vae_err), "nvae_create: %s", m); return NDIT_ERR_INVALID; };
    if (c->latent_channels < 1 || c->latent_channels > 16 || c->out_channels < 1 || c->out_channels > 4) return bad("latent_channels 1..16, out_channels 1..4");
    if (c->norm_num_groups != 32) return bad("norm_num_groups must be 32");
    if (c->layers_per_block < 1 || c->layers_per_block > 8) return bad("layers_per_block 1..8");
    for (int i = 0; i < 4; ++i)
        if (c->block_out_channels[i] != 128 && c->block_out_channels[i] != 256 && c->block_out_channels[i] != 512) return bad("block_out_channels: 128, 256 or 512");
    if (c->block_out_channels[3] < 256) return bad("the mid block needs at least 256 channels");
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return bad("no CUDA device");
    if (prop.major != 10) return bad("needs an sm_100 device (B200); there is no fallback path");
    nvae_engine* h = new nvae_engine();
    h->cfg = *c;
    h->num_sms = prop.multiProcessorCount;
    h->G = c->norm_num_groups;
    if (const char* e = getenv("NVAE_CONV_SHARE")) h->conv_share = atoi(e);
    for (int i = 0; i < 4; ++i) h->ch[i] = c->block_out_channels[3 - i];
    const int C0 = h->ch[0], Cz = c->latent_channels;
    int e = 0;
    e = e ? e : add_conv(h, "post_quant_conv", Cz, Cz, 1, &h->pq);
    e = e ? e : add_conv(h, "decoder.conv_in", Cz, C0, 9, &h->conv_in);
    if (!e) h->slots["decoder.conv_in.weight"].transposed = true;
    e = e ? e : add_res(h, "decoder.mid_block.resnets.0", C0, C0, &h->mid_res[0]);
    e = e ? e : add_res(h, "decoder.mid_block.resnets.1", C0, C0, &h->mid_res[1]);
    e = e ? e : add_norm(h, "decoder.mid_block.attentions.0.group_norm", C0, &h->mid_attn.gn);
    e = e ? e : add_conv(h, "decoder.mid_block.attentions.0.to_q", C0, C0, 1, &h->mid_attn.q);
    e = e ? e : add_conv(h, "decoder.mid_block.attentions.0.to_k", C0, C0, 1, &h->mid_attn.k);
    e = e ? e : add_conv(h, "decoder.mid_block.attentions.0.to_v", C0, C0, 1, &h->mid_attn.v);
    e = e ? e : add_conv(h, "decoder.mid_block.attentions.0.to_out.0", C0, C0, 1, &h->mid_attn.o);
    for (int i = 0; i < 4 && !e; ++i) {
        const int cin = i == 0 ? C0 : h->ch[i - 1], cout = h->ch[i];
        h->up_res[i].resize(c->layers_per_block + 1);
        for (int j = 0; j <= c->layers_per_block && !e; ++j) {
            char pre[96];
            snprintf(pre, sizeof(pre), "decoder.up_blocks.%d.resnets.%d", i, j);
            e = add_res(h, pre, j == 0 ? cin : cout, cout, &h->up_res[i][j]);
        }
        if (i < 3 && !e) {
            char pre[96];
            snprintf(pre, sizeof(pre), "decoder.up_blocks.%d.upsamplers.0.conv", i);
            e = add_conv(h, pre, cout, cout, 9, &h->up_conv[i]);
        }
    }
    e = e ? e : add_norm(h, "decoder.conv_norm_out", h->ch[3], &h->norm_out);
    e = e ? e : add_conv(h, "decoder.conv_out", h->ch[3], c->out_channels, 9, &h->conv_out);
    if (e) { snprintf(g_vae_err, sizeof(g_vae_err), "%s", h->err); nvae_destroy(h); return e; }
    const int osm = 9 * h->ch[3] * static_cast<int>(sizeof(float4));
    if (osm > 48 * 1024) {
        cudaFuncSetAttribute(vae_conv_out_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, osm);
        cudaFuncSetAttribute(vae_conv_out_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, osm);
    }
    *out = h;
    return NDIT_OK;
}

extern "C" int nvae_destroy(nvae_handle h) {
    if (!h) return NDIT_OK;
    cudaDeviceSynchronize();
    free_ws(h);
    for (void* p : h->w_allocs) cudaFree(p);
    delete h;
    return NDIT_OK;
}

extern "C" int nvae_set_weight(nvae_handle h, const char* key, const void* src, const int64_t* shape, int32_t ndim, int32_t dtype, void* stream) {
    if (!h || !key || !src || !shape) return NDIT_ERR_INVALID;
    if (dtype != NDIT_BF16 && dtype != NDIT_F32) return h->fail(NDIT_ERR_INVALID, "%s: dtype must be bf16 or f32", key);
    if (!strncmp(key, "encoder.", 8) || !strncmp(key, "quant_conv.", 11)) return NDIT_OK;      // not on the decode path
    std::string k(key);
    // diffusers < 0.18 attention names (AttentionBlock): query / key / value / proj_attn
    const char* olds[4] = {".query.", ".key.", ".value.", ".proj_attn."};
    const char* news[4] = {".to_q.", ".to_k.", ".to_v.", ".to_out.0."};
    for (int i = 0; i < 4; ++i) {
        const size_t pos = k.find(olds[i]);
        if (pos != std::string::npos && k.find("attentions.0") != std::string::npos) k.replace(pos, strlen(olds[i]), news[i]);
    }
    auto it = h->slots.find(k);
    if (it == h->slots.end()) return h->fail(NDIT_ERR_INVALID, "unexpected key %s", key);
    Slot& sl = it->second;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    h->finalized = false;
    if (sl.kind == 0) {
        // [cout, cin, kh, kw]; 1x1 weights also as [cout, cin] (Linear)
        const int side = sl.taps == 9 ? 3 : 1;
        bool ok = ndim >= 2 && shape[0] == sl.cout && shape[1] == sl.cin;
        if (ndim == 4) ok = ok && shape[2] == side && shape[3] == side;
        else ok = ok && ndim == 2 && sl.taps == 1;
        if (!ok) return h->fail(NDIT_ERR_INVALID, "%s: shape mismatch (want [%d, %d, %d, %d])", key, sl.cout, sl.cin, side, side);
        const size_t total = static_cast<size_t>(sl.cout) * sl.cin * sl.taps;
        const int grid = static_cast<int>(std::min<size_t>((total + 255) / 256, 4096));
        if (sl.transposed) vae_pack_conv_t_kernel<<<grid, 256, 0, s>>>(static_cast<bf16*>(sl.ptr), src, dtype == NDIT_F32, sl.cout, sl.cin, sl.taps);
        else vae_pack_conv_kernel<<<grid, 256, 0, s>>>(static_cast<bf16*>(sl.ptr), src, dtype == NDIT_F32, sl.cout, sl.cin, sl.taps);
    } else {
        if (ndim != 1 || shape[0] != sl.cout) return h->fail(NDIT_ERR_INVALID, "%s: shape mismatch (want [%d])", key, sl.cout);
        vae_copy_vec_kernel<<<(sl.cout + 255) / 256, 256, 0, s>>>(sl.ptr, sl.kind == 2, src, dtype == NDIT_F32, sl.cout);
    }
    VCK(cudaGetLastError());
    sl.loaded = true;
    return NDIT_OK;
}

extern "C" int nvae_finalize_weights(nvae_handle h, void* stream) {
    if (!h) return NDIT_ERR_INVALID;
    for (auto& kv : h->slots)
        if (!kv.second.loaded) return h->fail(NDIT_ERR_STATE, "missing key %s", kv.first.c_str());
    VCK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    h->finalized = true;
    return NDIT_OK;
}

extern "C" int nvae_decode(nvae_handle h, const void* z, int32_t batch, int32_t lat_h, int32_t lat_w, void* out, void* stream) {
    if (!h || !z || !out) return NDIT_ERR_INVALID;
    if (!h->finalized) return h->fail(NDIT_ERR_STATE, "weights not finalized");
    if (batch < 1 || batch > 64 || lat_h < 2 || lat_w < 2 || lat_h > 512 || lat_w > 512) return h->fail(NDIT_ERR_INVALID, "batch 1..64, latent side 2..512");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (int e = prepare(h, batch, lat_h, lat_w, s)) return e;
    return vae_walk(h, s, static_cast<const bf16*>(z), static_cast<bf16*>(out), batch, lat_h, lat_w);
}
