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

#include "kernels.h"
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
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
resid_rms_mod_kernel(bf16* __restrict__ X, const bf16* __restrict__ o, const bf16* __restrict__ w_post,
                     const float* __restrict__ tanh_g, const bf16* __restrict__ w_pre,
                     const float* __restrict__ onepls, bf16* __restrict__ u, int M, int rows_per_batch, int D,
                     int mod_stride, float eps) {
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
        const float rinv = rsqrtf(warp_sum(ss) / D + eps);
        const float* tg = tanh_g + static_cast<size_t>(b) * mod_stride;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + i * 32;
            if (v < nvec) {
                float w[8];
                load8(w_post + v * 8, w);
                const float4 g0 = *reinterpret_cast<const float4*>(tg + v * 8);
                const float4 g1 = *reinterpret_cast<const float4*>(tg + v * 8 + 4);
                const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float n = bf16_round(bf16_round(ov[i][e] * rinv) * w[e]);
                    x[i][e] = bf16_round(x[i][e] + bf16_round(g[e] * n));
                }
                store8(X + off + v * 8, x[i]);
            }
        }
    }
    if (u == nullptr) return;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
#pragma unroll
            for (int e = 0; e < 8; ++e) ss += x[i][e] * x[i][e];
        }
    }
    const float rinv = rsqrtf(warp_sum(ss) / D + eps);
    const float* op = onepls + static_cast<size_t>(b) * mod_stride;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            float w[8], r[8];
            load8(w_pre + v * 8, w);
            const float4 s0 = *reinterpret_cast<const float4*>(op + v * 8);
            const float4 s1 = *reinterpret_cast<const float4*>(op + v * 8 + 4);
            const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) r[e] = bf16_round(bf16_round(x[i][e] * rinv) * w[e]) * sc[e];
            store8(u + off + v * 8, r);
        }
    }
}

cudaError_t resid_rms_mod(bf16* X, const bf16* o, const bf16* w_post, const float* tanh_g, const bf16* w_pre,
                          const float* onepls, bf16* u, int M, int rows_per_batch, int D, int mod_stride, float eps,
                          cudaStream_t s) {
    if (D % 8 != 0 || D > MAX_VEC * 256 || mod_stride % 4 != 0) return cudaErrorInvalidValue;
    const int nv = (D / 8 + 31) / 32;
    const dim3 grid((M + ROW_WARPS - 1) / ROW_WARPS), block(ROW_WARPS * 32);
#define LAUNCH(NVV)                                                                                             \
    resid_rms_mod_kernel<NVV><<<grid, block, 0, s>>>(X, o, w_post, tanh_g, w_pre, onepls, u, M, rows_per_batch, \
                                                     D, mod_stride, eps)
    if (nv <= 3) LAUNCH(3);
    else if (nv <= 9) LAUNCH(9);
    else LAUNCH(MAX_VEC);
#undef LAUNCH
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Final layer: last gated residual + LayerNorm(no affine) * (1+scale) -> bf16 -> Linear(D->O)+bias.
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
final_layer_kernel(const bf16* __restrict__ X, const bf16* __restrict__ o, const bf16* __restrict__ w_post,
                   const float* __restrict__ tanh_g, const float* __restrict__ onepls, const bf16* __restrict__ Wout,
                   const bf16* __restrict__ bout, float* __restrict__ out, int M, int rows_per_batch, int D, int O,
                   int mod_stride, float eps_rms) {
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
        const float* tg = tanh_g + static_cast<size_t>(b) * mod_stride;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + i * 32;
            if (v < nvec) {
                float w[8];
                load8(w_post + v * 8, w);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float n = bf16_round(bf16_round(ov[i][e] * rinv) * w[e]);
                    x[i][e] = bf16_round(x[i][e] + bf16_round(tg[v * 8 + e] * n));
                }
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
    const float* op = onepls + static_cast<size_t>(b) * mod_stride;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
#pragma unroll
            for (int e = 0; e < 8; ++e) x[i][e] = bf16_round((x[i][e] - mean) * rstd * op[v * 8 + e]);
        }
    }
    for (int oc = 0; oc < O; ++oc) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = lane + i * 32;
            if (v < nvec) {
                float w[8];
                load8(Wout + static_cast<size_t>(oc) * D + v * 8, w);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc += x[i][e] * w[e];
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) out[static_cast<size_t>(row) * O + oc] = bf16_round(acc + __bfloat162float(bout[oc]));
    }
}

cudaError_t final_layer(const bf16* X, const bf16* o, const bf16* w_post, const float* tanh_g, const float* onepls,
                        const bf16* Wout, const bf16* bout, float* out, int M, int rows_per_batch, int D, int O,
                        int mod_stride, float eps, cudaStream_t s) {
    if (D % 8 != 0 || D > MAX_VEC * 256) return cudaErrorInvalidValue;
    const int nv = (D / 8 + 31) / 32;
    const dim3 grid((M + ROW_WARPS - 1) / ROW_WARPS), block(ROW_WARPS * 32);
#define LAUNCH(NVV)                                                                                              \
    final_layer_kernel<NVV><<<grid, block, 0, s>>>(X, o, w_post, tanh_g, onepls, Wout, bout, out, M,             \
                                                   rows_per_batch, D, O, mod_stride, eps)
    if (nv <= 3) LAUNCH(3);
    else if (nv <= 9) LAUNCH(9);
    else LAUNCH(MAX_VEC);
#undef LAUNCH
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Patchify + x_embedder.  One block per token, thread d-strided over D.
__global__ void patch_embed_kernel(const bf16* __restrict__ x, const bf16* __restrict__ Wx, const bf16* __restrict__ bx,
                                   bf16* __restrict__ X, int n_unique, int C, int Hh, int Ww, int D) {
    const int Wp = Ww >> 1, Hp = Hh >> 1;
    const int tok = blockIdx.x;               // over B * Hp * Wp
    const int N = Hp * Wp;
    const int b = tok / N, t = tok % N;
    const int bi = b % n_unique;              // second half of the batch re-uses the first (model.py:901-902)
    const int i = t / Wp, j = t % Wp;
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

cudaError_t patch_embed(const bf16* x, const bf16* Wx, const bf16* bx, bf16* X, int B, int n_unique, int C, int Hh,
                        int Ww, int D, cudaStream_t s) {
    if (C * 4 > 64 || (C * 4) % 8 != 0 || (Hh & 1) || (Ww & 1)) return cudaErrorInvalidValue;
    patch_embed_kernel<<<B * (Hh / 2) * (Ww / 2), 256, 0, s>>>(x, Wx, bx, X, n_unique, C, Hh, Ww, D);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Conditioning inputs: sinusoidal timestep features and LayerNorm'd masked-mean caption.
__global__ void cond_prepare_kernel(float t, const bf16* __restrict__ cap, const uint8_t* __restrict__ mask,
                                    const bf16* __restrict__ ln_w, const bf16* __restrict__ ln_b,
                                    float* __restrict__ tf, float* __restrict__ pool, int T, int C, int do_caption) {
    const int b = blockIdx.x;
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

cudaError_t cond_prepare(float t, const bf16* cap, const uint8_t* mask, const bf16* ln_w, const bf16* ln_b, float* tf,
                         float* pool, int B, int T, int C, int do_caption, cudaStream_t s) {
    cond_prepare_kernel<<<B, 256, do_caption ? C * sizeof(float) : 0, s>>>(t, cap, mask, ln_w, ln_b, tf, pool, T, C,
                                                                         do_caption);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Small-batch Linear: one warp per output feature, B <= 4 input rows share each weight row.
constexpr int GEMV_MAXB = 4;
__global__ void __launch_bounds__(256)
gemv_rows_kernel(const float* __restrict__ in, const bf16* __restrict__ W, const bf16* __restrict__ bias,
                 const float* __restrict__ addend, float* __restrict__ out, int B, int O, int K, int in_silu, int post,
                 int adaln_D, int adaln_blocks) {
    const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (o >= O) return;
    const int lane = threadIdx.x & 31;
    float acc[GEMV_MAXB] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane * 8; k < K; k += 256) {
        float w[8];
        load8(W + static_cast<size_t>(o) * K + k, w);
#pragma unroll
        for (int b = 0; b < GEMV_MAXB; ++b) {
            if (b < B) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float xv = in[static_cast<size_t>(b) * K + k + e];
                    if (in_silu) xv = bf16_round(silu_f(xv));
                    acc[b] += xv * w[e];
                }
            }
        }
    }
#pragma unroll
    for (int b = 0; b < GEMV_MAXB; ++b) acc[b] = warp_sum(acc[b]);
    if (lane == 0) {
        const float bv = bias ? __bfloat162float(bias[o]) : 0.f;
        for (int b = 0; b < B; ++b) {
            float y = bf16_round(acc[b] + bv);
            if (addend) y = bf16_round(y + addend[static_cast<size_t>(b) * O + o]);
            if (post == POST_SILU) y = bf16_round(silu_f(y));
            else if (post == POST_ADALN) {
                // per layer chunks [scale_msa | gate_msa | scale_mlp | gate_mlp] (model.py:595), then the
                // final layer's scale: scale -> bf16(1+scale); gate -> bf16(tanh(gate))
                const int chunk = o / adaln_D;
                const bool is_gate = (chunk < adaln_blocks * 4) && (chunk & 1);
                y = is_gate ? bf16_round(tanhf(y)) : bf16_round(1.0f + y);
            }
            out[static_cast<size_t>(b) * O + o] = y;
        }
    }
}

cudaError_t gemv_rows(const float* in, const bf16* W, const bf16* bias, const float* addend, float* out, int B, int O,
                      int K, int in_silu, int post, int adaln_D, int adaln_blocks, cudaStream_t s) {
    if (B > GEMV_MAXB || K % 8 != 0) return cudaErrorInvalidValue;
    gemv_rows_kernel<<<(O + 7) / 8, 256, 0, s>>>(in, W, bias, addend, out, B, O, K, in_silu, post, adaln_D,
                                                 adaln_blocks);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// RoPE table: tab[token][m] = (cos, sin)(pos * w_{m/2}), m even -> row index, m odd -> column index
// (model.py:951-961).  w_i = (theta)^(-4i/hd) / linear_factor computed like the reference in fp32.
__global__ void rope_table_kernel(float2* __restrict__ tab, int Hp, int Wp, int hd, float theta, float linear_factor) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = hd >> 1;
    if (idx >= Hp * Wp * half) return;
    const int m = idx % half, tok = idx / half;
    const int i = tok / Wp, j = tok % Wp;
    const int fi = m >> 1;
    const float freq = 1.0f / powf(theta, static_cast<float>(4 * fi) / static_cast<float>(hd)) / linear_factor;
    const float pos = static_cast<float>((m & 1) ? j : i);
    const float a = pos * freq;
    float sn, cs;
    sincosf(a, &sn, &cs);
    tab[idx] = make_float2(cs, sn);
}

cudaError_t rope_table(float2* tab, int Hp, int Wp, int hd, float theta, float linear_factor, cudaStream_t s) {
    const int n = Hp * Wp * (hd / 2);
    rope_table_kernel<<<(n + 255) / 256, 256, 0, s>>>(tab, Hp, Wp, hd, theta, linear_factor);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// q/k LayerNorm over ALL heads jointly + 2-D RoPE, in place on the fused qkv GEMM output.
// One warp per token row.  seg 0 = q (width H*hd), seg 1 = k (width Hkv*hd).
template <int NV>
__device__ __forceinline__ void ln_rope_segment(bf16* __restrict__ p, int width, const bf16* __restrict__ w,
                                                const bf16* __restrict__ bsh, const float2* __restrict__ rp, int hd,
                                                int lane) {
    const int nvec = width >> 3;
    float x[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = lane + i * 32;
        if (v < nvec) {
            load8(p + v * 8, x[i]);
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
    const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    bf16* p = qkv + static_cast<size_t>(row) * ld;
    const float2* rp = rope + static_cast<size_t>(row % N_tokens) * (hd >> 1);
    ln_rope_segment<NVQ>(p, H * hd, qw, qb, rp, hd, lane);
    ln_rope_segment<NVK>(p + H * hd, Hkv * hd, kw, kb, rp, hd, lane);
}

cudaError_t ln_rope_qk(bf16* qkv, int ld, const bf16* qw, const bf16* qb, const bf16* kw, const bf16* kb,
                       const float2* rope, int M, int N_tokens, int H, int Hkv, int hd, cudaStream_t s) {
    if (hd % 8 != 0 || H * hd > 9 * 256 || Hkv * hd > 9 * 256 || ld % 8 != 0) return cudaErrorInvalidValue;
    const dim3 grid((M + ROW_WARPS - 1) / ROW_WARPS), block(ROW_WARPS * 32);
    if (Hkv * hd <= 3 * 256)
        ln_rope_qk_kernel<9, 3><<<grid, block, 0, s>>>(qkv, ld, qw, qb, kw, kb, rope, M, N_tokens, H, Hkv, hd);
    else
        ln_rope_qk_kernel<9, 9><<<grid, block, 0, s>>>(qkv, ld, qw, qb, kw, kb, rope, M, N_tokens, H, Hkv, hd);
    return cudaGetLastError();
}

// LayerNorm(width)+affine in place on rows (ky_norm, model.py:421), batched over layers.
__global__ void __launch_bounds__(ROW_WARPS * 32)
ln_rows_kernel(bf16* __restrict__ x, int ld, size_t lsx, const bf16* __restrict__ w, const bf16* __restrict__ b,
               size_t lsw, int M, int width) {
    const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const int l = blockIdx.y;
    ln_rope_segment<9>(x + l * lsx + static_cast<size_t>(row) * ld, width, w + l * lsw, b + l * lsw, nullptr, 8,
                       threadIdx.x & 31);
}

cudaError_t ln_rows(bf16* x, int ld, size_t layer_stride_x, const bf16* w, const bf16* b, size_t layer_stride_w, int M,
                    int width, int layers, cudaStream_t s) {
    if (width % 8 != 0 || width > 9 * 256) return cudaErrorInvalidValue;
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
__global__ void transpose_v_kernel(const bf16* __restrict__ src, int ld, int col0, size_t sls, bf16* __restrict__ dst,
                                   int ld_dst, size_t dls, int N, int G, int hd) {
    __shared__ bf16 tile[32][33];
    const int l = blockIdx.z;
    const int bg = blockIdx.y;                 // b * G + g
    const int b = bg / G, g = bg % G;
    const int tiles_d = (hd + 31) / 32;
    const int n0 = (blockIdx.x / tiles_d) * 32, d0 = (blockIdx.x % tiles_d) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r, d = d0 + tx;
        bf16 v = __float2bfloat16(0.f);
        if (n < N && d < hd) v = src[l * sls + (static_cast<size_t>(b) * N + n) * ld + col0 + g * hd + d];
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int d = d0 + r, n = n0 + tx;
        if (d < hd && n < N) dst[l * dls + (static_cast<size_t>(bg) * hd + d) * ld_dst + n] = tile[tx][r];
    }
}

cudaError_t transpose_v(const bf16* src, int ld, int col0, size_t src_layer_stride, bf16* dst, int ld_dst,
                        size_t dst_layer_stride, int B, int N, int G, int hd, int layers, cudaStream_t s) {
    const dim3 grid(((N + 31) / 32) * ((hd + 31) / 32), B * G, layers);
    transpose_v_kernel<<<grid, 256, 0, s>>>(src, ld, col0, src_layer_stride, dst, ld_dst, dst_layer_stride, N, G, hd);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// unpatchify (token feature order (ph,pw,c_out), model.py:753-754) + keep first C of 2C channels (:859-861)
// + CFG on channels 0..2 only (:904-913).
__global__ void unpatchify_cfg_kernel(const float* __restrict__ tok, bf16* __restrict__ v_out, int n, int C, int Hh,
                                      int Ww, int O, float cfg_scale) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over n * C * Hh * Ww
    const int total = n * C * Hh * Ww;
    if (idx >= total) return;
    const int x = idx % Ww, y = (idx / Ww) % Hh, c = (idx / (Ww * Hh)) % C, s = idx / (Ww * Hh * C);
    const int Wp = Ww >> 1, N = (Hh >> 1) * Wp;
    const int t = (y >> 1) * Wp + (x >> 1);
    const int Cout = O / 4;
    const int f = ((y & 1) * 2 + (x & 1)) * Cout + c;
    const float cond = tok[(static_cast<size_t>(s) * N + t) * O + f];
    const float unc = tok[(static_cast<size_t>(s + n) * N + t) * O + f];
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

cudaError_t unpatchify_cfg(const float* tok, bf16* v_out, int n, int C, int Hh, int Ww, int O, float cfg_scale,
                           cudaStream_t s) {
    const int total = n * C * Hh * Ww;
    unpatchify_cfg_kernel<<<(total + 255) / 256, 256, 0, s>>>(tok, v_out, n, C, Hh, Ww, O, cfg_scale);
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

cudaError_t axpy_bf16(bf16* y_out, const bf16* y_in, const bf16* v, float dt, size_t count, cudaStream_t s) {
    axpy_bf16_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, s>>>(y_out, y_in, v, dt, count);
    return cudaGetLastError();
}

}  // namespace ndit
