// Caption-encoder end of the Lumina-Next-T2I sampling path (include/ndit_text.h): the Gemma decoder stack behind
//   text_encoder(input_ids, attention_mask, output_hidden_states=True).hidden_states[-2]     (lumina_next_t2i/sample.py:46-50)
// restated from transformers 5.5.0 models/gemma/modeling_gemma.py (cited per function below).  The projections run on the same
// tcgen05 GEMM kernels as the denoiser (fused q|k|v, o, gate|up with a GeGLU epilogue, down); the work is weight-streaming
// bound (3.7 GB of bf16 weights for 17 layers of gemma-2b against at most 512 rows of activations), so everything else is
// small row-wise CUDA: scaled embedding gather, RMSNorm with (1 + w), rotate-half RoPE with bf16 cos / sin, and a causal,
// padding-masked attention over at most 1024 tokens at head_dim 256 (one warp per query row).
// Rounding points follow the bf16 module: every Linear output, every tensor op of apply_rotary_pos_emb, the softmax
// probabilities, both residual additions and the RMSNorm output are bf16; reductions and the softmax are fp32.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <set>
#include <string>
#include <vector>

#include "../../include/ndit.h"
#include "../../include/ndit_text.h"
#include "kernels.h"
#include "ptx.cuh"

using namespace ndit;

namespace {

constexpr int TX_WARPS = 8;

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
inline float host_bf16_round(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    const uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    u &= 0xffff0000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}

// hidden = embed_tokens(ids) * bf16(sqrt(hidden_size))            (GemmaTextScaledWordEmbedding.forward, modeling_gemma.py:60-61)
__global__ void tx_embed_kernel(const long long* __restrict__ ids, const bf16* __restrict__ emb, bf16* __restrict__ x, int M, int D,
                                int V, float scale_b) {
    const int row = blockIdx.x;
    if (row >= M) return;
    long long id = ids[row];
    if (id < 0 || id >= V) id = 0;
    const bf16* e = emb + static_cast<size_t>(id) * D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) x[static_cast<size_t>(row) * D + d] = __float2bfloat16_rn(__bfloat162float(e[d]) * scale_b);
}

// x <- bf16(x + o) when o != nullptr (decoder-layer residual, :333,:339); u = bf16((x * rsqrt(mean(x^2) + eps)) * (1 + w))   (GemmaRMSNorm, :70-78)
__global__ void __launch_bounds__(TX_WARPS * 32)
tx_resid_rmsnorm_kernel(bf16* __restrict__ x, const bf16* __restrict__ o, const bf16* __restrict__ w, bf16* __restrict__ u, int M, int D,
                        float eps) {
    const int row = blockIdx.x * TX_WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31;
    bf16* xr = x + static_cast<size_t>(row) * D;
    float ss = 0.f;
    for (int d = lane * 8; d < D; d += 256) {
        float v[8];
        load8(xr + d, v);
        if (o != nullptr) {
            float ov[8];
            load8(o + static_cast<size_t>(row) * D + d, ov);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = bf16_round(v[e] + ov[e]);
            store8(xr + d, v);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += v[e] * v[e];
    }
    ss = warp_sum(ss);
    if (u == nullptr) return;
    const float r = rsqrtf(ss / D + eps);
    __syncwarp();
    for (int d = lane * 8; d < D; d += 256) {
        float v[8], g[8], out[8];
        load8(xr + d, v);
        load8(w + d, g);
#pragma unroll
        for (int e = 0; e < 8; ++e) out[e] = (v[e] * r) * (1.0f + g[e]);
        store8(u + static_cast<size_t>(row) * D + d, out);
    }
}

// cos / sin of GemmaRotaryEmbedding.forward (:151-163): fp32 angles pos * inv_freq, inv_freq = 1 / theta^(2i / head_dim), cast to bf16
__global__ void tx_rope_table_kernel(float2* __restrict__ tab, int T, int half, float theta) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * half) return;
    const int pos = idx / half, i = idx % half;
    const float inv = 1.0f / powf(theta, static_cast<float>(2 * i) / static_cast<float>(2 * half));
    const float a = static_cast<float>(pos) * inv;
    tab[idx] = make_float2(bf16_round(cosf(a)), bf16_round(sinf(a)));
}

// apply_rotary_pos_emb (:165-195) on the q heads and k heads of the fused projection, in place:
//   q_embed = (q * cos) + (rotate_half(q) * sin), rotate_half(x) = cat(-x2, x1); every product and the sum are bf16 tensor ops
__global__ void tx_rope_kernel(bf16* __restrict__ qkv, int ld, const float2* __restrict__ tab, int M, int T, int n_heads, int hd) {
    const int half = hd >> 1;
    const size_t total = static_cast<size_t>(M) * n_heads * half;
    const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int i = static_cast<int>(idx % half);
    const int head = static_cast<int>((idx / half) % n_heads);
    const int row = static_cast<int>(idx / (static_cast<size_t>(half) * n_heads));
    const float2 cs = tab[static_cast<size_t>(row % T) * half + i];
    bf16* p = qkv + static_cast<size_t>(row) * ld + head * hd;
    const float a = __bfloat162float(p[i]), b = __bfloat162float(p[i + half]);
    p[i] = __float2bfloat16_rn(bf16_round(a * cs.x) + bf16_round(-b * cs.y));
    p[i + half] = __float2bfloat16_rn(bf16_round(b * cs.x) + bf16_round(a * cs.y));
}

// eager_attention_forward (:210-232) with the causal + padding mask of create_causal_mask: one warp per query row.
//   scores = (q . k) * head_dim^-0.5 (fp32), masked keys excluded, softmax in fp32, probabilities to bf16, out = bf16(P V)
// qkv rows: [q heads | k heads | v heads]; kv head of query head h is h / (H / Hkv).  Scores live in shared memory (T <= 1024).
template <int HD>
__global__ void __launch_bounds__(TX_WARPS * 32)
tx_attention_kernel(const bf16* __restrict__ qkv, int ld, const long long* __restrict__ mask, bf16* __restrict__ out, int T, int H, int Hkv,
                    float scale) {
    extern __shared__ float sh[];                 // [TX_WARPS][HD] q rows, then [TX_WARPS][T] scores
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * TX_WARPS + warp;   // query position
    const int h = blockIdx.y, b = blockIdx.z;
    if (i >= T) return;
    const int g = h / (H / Hkv);
    float* qs = sh + warp * HD;
    float* sc = sh + TX_WARPS * HD + warp * T;
    const bf16* qp = qkv + (static_cast<size_t>(b) * T + i) * ld + h * HD;
    for (int d = lane; d < HD; d += 32) qs[d] = __bfloat162float(qp[d]);
    __syncwarp();
    const bf16* kbase = qkv + static_cast<size_t>(b) * T * ld + (H + g) * HD;
    const bf16* vbase = qkv + static_cast<size_t>(b) * T * ld + (H + Hkv + g) * HD;
    float mx = -INFINITY;
    for (int j0 = 0; j0 <= i; j0 += 32) {
        const int j = j0 + lane;
        float s = -INFINITY;
        if (j <= i && (mask == nullptr || mask[static_cast<size_t>(b) * T + j] != 0)) {
            const bf16* kp = kbase + static_cast<size_t>(j) * ld;
            float acc = 0.f;
#pragma unroll 4
            for (int d = 0; d < HD; d += 8) {
                float kv[8];
                load8(kp + d, kv);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc = fmaf(qs[d + e], kv[e], acc);
            }
            s = bf16_round(acc) * scale;          // matmul output is a bf16 tensor, the scaling a bf16 tensor op
            s = bf16_round(s);
        }
        if (j < T) sc[j] = s;
        mx = fmaxf(mx, s);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncwarp();
    float sum = 0.f;
    for (int j = lane; j <= i; j += 32) {
        const float p = (mx == -INFINITY) ? 0.f : __expf(sc[j] - mx);
        sc[j] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.0f / sum : 0.f;
    __syncwarp();
    float acc[HD / 32];
#pragma unroll
    for (int e = 0; e < HD / 32; ++e) acc[e] = 0.f;
    for (int j = 0; j <= i; ++j) {
        const float p = bf16_round(sc[j] * inv);  // softmax(dtype=float32).to(query.dtype)
        if (p == 0.f) continue;
        const bf16* vp = vbase + static_cast<size_t>(j) * ld;
#pragma unroll
        for (int e = 0; e < HD / 32; ++e) acc[e] = fmaf(p, __bfloat162float(vp[lane + 32 * e]), acc[e]);
    }
    bf16* op = out + (static_cast<size_t>(b) * T + i) * (static_cast<size_t>(H) * HD) + h * HD;
#pragma unroll
    for (int e = 0; e < HD / 32; ++e) op[lane + 32 * e] = __float2bfloat16_rn(acc[e]);
}

// dst[row_map(r)][c] = src[r][c] (bf16 or f32 source); blk != 0: rows go to (r / blk) * blk_stride + r % blk + row0 (gate | up interleave)
__global__ void tx_place_kernel(bf16* __restrict__ dst, size_t dst_ld, const void* __restrict__ src, int src_f32, size_t rows, size_t cols,
                                size_t blk, size_t blk_stride, size_t row0) {
    const size_t total = rows * cols;
    for (size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t r = idx / cols, c = idx % cols;
        const size_t dr = blk ? (r / blk) * blk_stride + (r % blk) + row0 : r + row0;
        const float v = src_f32 ? static_cast<const float*>(src)[idx] : __bfloat162float(static_cast<const bf16*>(src)[idx]);
        dst[dr * dst_ld + c] = __float2bfloat16_rn(v);
    }
}

}  // namespace

struct ntxt_engine {
    ntxt_config cfg;
    int D, L, Lrun, H, Hkv, hd, F, Wq, Mmax, num_sms;
    bf16 *emb = nullptr, *Wqkv = nullptr, *Wo = nullptr, *W13 = nullptr, *W2 = nullptr, *ln1 = nullptr, *ln2 = nullptr;
    bf16 *x = nullptr, *u = nullptr, *qkv = nullptr, *attn = nullptr, *o = nullptr, *hbuf = nullptr;
    float2* rope = nullptr;
    int rope_T = 0;
    std::vector<void*> allocs;
    std::set<std::string> seen;
    bool finalized = false;
    int plan_M = 0;
    std::vector<GemmPlan> p_qkv, p_o, p_w13, p_w2;
    char err[512] = {0};
    int fail(int code, const char* fmt, ...) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, sizeof(err), fmt, ap);
        va_end(ap);
        return code;
    }
};

static thread_local char g_txt_err[512] = "";

#define TCK(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) return h->fail(NDIT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

template <typename T>
static int tx_alloc(ntxt_engine* h, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
    if (e != cudaSuccess) return h->fail(NDIT_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
    cudaMemset(q, 0, count * sizeof(T) + 256);
    h->allocs.push_back(q);
    *p = static_cast<T*>(q);
    return 0;
}

extern "C" const char* ntxt_last_error(ntxt_handle h) { return h ? h->err : g_txt_err; }

extern "C" int ntxt_create(const ntxt_config* c, ntxt_handle* out) {
    if (!c || !out) return NDIT_ERR_INVALID;
    auto bad = [&](const char* m) { snprintf(g_txt_err, sizeof(g_txt_err), "ntxt_create: %s", m); return NDIT_ERR_INVALID; };
    if (c->head_dim != 256) return bad("head_dim must be 256 (Gemma)");
    if (c->num_hidden_layers < 2 || c->num_attention_heads < 1 || c->num_key_value_heads < 1 || c->num_attention_heads % c->num_key_value_heads)
        return bad("bad layer / head counts");
    if (c->hidden_size % 8 || c->intermediate_size % 128 || c->vocab_size < 1 || c->max_tokens < 1) return bad("hidden_size % 8, intermediate_size % 128");
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return bad("no CUDA device");
    if (prop.major != 10) return bad("needs an sm_100 device (B200); there is no fallback path");
    ntxt_engine* h = new ntxt_engine();
    h->cfg = *c;
    h->D = c->hidden_size; h->L = c->num_hidden_layers; h->Lrun = h->L - 1; h->H = c->num_attention_heads; h->Hkv = c->num_key_value_heads;
    h->hd = c->head_dim; h->F = c->intermediate_size; h->Wq = (h->H + 2 * h->Hkv) * h->hd; h->Mmax = c->max_tokens; h->num_sms = prop.multiProcessorCount;
    const size_t D = h->D, F = h->F, Wq = h->Wq, R = h->Lrun, M = h->Mmax, HD = (size_t)h->H * h->hd;
    int e = 0;
    e |= tx_alloc(h, &h->emb, (size_t)c->vocab_size * D);
    e |= tx_alloc(h, &h->Wqkv, R * Wq * D); e |= tx_alloc(h, &h->Wo, R * D * HD); e |= tx_alloc(h, &h->W13, R * 2 * F * D); e |= tx_alloc(h, &h->W2, R * D * F);
    e |= tx_alloc(h, &h->ln1, R * D); e |= tx_alloc(h, &h->ln2, R * D);
    e |= tx_alloc(h, &h->x, M * D); e |= tx_alloc(h, &h->u, M * D); e |= tx_alloc(h, &h->qkv, M * Wq); e |= tx_alloc(h, &h->attn, M * HD);
    e |= tx_alloc(h, &h->o, M * D); e |= tx_alloc(h, &h->hbuf, M * F);
    e |= tx_alloc(h, &h->rope, (size_t)1024 * (h->hd / 2));
    if (e) { snprintf(g_txt_err, sizeof(g_txt_err), "%s", h->err); ntxt_destroy(h); return e; }
    *out = h;
    return NDIT_OK;
}

extern "C" int ntxt_destroy(ntxt_handle h) {
    if (!h) return NDIT_OK;
    cudaDeviceSynchronize();
    for (void* p : h->allocs) cudaFree(p);
    delete h;
    return NDIT_OK;
}

static int tx_place(ntxt_engine* h, bf16* dst, size_t dst_ld, const void* src, int dtype, size_t rows, size_t cols, size_t blk, size_t blk_stride,
                    size_t row0, cudaStream_t s) {
    const size_t total = rows * cols;
    const int grid = (int)((total + 255) / 256 > 8192 ? 8192 : (total + 255) / 256);
    tx_place_kernel<<<grid, 256, 0, s>>>(dst, dst_ld, src, dtype == NDIT_F32, rows, cols, blk, blk_stride, row0);
    TCK(cudaGetLastError());
    return 0;
}

extern "C" int ntxt_set_weight(ntxt_handle h, const char* key, const void* src, const int64_t* shape, int32_t ndim, int32_t dtype, void* stream) {
    if (!h || !key || !src || !shape) return NDIT_ERR_INVALID;
    if (dtype != NDIT_BF16 && dtype != NDIT_F32) return h->fail(NDIT_ERR_INVALID, "%s: dtype must be bf16 or f32", key);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t D = h->D, F = h->F, Wq = h->Wq, hd = h->hd, HD = (size_t)h->H * hd, KV = (size_t)h->Hkv * hd;
    const size_t r = ndim >= 1 ? (size_t)shape[0] : 1, c = ndim >= 2 ? (size_t)shape[1] : 1;
    auto want = [&](size_t er, size_t ec) -> int {
        const bool ok = (ec == 0) ? (ndim == 1 && r == er) : (ndim == 2 && r == er && c == ec);
        if (!ok) return h->fail(NDIT_ERR_INVALID, "%s: shape mismatch (got [%zu,%zu] ndim %d, want [%zu,%zu])", key, r, c, ndim, er, ec);
        return 0;
    };
    h->finalized = false;
    if (!strcmp(key, "embed_tokens.weight")) {
        if (int e = want((size_t)h->cfg.vocab_size, D)) return e;
        h->seen.insert(key);
        return tx_place(h, h->emb, D, src, dtype, r, c, 0, 0, 0, s);
    }
    if (!strcmp(key, "norm.weight")) { if (int e = want(D, 0)) return e; return NDIT_OK; }   // final norm: not on the path to hidden_states[-2]
    int l = -1, n = 0;
    if (sscanf(key, "layers.%d.%n", &l, &n) != 1 || l < 0 || l >= h->L) return h->fail(NDIT_ERR_INVALID, "unexpected key %s", key);
    const char* sub = key + n;
    const bool skip = l >= h->Lrun;            // the last layer is not evaluated
    struct { const char* name; size_t er, ec; int kind; } tab[] = {
        {"self_attn.q_proj.weight", HD, D, 0}, {"self_attn.k_proj.weight", KV, D, 1}, {"self_attn.v_proj.weight", KV, D, 2},
        {"self_attn.o_proj.weight", D, HD, 3}, {"mlp.gate_proj.weight", F, D, 4}, {"mlp.up_proj.weight", F, D, 5},
        {"mlp.down_proj.weight", D, F, 6}, {"input_layernorm.weight", D, 0, 7}, {"post_attention_layernorm.weight", D, 0, 8}};
    for (auto& t : tab) {
        if (strcmp(sub, t.name)) continue;
        if (int e = want(t.er, t.ec)) return e;
        if (skip) return NDIT_OK;
        h->seen.insert(key);
        const size_t L = (size_t)l;
        switch (t.kind) {
            case 0: return tx_place(h, h->Wqkv + L * Wq * D, D, src, dtype, r, c, 0, 0, 0, s);
            case 1: return tx_place(h, h->Wqkv + L * Wq * D, D, src, dtype, r, c, 0, 0, HD, s);
            case 2: return tx_place(h, h->Wqkv + L * Wq * D, D, src, dtype, r, c, 0, 0, HD + KV, s);
            case 3: return tx_place(h, h->Wo + L * D * HD, HD, src, dtype, r, c, 0, 0, 0, s);
            case 4: return tx_place(h, h->W13 + L * 2 * F * D, D, src, dtype, r, c, 128, 256, 0, s);      // [128 gate rows | 128 up rows] per 256-row block
            case 5: return tx_place(h, h->W13 + L * 2 * F * D, D, src, dtype, r, c, 128, 256, 128, s);
            case 6: return tx_place(h, h->W2 + L * D * F, F, src, dtype, r, c, 0, 0, 0, s);
            case 7: return tx_place(h, h->ln1 + L * D, D, src, dtype, 1, r, 0, 0, 0, s);
            default: return tx_place(h, h->ln2 + L * D, D, src, dtype, 1, r, 0, 0, 0, s);
        }
    }
    return h->fail(NDIT_ERR_INVALID, "unexpected key %s", key);
}

extern "C" int ntxt_finalize_weights(ntxt_handle h, void* stream) {
    if (!h) return NDIT_ERR_INVALID;
    const char* per_layer[] = {"self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight", "self_attn.o_proj.weight",
                               "mlp.gate_proj.weight", "mlp.up_proj.weight", "mlp.down_proj.weight", "input_layernorm.weight",
                               "post_attention_layernorm.weight"};
    if (!h->seen.count("embed_tokens.weight")) return h->fail(NDIT_ERR_STATE, "missing key embed_tokens.weight");
    for (int l = 0; l < h->Lrun; ++l)
        for (const char* n : per_layer) {
            char k[128];
            snprintf(k, sizeof(k), "layers.%d.%s", l, n);
            if (!h->seen.count(k)) return h->fail(NDIT_ERR_STATE, "missing key %s", k);
        }
    TCK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    h->finalized = true;
    return NDIT_OK;
}

static int tx_plans(ntxt_engine* h, int M) {
    if (M == h->plan_M) return 0;
    const size_t D = h->D, F = h->F, Wq = h->Wq, HD = (size_t)h->H * h->hd, R = h->Lrun;
    h->p_qkv.resize(R); h->p_o.resize(R); h->p_w13.resize(R); h->p_w2.resize(R);
    for (size_t l = 0; l < R; ++l) {
        int e = 0;
        e |= make_gemm_plan(&h->p_qkv[l], h->u, (int)D, h->Wqkv + l * Wq * D, h->qkv, (int)Wq, M, (int)Wq, (int)D, EPI_STORE, h->num_sms);
        e |= make_gemm_plan(&h->p_o[l], h->attn, (int)HD, h->Wo + l * D * HD, h->o, (int)D, M, (int)D, (int)HD, EPI_STORE, h->num_sms);
        e |= make_gemm_plan(&h->p_w13[l], h->u, (int)D, h->W13 + l * 2 * F * D, h->hbuf, (int)F, M, (int)(2 * F), (int)D, EPI_GEGLU, h->num_sms);
        e |= make_gemm_plan(&h->p_w2[l], h->hbuf, (int)F, h->W2 + l * D * F, h->o, (int)D, M, (int)D, (int)F, EPI_STORE, h->num_sms);
        if (e) return h->fail(NDIT_ERR_CUDA, "%s", tmap_last_error());
    }
    h->plan_M = M;
    return 0;
}

extern "C" int ntxt_encode(ntxt_handle h, const int64_t* ids, const int64_t* mask, int32_t batch, int32_t T, void* out, void* stream) {
    if (!h || !ids || !out) return NDIT_ERR_INVALID;
    if (!h->finalized) return h->fail(NDIT_ERR_STATE, "weights not finalized");
    if (batch < 1 || T < 1 || T > 1024 || (int64_t)batch * T > h->Mmax) return h->fail(NDIT_ERR_INVALID, "batch * T must be in 1..%d, T <= 1024", h->Mmax);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int M = batch * T, D = h->D, hd = h->hd, H = h->H, Hkv = h->Hkv;
    if (int e = tx_plans(h, M)) return e;
    if (h->rope_T < T) {
        const int n = T * (hd / 2);
        tx_rope_table_kernel<<<(n + 255) / 256, 256, 0, s>>>(h->rope, T, hd / 2, h->cfg.rope_theta);
        TCK(cudaGetLastError());
        h->rope_T = T;
    }
    const float scale_b = host_bf16_round(sqrtf((float)D));     // embed_scale.to(weight.dtype)
    tx_embed_kernel<<<M, 256, 0, s>>>(reinterpret_cast<const long long*>(ids), h->emb, h->x, M, D, h->cfg.vocab_size, scale_b);
    TCK(cudaGetLastError());
    const dim3 ngrid((M + TX_WARPS - 1) / TX_WARPS), nblock(TX_WARPS * 32);
    const size_t attn_sh = (size_t)TX_WARPS * (hd + T) * sizeof(float);
    static bool attn_cfg = false;
    if (!attn_cfg && attn_sh > 48 * 1024) {
        TCK(cudaFuncSetAttribute(tx_attention_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)TX_WARPS * (hd + 1024) * sizeof(float))));
        attn_cfg = true;
    }
    const float scale = 1.0f / sqrtf((float)hd);
    for (int l = 0; l < h->Lrun; ++l) {
        // input_layernorm (for l > 0 the kernel first folds the previous layer's MLP output into the residual stream)
        tx_resid_rmsnorm_kernel<<<ngrid, nblock, 0, s>>>(h->x, l == 0 ? nullptr : h->o, h->ln1 + (size_t)l * D, h->u, M, D, h->cfg.rms_norm_eps);
        TCK(cudaGetLastError());
        TCK(gemm_bf16_tn(h->p_qkv[l], s));
        {
            const size_t total = (size_t)M * (H + Hkv) * (hd / 2);
            tx_rope_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(h->qkv, h->Wq, h->rope, M, T, H + Hkv, hd);
            TCK(cudaGetLastError());
        }
        tx_attention_kernel<256><<<dim3((T + TX_WARPS - 1) / TX_WARPS, H, batch), TX_WARPS * 32, attn_sh, s>>>(
            h->qkv, h->Wq, reinterpret_cast<const long long*>(mask), h->attn, T, H, Hkv, scale);
        TCK(cudaGetLastError());
        TCK(gemm_bf16_tn(h->p_o[l], s));
        // residual + post_attention_layernorm
        tx_resid_rmsnorm_kernel<<<ngrid, nblock, 0, s>>>(h->x, h->o, h->ln2 + (size_t)l * D, h->u, M, D, h->cfg.rms_norm_eps);
        TCK(cudaGetLastError());
        TCK(gemm_bf16_tn(h->p_w13[l], s));       // gelu_tanh(gate) * up in the epilogue (GemmaMLP.forward :95-97)
        TCK(gemm_bf16_tn(h->p_w2[l], s));
    }
    // last residual addition: hidden_states[-2] = x + mlp output of layer Lrun-1
    tx_resid_rmsnorm_kernel<<<ngrid, nblock, 0, s>>>(h->x, h->o, nullptr, nullptr, M, D, h->cfg.rms_norm_eps);
    TCK(cudaGetLastError());
    TCK(cudaMemcpyAsync(out, h->x, (size_t)M * D * sizeof(bf16), cudaMemcpyDeviceToDevice, s));
    return NDIT_OK;
}
