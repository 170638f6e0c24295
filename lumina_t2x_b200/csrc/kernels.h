// Internal (C++) launch API of the sm_100a kernels.  The public boundary is include/ndit.h.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ndit {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- tensor maps (tensormap.cu)
// bf16 2-D map over a row-major [rows, cols] matrix with row stride ld (elements):
// box = box_rows x box_cols, swizzle = 128B (box_cols*2 must be 128) .
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_rows, uint32_t box_cols, int swizzle_bytes);
// bf16 3-D map: dims (d0 fastest, d1, d2) with byte strides s1, s2 (multiples of 16), box (b0,b1,b2).
int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes,
                 uint64_t s2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, int swizzle_bytes);
const char* tmap_last_error();

// ---------------------------------------------------------------- GEMM (gemm_tcgen05.cu)
enum { EPI_STORE = 0, EPI_SWIGLU = 1, EPI_GEGLU = 2 };   // gated epilogues: silu(a) * b (Next-DiT FFN) / gelu_tanh(a) * b (Gemma MLP)
inline bool epi_gated(int epi) { return epi == EPI_SWIGLU || epi == EPI_GEGLU; }
// Optional second destination of an EPI_STORE GEMM (the fused q|k|v projection): output columns >= col0 are the value heads and go,
// transposed, straight into the attention kernel's V^T buffer [batch * Hkv][vrows][npad] (replaces transpose_v and the 2 x 9.4 MB
// round trip through the qkv buffer).  ptr == nullptr: plain store.
struct GemmVtOut {
    bf16* ptr;
    int col0;     // first value column = (H + Hkv) * hd
    int hd, hkv, vrows, npad, ntok;
};
// Optional device-side row window (token-routed mixture of experts: how many gathered rows an expert owns is only known on the
// device): the kernel works on rows [*offset, *offset + *count) of A and C, count <= the plan's M (tiles past it are skipped).
// count == nullptr: rows [0, M).
struct GemmRowWin {
    const int* count;
    const int* offset;
};
// Optional extensions of the CTA-pair EPI_STORE GEMM for the VAE decoder (vae_decoder.cu).  The image stack lives in HBM as
// [batch][hp][wp][C] bf16 with a one-pixel zero border, i.e. as a [batch * hp * wp, C] matrix: a 3x3 convolution with padding 1 is
// then nine accumulating GEMMs whose A tiles are the same rows shifted by (dy * wp + dx) - the producer only offsets the TMA row
// coordinate (rows before / after the matrix are zero-filled) and W is packed [Cout][tap][Cin].
struct GemmExt {
    int kpt;             // > 0: implicit 3x3 convolution, k-blocks per tap (= Cin / 64); 0: plain GEMM
    int share;           // 1 (with kpt > 0): one (128 + 8)-row A tile per (dy, k-block) serves the three dx taps (A map box = 136 rows):
                         // A crosses L2 -> shared memory 3 times per tile instead of 9
    int wp, hp;          // padded width / height; hp > 0: rows on the border of their image are stored as zero
    const bf16* bias;    // optional [N]: bf16(acc + bias)
    const bf16* resid;   // optional [M, ldr]: out = bf16(bf16(acc + bias) + resid)   (ResnetBlock2D / Attention residual)
    int ldr;
    float* out_f32;      // non-null: the fp32 accumulators go to out_f32[M, ldc] instead of C (attention scores)
};
struct GemmPlan {
    CUtensorMap tmA;  // A [M,K], box 128 x 64
    CUtensorMap tmB;  // W [N,K], box bn  x 64
    bf16* C;
    const int* w_row_off;   // optional device int: W is a stack of [N, K] matrices, use the one at row (*w_row_off) * w_row_mul
    int w_row_mul;
    const bf16* bias;   // optional [N] bias added to the fp32 accumulator (single-CTA kernel, EPI_STORE only); nullptr: none
    int M, N, K, ldc;
    int bn;   // 128 or 256
    int pair; // 1: CTA-pair kernel (cta_group::2, 256x256 tiles; W box is 128 rows)
    int epi;  // EPI_*
    int num_sms;
    GemmVtOut vt;   // zero-initialised by make_gemm_plan
    GemmRowWin rows;
    int use_ext;    // 1: ext applies (pair kernel, EPI_STORE; bn 256 or 128)
    GemmExt ext;
};
cudaError_t gemm_bf16_tn(const GemmPlan& p, cudaStream_t stream);
// builds the maps of a plan (A: [M,K] ld=lda; W: [N,K] ld=K)
// w_total_rows > N: W is a stack of matrices (see GemmPlan::w_row_off); the W map then spans all of them
int make_gemm_plan(GemmPlan* p, const bf16* A, int lda, const bf16* W, bf16* C, int ldc, int M, int N, int K, int epi,
                   int num_sms, int allow_pair = 1, int w_total_rows = 0);

// ---------------------------------------------------------------- attention (attention_tcgen05.cu)
// Region-masked caption cross-attention of the compositional model (lumina_next_compositional_generation/models/model.py:421-446,
// 872-887): the caption buffers hold n_cond region captions followed by the unconditional one; the latent's token grid (Wp tokens per
// row) is cut into hs x ws rectangles of hp x wp tokens and rectangle (i, j) belongs to caption (i + 1) * (j + 1) - 1.  n_cond == 0: off.
struct AttnRegion {
    int n_cond;
    int Wp, hp, wp, hs, ws;
};
struct AttnPlan {
    CUtensorMap tmQ64, tmQ16;    // q  : dims (hd, H,   B*N) on the qkv buffer, boxes (64,1,128) / (16,1,128)
    CUtensorMap tmK64, tmK16;    // k  : dims (hd, Hkv, B*N)
    CUtensorMap tmVt;            // v^T: dims (N, 80, B*Hkv), box (64, 80, 1); row 72 = ones
    CUtensorMap tmKy64, tmKy16;  // ky : dims (hd, Hkv, B*T)
    CUtensorMap tmVyt;           // vy^T: dims (Tpad, 80, B*Hkv), box (64, 80, 1); row 72 = ones
    const uint8_t* ymask;        // [B, T] bytes (0/1)
    const int* kv_len;           // optional device [B]: valid image tokens of each batch row (variable-resolution list input: rows are
                                 // padded to N tokens, keys beyond kv_len[b] are masked, model.py:387-404 / flash-attn varlen); nullptr: N
    const float* gate_tanh;      // [H] bf16-rounded tanh(gate)
    bf16* out;                   // [B*N, H*hd]
    int B, N, T, H, Hkv, hd;     // T = 0: no caption segment (class-conditional model); hd = 72, 48 or 96
    float scale_self, scale_cross;
    int bkv;                     // kv rows per K box: 128 (attention_fused) or attention_hr_bkv(hd) (attention_fused_hr)
    AttnRegion region;           // n_cond > 0: region-masked captions (attention_fused and attention_ref only; B = 2)
};
// first-generation kernel: one softmax thread per row, P through shared memory (attention_tcgen05.cu)
cudaError_t attention_fused(const AttnPlan& p, cudaStream_t stream);
// third-generation kernel: P in tensor memory (aliased onto S), two softmax threads per row (attention_hr_tcgen05.cu);
// attention_hr_bkv(hd) = its kv block size for this head_dim (the K boxes of the plan must be built with it)
int attention_hr_bkv(int hd);
cudaError_t attention_fused_hr(const AttnPlan& p, cudaStream_t stream);
// slow CUDA-core reference of the same op (debug / NDIT_ATTN=ref); same inputs in plain layouts
cudaError_t attention_ref(const bf16* qkv, int ld_qkv, const bf16* kvy, int ld_kvy, const uint8_t* ymask,
                          const float* gate_tanh, bf16* out, int B, int N, int T, int H, int Hkv, int hd,
                          float scale_self, float scale_cross, cudaStream_t stream, const int* kv_len = nullptr,
                          AttnRegion region = AttnRegion{0, 0, 0, 0, 0, 0});

// ---------------------------------------------------------------- row-wise kernels (rowwise.cu)
// X[token, :] = bf16(patch(x[b % n]) . Wx^T + bx)
// eol != nullptr (Flag-DiT): every row of patches is closed by the learned [eol] token -> Hp * (Wp + 1) tokens
cudaError_t patch_embed(const bf16* x, const bf16* Wx, const bf16* bx, const bf16* eol, bf16* X, int B, int n_unique, int C,
                        int Hh, int Ww, int D, cudaStream_t s);
// tf[b, 0:256] = bf16(sinusoid(t)); pool[b, :] = bf16(LN(masked mean of cap[b]))  (fp32 storage)
cudaError_t cond_prepare(float t, const float* t_rows, const bf16* cap, const uint8_t* mask, const bf16* ln_w, const bf16* ln_b, float* tf,
                         float* pool, int B, int T, int C, int do_caption, cudaStream_t s);
enum { POST_NONE = 0, POST_SILU = 1, POST_ADALN = 2 };
enum { ADALN_NEXT = 0, ADALN_CLASS = 1, ADALN_FLAG = 2 };   // chunk layout of the packed adaLN output (see gemv_rows_kernel)
// out[b,o] = post(bf16(sum_k in'[b,k] W[o,k] + bias[o]) (+ addend[b,o]));  in' = silu(in) if in_silu
// result goes to out_b (bf16) when non-null, else to out (fp32 storage of bf16 values)
cudaError_t gemv_rows(const float* in, const bf16* W, const bf16* bias, const float* addend, float* out, bf16* out_b,
                      int B, int O, int K, int in_silu, int post, int adaln_D, int adaln_blocks, int adaln_kind, cudaStream_t s);
// optional residual update  X += tanh_g * RMS(o; w_post)  (w_post == nullptr: X += tanh_g * o)
// then   u = RMS(X; w_pre) * onepls (+ shift)
cudaError_t resid_rms_mod(bf16* X, const bf16* o, const bf16* w_post, const bf16* tanh_g, const bf16* w_pre,
                          const bf16* onepls, const bf16* shift, bf16* u, int M, int rows_per_batch, int D, int mod_stride,
                          float eps, cudaStream_t s);
// last gated residual update then xn = bf16(LN(no affine, eps 1e-6)(X) * onepls (+ shift)) -> xn [M, D]; the Linear(D->O)+bias
// of the final layer runs as a tcgen05 GEMM with a bias epilogue on xn (shift: class-conditional model / Flag-DiT)
cudaError_t final_norm(const bf16* X, const bf16* o, const bf16* w_post, const bf16* tanh_g, const bf16* onepls,
                       const bf16* shift, bf16* xn, int M, int rows_per_batch, int D, int mod_stride, float eps, cudaStream_t s, bf16* x_out = nullptr);
cudaError_t gather_label_rows(const bf16* table, const long long* labels, float* out, int B, int n_rows, int width, cudaStream_t s);
// rope table [N][hd/2] (cos,sin)
// dst[r][:] = src[:] for r in [0, rows): row_bytes a multiple of 16 (pad tokens / pad rope rows of the list input, model.py:811-826)
cudaError_t broadcast_row(void* dst, const void* src, int rows, int row_bytes, cudaStream_t s);
cudaError_t rope_table(float2* tab, int Hp, int Wp, int hd, float theta, float linear_factor, int one_d, cudaStream_t s);
// in place on qkv [M, ld]: q = bf16(rope(LN(q))), k = bf16(rope(LN(k)))
cudaError_t ln_rope_qk(bf16* qkv, int ld, const bf16* qw, const bf16* qb, const bf16* kw, const bf16* kb,
                       const float2* rope, int M, int N_tokens, int H, int Hkv, int hd, cudaStream_t s);
// qk_norm=False: rotary embedding only, in place on the q|k columns of qkv [M, ld]
cudaError_t rope_qk(bf16* qkv, int ld, const float2* rope, int M, int N_tokens, int H, int Hkv, int hd, cudaStream_t s);
// in place LayerNorm(width) + affine on rows of a [M, ld] matrix (no rope), layer-batched via blockIdx.y
cudaError_t ln_rows(bf16* x, int ld, size_t layer_stride_x, const bf16* w, const bf16* b, size_t layer_stride_w,
                    int M, int width, int layers, cudaStream_t s);
// out[l][row,:] = RMS(y[row,:]; w[l]) for all layers
cudaError_t rms_rows_layers(const bf16* y, const bf16* w, bf16* out, int M, int C, int layers, float eps, cudaStream_t s);
// dst[(b*G + g)*grows + d][n] = src[(b*N + n)*ld + col0 + g*hd + d]   (v -> v^T), layers via blockIdx.z
cudaError_t transpose_v(const bf16* src, int ld, int col0, size_t src_layer_stride, bf16* dst, int ld_dst,
                        size_t dst_layer_stride, int B, int N, int G, int hd, int grows, int layers, cudaStream_t s);
// dst[group*grows + hd][0:n_cols] = 1  (the all-ones row of V^T: column hd of P.V becomes the softmax row sum)
cudaError_t fill_ones_row(bf16* dst, int ld_dst, size_t dst_layer_stride, int groups, int n_cols, int hd, int grows,
                          int layers, cudaStream_t s);
// rows per (batch, kv head) group in the V^T buffers: head_dim data rows + the all-ones row, padded to a multiple of 16
__host__ __device__ constexpr int attn_vrows(int hd) { return (hd + 1 + 15) / 16 * 16; }
// unpatchify + learn_sigma slice + 3-channel CFG combine (+ optional fused Euler update)
//   tok [2n*N, O] bf16 -> v [2n,4,Hh,Ww] bf16
cudaError_t unpatchify_plain(const bf16* tok, bf16* v_out, int n, int C, int Hh, int Ww, int O, int eol, cudaStream_t s);
cudaError_t unpatchify_cfg(const bf16* tok, bf16* v_out, int n, int C, int Hh, int Ww, int O, float cfg_scale, int eol,
                           cudaStream_t s);
// mixture-of-experts (class-conditional Next-DiT-MoE): token gate + expert-order bf16 accumulation (see rowwise.cu)
cudaError_t moe_space_gate(const bf16* u, const bf16* Wg, bf16* wtok, int M, int D, int E, cudaStream_t s);
// uniform_w: E weights in DEVICE memory (time-gated layer: written by moe_time_select)
// token-routed experts as a grouped GEMM: segment bookkeeping + gather (see moe_route_count_kernel) and the routed combine
cudaError_t moe_route(const bf16* u, const bf16* wtok, int* cnt, int* cntp, int* off, int* cursor, int* pos, bf16* u_perm, int M, int D, int E,
                      cudaStream_t s);
cudaError_t moe_combine_routed(const bf16* o_perm, const int* pos, const bf16* wtok, bf16* out, int M, int D, int E, cudaStream_t s);
cudaError_t moe_combine(const bf16* oe, size_t estride, int E, const bf16* wtok, const float* uniform_w, bf16* out, int M, int D,
                        cudaStream_t s);
// time gate (Next-DiT-MoE models.py:459-477): per layer the top-2 experts of the gate logits of batch row 0 (ascending expert
// index = accumulation order) and their bf16-rounded softmax weights: sel [L][2], w [L][2] in device memory
cudaError_t moe_time_select(const float* logits, int L, int E, int* sel, float* w, cudaStream_t s);
// one stage of torchdiffeq's fixed-grid rk4 (3/8 rule) on a bf16 state (see rowwise.cu); dt already rounded to bf16
cudaError_t rk4_stage(int stage, bf16* out, const bf16* y, const bf16* k1, const bf16* k2, const bf16* k3, const bf16* k4, float dt,
                      size_t count, cudaStream_t s);
// one elementwise stage of the SDE samplers on a bf16 state (see sde_step_kernel)
cudaError_t sde_step(int mode, bf16* out, bf16* kout, const bf16* x, const bf16* v, const bf16* w, const bf16* a, const bf16* kin,
                     float ratio, float var, float diffusion, float sqrt2d, float dt, float sqrt_dt, float half_dt, size_t count,
                     cudaStream_t s);
cudaError_t axpy_bf16(bf16* y_out, const bf16* y_in, const bf16* v, float dt, size_t count, cudaStream_t s);

}  // namespace ndit
