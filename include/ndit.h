/* ndit.h - C ABI of the B200-native Next-DiT denoising engine (libndit_b200.so).
 *
 * The reference (Alpha-VLLM/Lumina-T2X) has no FFI: its boundary for this path is the Python module
 * API.  Each entry point below names the reference interface it stands behind (paths relative to the
 * reference root).  The Python mirror of that interface lives in lumina_t2x_b200/{models,transport} and
 * reaches this library through ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions: plain C types only; every call returns 0 on success or a negative ndit_status; the
 * message is available from ndit_last_error().  Never aborts.  A handle is bound to the CUDA device that
 * was current at ndit_create(), is not re-entrant, and launches on the stream passed per call
 * (a cudaStream_t cast to void*; NULL = legacy default stream).  "dev" pointers are device pointers owned
 * by the caller; "host" pointers are host memory (pinned recommended).  bf16 = IEEE bfloat16.
 */
#ifndef NDIT_H_
#define NDIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDIT_ABI_VERSION 5

typedef struct ndit_engine* ndit_handle;

typedef enum ndit_status {
    NDIT_OK = 0,
    NDIT_ERR_INVALID = -1,   /* bad argument / unsupported shape */
    NDIT_ERR_CUDA = -2,      /* CUDA runtime / driver error */
    NDIT_ERR_STATE = -3,     /* call order violated (weights not finalized, caption not set, ...) */
    NDIT_ERR_NOMEM = -4
} ndit_status;

typedef enum ndit_dtype { NDIT_BF16 = 0, NDIT_F32 = 1 } ndit_dtype;
typedef enum ndit_method { NDIT_EULER = 0, NDIT_MIDPOINT = 1, NDIT_RK4 = 2 /* torchdiffeq's fixed-grid rk4 (3/8 rule) */ } ndit_method;

/* Architecture of a NextDiT instance: the ctor arguments of
 * lumina_next_t2i/models/model.py:665-741 (NextDiT.__init__) fixed by the factories at :994-999. */
typedef struct ndit_config {
    int32_t dim;            /* 2304 */
    int32_t n_layers;       /* 24   */
    int32_t n_heads;        /* 32   */
    int32_t n_kv_heads;     /* 8 (GQA) or n_heads */
    int32_t cap_feat_dim;   /* 2048 */
    int32_t in_channels;    /* 4 (sdxl VAE latents; 16 = sd3 VAE, lumina_next_t2i/train.py:323); even, 2..16 */
    int32_t patch_size;     /* 2    */
    int32_t multiple_of;    /* 256  */
    int32_t learn_sigma;    /* 1    */
    float norm_eps;         /* 1e-5 */
    int32_t max_tokens;     /* largest H/2*W/2 the workspace is sized for (e.g. 4096 for 1024x1024) */
    int32_t max_cap_len;    /* largest caption length T (e.g. 256) */
    int32_t max_batch;      /* rows of one forward_with_cfg call = 2 (cond + uncond) * samples; <= 4 */
    int32_t num_classes;    /* 0: caption-conditioned Lumina-Next-T2I NextDiT (above).  > 0: class-conditional Next-DiT
                             * (Next-DiT-ImageNet/models/models.py:836-1056 DiT_Llama + TransformerBlockSandwichNorm2):
                             * label embedding table [num_classes + 1, min(dim,1024)] instead of the caption path, no
                             * cross-attention, weight-free pre-norms, final layer with shift + scale; cap_feat_dim and
                             * max_cap_len are ignored. */
    int32_t flag_dit;       /* != 0 (with num_classes == 0): the Flag-DiT of Lumina-T2I, lumina_t2i/models/model.py:661-991
                             * (DiT_Llama; DiT_Llama_5B_patch2 :989-990): caption-conditioned like NextDiT but with a 6-chunk
                             * adaLN (shift, scale, plain gate) :596-609, one weighted RMSNorm per sub-block and no post-norms,
                             * a 1-D RoPE over the token index :925-960, a learned [eol] token closing each row of patches
                             * :779-785 (max_tokens counts them: H/2 * (W/2 + 1)) and shift + scale in the final layer
                             * :655-656.  head_dim = dim / n_heads must be 72, 48 or 96 in all variants. */
    int32_t moe_time_experts;   /* class-conditional model only (Next-DiT-MoE/models/): mixture-of-experts FFN, top-2.   */
    int32_t moe_space_experts;  /* (8, 0): models.py, gate = Linear(timestep embedding) :451-477 -> one pair of experts per
                                 * sample and layer; (0, 8): models1.py, gate = Linear(token) :451-477; (4, 4): models2.py,
                                 * time MoE then space MoE, 6-chunk adaLN :451-506,760-808.  (0, 0): dense FFN.  Expert
                                 * outputs are accumulated in expert-index order in bf16, like the reference's
                                 * ``results[idx] += w * expert(x[idx])``; top-k ties go to the lower expert index. */
    int32_t ffn_dim;        /* 0: FeedForward's hidden width is derived as in model.py:470-473 without ffn_dim_multiplier,
                             * multiple_of * ceil(int(2 * 4 * dim / 3) / multiple_of).  > 0: the hidden width itself (the host computes
                             * int(ffn_dim_multiplier * int(2 * 4 * dim / 3)) rounded up to multiple_of in Python's arithmetic, :471-473);
                             * a multiple of 128. */
    int32_t no_qk_norm;     /* != 0: the qk_norm=False architecture (model.py:219-220: q_norm = k_norm = ky_norm = Identity): no
                             * q/k LayerNorm before the rotary embedding, no LayerNorm on the caption keys, no *_norm keys in the state dict */
} ndit_config;

/* Per-call arguments of NextDiT.forward_with_cfg (model.py:866-913) that are not tensors. */
typedef struct ndit_step_params {
    float cfg_scale;
    float scale_factor;      /* time-aware RoPE scaling (model.py:944-952) */
    float scale_watershed;
    int32_t proportional_attn;
    int32_t base_seqlen;     /* used when proportional_attn != 0 (model.py:373-376) */
    float ntk_factor;        /* class-conditional model and Flag-DiT: DiT_Llama.forward_with_cfg(rope_scaling_factor=scale_factor,
                              * ntk_factor=...) (models.py:946-1012; lumina_t2i model.py:868-899); 0 is treated as 1 */
} ndit_step_params;

/* --- lifecycle: models.NextDiT_2B_GQA_patch2(...) / .to("cuda") / del (sample.py:125-129) */
int ndit_abi_version(void);
int ndit_create(const ndit_config* cfg, ndit_handle* out);
int ndit_destroy(ndit_handle h);
const char* ndit_last_error(ndit_handle h);   /* h may be NULL: last error of a failed ndit_create */

/* Grow the workspace of an existing handle (never shrinks; weights stay packed): the reference allocates activations per
 * call, so an unmodified caller may pass a larger latent (e.g. 2048x2048 = 16384 tokens) or a longer caption than the handle
 * was created for - the Python mirror calls this instead of failing.  Invalidates the caption / label state (set it again). */
int ndit_reserve(ndit_handle h, int32_t max_tokens, int32_t max_cap_len, int32_t max_batch);

/* --- weights: nn.Module.load_state_dict(strict=True) (sample.py:135-142).  `key` is the reference
 * state-dict key; the tensor is copied (bf16 or f32 source, row-major) into
 * engine-owned, GEMM-ready storage.  ndit_finalize_weights fails if a key is missing (strict). */
int ndit_set_weight(ndit_handle h, const char* key, const void* dev_ptr, const int64_t* shape, int32_t ndim,
                    int32_t dtype, void* stream);
int ndit_finalize_weights(ndit_handle h, void* stream);
int64_t ndit_parameter_count(ndit_handle h);   /* NextDiT.parameter_count (model.py:965-982) */
/* Checkpoint tooling for this engine (the reference converts .pth <-> .safetensors, lumina_next_t2i/entry_point.py:115-156):
 * the finalized weights in the engine's own GEMM-ready layout, written to / read from one flat file.  Loading replaces
 * ndit_set_weight x N + ndit_finalize_weights on a cold start (no re-packing kernels, disk reads overlapped with the H2D
 * copies); the file is tied to the architecture (ndit_config without the workspace limits) and to the ABI version. */
int ndit_save_packed(ndit_handle h, const char* path);
int ndit_load_packed(ndit_handle h, const char* path);

/* --- caption conditioning: the cap_feats / cap_mask kwargs of forward_with_cfg.  They are constant over
 * an ODE solve, so the caption-side work (pooling + cap_embedder, model.py:847-850; attention_y_norm,
 * wk_y/wv_y, ky_norm for all layers, :421-422,:602) is done once here.
 * cap_feats_dev: bf16 [batch, T, cap_feat_dim]; cap_mask_dev: uint8 [batch, T] (non-zero = valid). */
int ndit_set_caption(ndit_handle h, const void* cap_feats_dev, const uint8_t* cap_mask_dev, int32_t batch, int32_t T,
                     void* stream);

/* --- region-masked captions of the compositional model: the cap_feats / cap_mask / global_cap_feats / global_cap_mask / h_split_num /
 * w_split_num kwargs of lumina_next_compositional_generation/models/model.py:902-953 (NextDiT.forward_with_cfg) -> :852-899 (forward).
 * cap_feats_dev: bf16 [n_caps, T, cap_feat_dim] = the region captions followed by the negative (unconditional) caption (demo.py:211-223),
 * cap_mask_dev: uint8 [n_caps, T]; global_cap_feats_dev bf16 [1, global_T, cap_feat_dim] / global_cap_mask_dev uint8 [1, global_T]: the
 * caption the adaLN conditioning pools over for BOTH rows (:866-870).  The latent's token grid is cut into h_split x w_split rectangles
 * of (H // h_split // 2) x (W // w_split // 2) tokens; rectangle (i, j) belongs to caption (i + 1) * (j + 1) - 1 (:879) and its tokens
 * cross-attend to that caption only (Attention.forward :421-446: per-caption masked SDPA, nan_to_num, sum over the cond captions); every
 * token of the unconditional row attends to the last caption.  Drives ndit_forward_cfg / ndit_sample / ndit_sample_sde (and ndit_forward:
 * the guidance-free forward :852-899, row 0 = cond, row 1 = uncond) with batch = 2 (one cond / uncond pair) until the next ndit_set_caption.  Grows the caption buffers when n_caps exceeds what the workspace holds.  head_dim 72. */
int ndit_set_caption_regions(ndit_handle h, const void* cap_feats_dev, const uint8_t* cap_mask_dev, int32_t n_caps, int32_t T,
                             const void* global_cap_feats_dev, const uint8_t* global_cap_mask_dev, int32_t global_T, int32_t h_split,
                             int32_t w_split, void* stream);

/* --- class-conditional model: the `y` argument of DiT_Llama.forward_with_cfg (models.py:946): labels_dev int64 [batch]
 * (second half = the null class num_classes for CFG).  Replaces ndit_set_caption for num_classes > 0. */
int ndit_set_labels(ndit_handle h, const int64_t* labels_dev, int32_t batch, void* stream);

/* --- NextDiT.forward_with_cfg (model.py:866-913).  x_dev/out_dev: bf16 [batch, C, height, width] NCHW
 * (latent size; batch = 2 * samples, second half of x ignored as in the reference); t is the (common)
 * timestep.  Uses the caption set by ndit_set_caption (same batch). */
int ndit_forward_cfg(ndit_handle h, const void* x_dev, float t, int32_t batch, int32_t height, int32_t width,
                     const ndit_step_params* sp, void* out_dev, void* stream);

/* --- NextDiT.forward (model.py:836-864): no classifier-free guidance.  x_dev/out_dev: bf16 [batch, C, height, width]; every row is
 * its own sample with its own timestep t_host[b] (HOST array of `batch` floats) and its own caption row (ndit_set_caption with
 * the same batch, 1 <= batch <= max_batch).  out = the first C of the 2C output channels (learn_sigma chunk, :859-861).
 * The reference uses whatever RoPE table / attention scaling the module currently holds (self.freqs_cis, set by __init__ or by the
 * last forward_with_cfg; layer.attention.base_seqlen / proportional_attn): sp->scale_factor = its linear factor, sp->ntk_factor =
 * its NTK factor (0 = 1.0), sp->proportional_attn / base_seqlen as stored; sp->cfg_scale and sp->scale_watershed are ignored. */
int ndit_forward(ndit_handle h, const void* x_dev, const float* t_host, int32_t batch, int32_t height, int32_t width,
                 const ndit_step_params* sp, void* out_dev, void* stream);

/* --- NextDiT.forward with a LIST of latents of different sizes (model.py:789-834, 836-864): x_dev / out_dev are HOST arrays of
 * `batch` device pointers, bf16 [C, heights[i], widths[i]].  Rows are padded to the longest one with the learned pad token, every
 * row gets the rope positions of its own (H/2, W/2) grid, keys beyond a row's own tokens are masked in the self-attention (the
 * reference's flash-attn varlen path), and proportional attention uses the padded length, as the reference does.  Otherwise
 * like ndit_forward (caption rows from ndit_set_caption, t_host[batch], sp as there).  Text-conditioned Next-DiT only. */
int ndit_forward_list(ndit_handle h, const void* const* x_dev, const int32_t* heights, const int32_t* widths, const float* t_host,
                      int32_t batch, const ndit_step_params* sp, void* const* out_dev, void* stream);

/* --- transport.Sampler.sample_ode(...)(z, model.forward_with_cfg, **kw) (transport/transport.py:346-391,
 * transport/integrators.py:79-116) with torchdiffeq's fixed-grid euler / midpoint / rk4.  t_grid_host: the
 * n_grid time points (fp32, host).  z_dev: bf16 initial state [batch,C,height,width]; traj_dev: bf16
 * [n_grid, batch, C, height, width] receiving every grid state (traj[0] = z), or NULL to keep only
 * the final state, which is always written to final_dev (may alias z_dev). */
int ndit_sample(ndit_handle h, const void* z_dev, int32_t batch, int32_t height, int32_t width,
                const float* t_grid_host, int32_t n_grid, int32_t method, const ndit_step_params* sp,
                void* traj_dev, void* final_dev, void* stream);

/* --- transport.Sampler.sample_sde(...)(init, model_fn, **kw) (transport/transport.py:285-344, integrators.py:5-76), velocity model on
 * the Linear path: the stochastic loop inside the engine.  The reference draws torch.randn on the host in every step; the caller draws
 * the same tensors in the same order and passes them as noise_dev [n_steps][batch,C,H,W] bf16.  Every PyTorch op of the reference rounds
 * to bf16, so the caller also supplies the t-dependent scalars of every evaluation point, computed with the same tensor ops on a
 * one-element bf16 tensor (floats that hold bf16 values): Euler-Maruyama (method 0): pts[i] = point t_i; Heun (method 1): pts[2i] = t_i,
 * pts[2i+1] = t_i + dt.  dt, sqrt_dt, half_dt enter the products as fp32 scalars: the reference multiplies CUDA bf16 tensors with 0-dim
 * CPU tensors, which a CUDA op keeps in fp32 (a CPU op would round them to bf16 first).  traj_dev receives the state after every step
 * [n_steps][batch,C,H,W] (the reference's list xs, without its fp32 last step, which stays with the caller). */
typedef struct ndit_sde_point {
    float t;                /* timestep the model sees */
    float ratio;            /* alpha_t / d_alpha_t */
    float var;              /* sigma_t^2 - ratio * d_sigma_t * sigma_t */
    float diffusion;        /* compute_diffusion(t, form, norm) */
    float sqrt_2diffusion;  /* sqrt(2 * diffusion) */
} ndit_sde_point;
int ndit_sample_sde(ndit_handle h, const void* z_dev, int32_t batch, int32_t height, int32_t width, int32_t n_steps, int32_t method,
                    const ndit_sde_point* pts, float dt, float sqrt_dt, float half_dt, const void* noise_dev,
                    const ndit_step_params* sp, void* traj_dev, void* stream);

/* Same solve with HOST buffers (the end-to-end entry point): copies z / caption host->device, runs
 * ndit_set_caption + ndit_sample, copies the final latent device->host and synchronises the stream. */
int ndit_sample_host(ndit_handle h, const void* z_host, const void* cap_feats_host, const uint8_t* cap_mask_host,
                     int32_t batch, int32_t height, int32_t width, int32_t T, const float* t_grid_host, int32_t n_grid,
                     int32_t method, const ndit_step_params* sp, void* final_host, void* stream);

/* --- instrumentation */
int64_t ndit_launch_count(ndit_handle h);          /* kernels launched by this handle so far (a graph replay counts the launches it contains) */
int64_t ndit_graph_replay_count(ndit_handle h);    /* ndit_sample calls that ran as ONE CUDA-graph launch: a solve is recorded the second time
                                                    * it is requested with the same shape / grid / parameters and replayed from then on */
int ndit_set_option(ndit_handle h, const char* name, int32_t value);
/* options: "attn_ref" = 1: CUDA-core debug attention kernel; "attn_tp" = 1: experimental attention kernel with P in tensor
 * memory (head_dim 72); "pdl" = 1: programmatic dependent launch for the hot-loop kernels (process-wide); "profile" = 1:
 * record a CUDA-event pair around every kernel launch, read back with ndit_profile_read. */
/* Sums the per-launch CUDA-event durations recorded since "profile" was switched on (or since the last read),
 * per kernel class: 0 gemm_qkv, 1 gemm_wo, 2 gemm_w13(swiglu), 3 gemm_w2, 4 attention, 5 row-wise, 6 conditioning. */
#define NDIT_PROFILE_CLASSES 7
int ndit_profile_read(ndit_handle h, float* ms_out, int64_t* count_out, int32_t n_classes);

/* Debug tap for per-block parity tests: with ndit_set_option(h, "tap_layer", l) every forward copies the residual stream x after
 * TransformerBlock l (model.py:624, [batch * tokens, dim] bf16, token-major) aside; this reads it back (device to device). */
int ndit_debug_read_residual(ndit_handle h, void* out_dev, int64_t rows, void* stream);

/* --- single-operator entry points (parity tests and micro-benchmarks call the kernels through these) */
/* C[M,N] = A[M,K] W[N,K]^T, bf16, fp32 accumulate.  swiglu != 0: W is [2F,K] block-interleaved
 * (128 rows w1 | 128 rows w3) and C is [M,F] = silu(a)*b.  (F.linear call sites: model.py:358,438,502) */
int ndit_op_gemm(const void* A_dev, const void* W_dev, void* C_dev, int32_t M, int32_t N, int32_t K, int32_t swiglu,
                 void* stream);
/* micro-benchmark of the same GEMM: `iters` launches after a warm-up, *ms_out = average device time per launch.
 * allow_pair = 0 forces the single-CTA kernel.  Returns 1 if the CTA-pair kernel ran, 0 if the single-CTA one, < 0 on error. */
int ndit_op_gemm_bench(const void* A_dev, const void* W_dev, void* C_dev, int32_t M, int32_t N, int32_t K, int32_t swiglu,
                       int32_t allow_pair, int32_t iters, float* ms_out, void* stream);
/* in place on qkv [M, (H+2Hkv)*hd]: q,k <- bf16(rope(LayerNorm(.))) (model.py:361-371); angles are built from
 * (Hp, Wp, theta, linear_factor) like precompute_freqs_cis (:916-963). M = batch*Hp*Wp.
 * one_d != 0: the 1-D table of Flag-DiT over Hp*Wp token positions (lumina_t2i model.py:925-960), linear_factor = rope
 * scaling factor */
int ndit_op_ln_rope(void* qkv_dev, const void* qw, const void* qb, const void* kw, const void* kb, int32_t batch,
                    int32_t Hp, int32_t Wp, int32_t H, int32_t Hkv, int32_t hd, float theta, float linear_factor,
                    int32_t one_d, void* stream);
/* fused self + gated cross attention (model.py:373-434).  qkv [B*N,(H+2Hkv)*72] (q,k already normed+roped),
 * kvy [B*T, 2*Hkv*72] (ky normed | vy), ymask uint8 [B,T], gate_tanh f32 [H]; out bf16 [B*N, H*72].
 * use_ref: 0 = default tcgen05 kernel (environment NDIT_ATTN_GEN = 1 / 3 overrides), 1 = CUDA-core reference kernel,
 * 2 = third-generation kernel (P in tensor memory, two softmax threads per row; engine option "attn_gen" = 3),
 * 3 = first-generation kernel (one softmax thread per row, P through shared memory; "attn_gen" = 1). */
int ndit_op_attention(const void* qkv_dev, const void* kvy_dev, const uint8_t* ymask_dev, const float* gate_tanh_dev,
                      void* out_dev, int32_t B, int32_t N, int32_t T, int32_t H, int32_t Hkv, float scale_self,
                      float scale_cross, int32_t use_ref, void* stream);
/* general form: head_dim hd = 72, 48 or 96; T = 0 (kvy/ymask/gate may be NULL) = no caption segment */
int ndit_op_attention_hd(const void* qkv_dev, const void* kvy_dev, const uint8_t* ymask_dev, const float* gate_tanh_dev,
                         void* out_dev, int32_t B, int32_t N, int32_t T, int32_t H, int32_t Hkv, int32_t hd, float scale_self,
                         float scale_cross, int32_t use_ref, void* stream);
/* same op, launched `iters` times after a warm-up; *ms_out = average device time per launch (CUDA events) */
int ndit_op_attention_bench(const void* qkv_dev, const void* kvy_dev, const uint8_t* ymask_dev, const float* gate_tanh_dev,
                            void* out_dev, int32_t B, int32_t N, int32_t T, int32_t H, int32_t Hkv, float scale_self,
                            float scale_cross, int32_t iters, float* ms_out, void* stream);
/* token gate of the mixture-of-experts FFN (Next-DiT-MoE models1.py:461-470): logits = bf16(u Wg^T) (fp32 accumulate), top-2 per
 * token (ties: lower expert index), weights = bf16(softmax over the two).  u bf16 [M, D], Wg bf16 [E, D], 2 <= E <= 8;
 * wtok bf16 [M, E]: the weight of each selected expert, 0 for the others. */
int ndit_op_moe_gate(const void* u_dev, const void* Wg_dev, void* wtok_dev, int32_t M, int32_t D, int32_t E, void* stream);
/* X += tanh_g * RMS(o; w_post) (skipped if o NULL; w_post NULL: X += tanh_g * o, Flag-DiT);
 * u = RMS(X; w_pre) * onepls (+ shift if not NULL)   (model.py:597-610; lumina_t2i model.py:596-609);
 * tanh_g / onepls / shift: bf16 [M / rows_per_batch, D] */
int ndit_op_resid_rms_mod(void* X_dev, const void* o_dev, const void* w_post, const void* tanh_g, const void* w_pre,
                          const void* onepls, const void* shift, void* u_dev, int32_t M, int32_t rows_per_batch, int32_t D,
                          float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NDIT_H_ */
