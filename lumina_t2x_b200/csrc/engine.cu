// C-ABI engine (include/ndit.h): weight packing, workspace, the per-block kernel schedule of
// NextDiT.forward_with_cfg and the fixed-grid ODE loop.  No torch types; plain CUDA runtime.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/ndit.h"
#include "kernels.h"
#include "launch.cuh"

using namespace ndit;

namespace ndit {
// programmatic dependent launch for the hot-loop kernels (launch.cuh): per engine (option "pdl", default NDIT_PDL), copied into
// this thread-local for the duration of an engine call.  Measured: no gain on the power-capped B200 (934.4 vs 934.5 ms / latent).
thread_local int g_pdl = 0;
}  // namespace ndit

namespace {

inline float host_bf16_round(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return x;   // NaN
    const uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    u &= 0xffff0000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}
inline float host_bf16_to_float(uint16_t h) {
    uint32_t u = static_cast<uint32_t>(h) << 16;
    float y;
    memcpy(&y, &u, 4);
    return y;
}

// dst[(r / blk) * dst_blk_stride + r % blk + dst_row0][c] = src[r][c]   (bf16 or f32 source)
__global__ void place_rows_kernel(bf16* __restrict__ dst, size_t dst_ld, const void* __restrict__ src, int src_f32,
                                  size_t rows, size_t cols, size_t blk, size_t dst_blk_stride, size_t dst_row0) {
    const size_t total = rows * cols;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t r = i / cols, c = i % cols;
        const size_t dr = (r / blk) * dst_blk_stride + (r % blk) + dst_row0;
        const float v = src_f32 ? static_cast<const float*>(src)[i] : __bfloat162float(static_cast<const bf16*>(src)[i]);
        dst[dr * dst_ld + c] = __float2bfloat16_rn(v);
    }
}

__global__ void mask_to_u8_kernel(uint8_t* dst, const uint8_t* src, size_t n) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] ? 1 : 0;
}

struct RopeSlot {
    int Hp = 0, Wp = 0;
    float theta = 0.f, lin = 0.f;
    float2* tab = nullptr;
};

}  // namespace

// attention kernel generation used when the option / environment says 0
constexpr int ATTN_DEFAULT_GEN = 1;
static inline int attn_gen(int opt) { return opt == 1 || opt == 3 ? opt : ATTN_DEFAULT_GEN; }

struct ndit_engine {
    ndit_config cfg;
    int D, L, H, Hkv, hd, F, C, cd, O, Wq;   // Wq = fused qkv width
    int vrows = 80;                          // V^T rows per (batch, kv head): head_dim + ones row, padded to 16
    bool cls = false;                        // class-conditional variant (DiT_Llama): labels instead of captions
    bool flag = false;                       // Flag-DiT (Lumina-T2I DiT_Llama): shift/scale/gate adaLN, 1-D RoPE, eol tokens
    int NCH = 4;                             // adaLN chunks per layer: 4 (scale, gate) x2, 6 (shift, scale, gate) x2, 6 = MoE "both"
    // FFN sub-blocks of one layer (Next-DiT-MoE): kind 0 dense, 1 time-gated MoE, 2 token-gated MoE
    int NF = 1;
    int ffn_kind[2] = {0, 0}, ffn_E[2] = {0, 0}, ffn_slot0[2] = {0, 0};
    const char* ffn_name[2] = {"feed_forward", nullptr};
    const char* ffn_norm_name[2] = {"ffn_norm", nullptr};
    int S = 1;                               // FFN weight slots per layer (sum over sub-blocks of max(E, 1))
    bf16 *Wg_time = nullptr, *Wg_space = nullptr;   // [L][E][cd] / [L][E][D]
    bf16 *oE = nullptr, *wtok = nullptr;     // expert outputs [E][M][D], token weights [M][E]
    // token-routed experts as a grouped GEMM (option moe_grouped, default on): gathered rows of u / hidden / expert output, per
    // expert a 256-row padded segment; rt = [cnt 8 | cntp 8 | off 9 | cursor 8] ints, pos = slot of each token in its two segments
    bf16 *u_perm = nullptr, *h_perm = nullptr, *o_perm = nullptr;
    int *rt = nullptr, *rpos = nullptr;
    int moe_grouped = 1;
    std::vector<GemmPlan> p_w13g, p_w2g;
    float *temb = nullptr, *tlogits = nullptr;      // [B][cd], [B][L*E]
    int* tsel = nullptr;                     // [L][2] experts chosen by the time gate (device; ascending index)
    float* tw = nullptr;                     // [L][2] their bf16-rounded softmax weights (device)
    std::vector<GemmPlan> p_w13t, p_w2t;     // [L][2] time-gated layers: W = stack of the layer's experts, chosen on the device
    int FD = 1;                              // final-layer adaLN chunks: 1 (scale) or 2 (shift, scale)
    bf16* Yemb = nullptr;                    // [num_classes + 1, cd] label embedding table
    int device = 0, num_sms = 148;
    char err[512];
    int64_t launches = 0;
    int64_t n_params = 0;
    bool finalized = false;
    int attn_ref = 0;
    int tap_layer = -1;                      // debug: >= 0 copies the residual stream after that block into tap_buf (ndit_debug_read_residual)
    bf16* tap_buf = nullptr;
    size_t tap_rows = 0;
    int vt_epi = 1;                          // 1: the q|k|v GEMM epilogue writes V^T itself (no transpose_v launch); 0: separate kernel
    int attn_tp = 0;                         // attention kernel generation: 0 default (ATTN_DEFAULT_GEN), 1 one thread per row + P through
                                             // shared memory (attention_tcgen05.cu), 3 half rows + P in tensor memory (attention_hr_tcgen05.cu)
    int profile = 0;
    int pdl = getenv("NDIT_PDL") ? atoi(getenv("NDIT_PDL")) : 0;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_class;       // class of event pair i (events 2i, 2i+1)
    size_t ev_used = 0;
    std::set<std::string> seen;
    std::vector<void*> allocs;               // weights (live as long as the handle)
    std::vector<size_t> alloc_bytes;         // payload bytes of each weight buffer (ndit_save_packed / ndit_load_packed)
    std::vector<void*> ws_allocs;            // workspace (re-created by ndit_reserve)

    // weights
    bf16 *Wx, *bx, *Wt0, *bt0, *Wt2, *bt2, *capln_w, *capln_b, *Wcap, *bcap, *Wada, *bada, *Wout, *bout, *pad_token, *eol_token;
    bf16 *Wqkv, *Wo, *W13, *W2, *Wkvy;                       // [L][...]
    bf16 *qn_w, *qn_b, *kn_w, *kn_b, *kyn_w, *kyn_b;         // [L][...]
    bf16 *an1, *an2, *fn1, *fn2, *yn;                        // [L][...]
    bf16* gate_raw;                                          // [L][H]
    float* gate_tanh;                                        // [L][H]
    // workspace
    int Mmax, Tmax, Bmax, Tpad_max;
    bf16 *X, *u, *qkv, *attn, *o, *hbuf, *vt;
    bf16 *yhat, *kvy, *vyt;
    uint8_t* ymask;
    float *pool, *capemb, *tf, *h1, *sc, *trow;   // trow: per-row timesteps of a plain forward
    int* kvlen = nullptr;          // per-row valid token counts of a variable-resolution (list) forward
    float2* rope_rows = nullptr;   // per-row rope table of a list forward [batch * Lmax][hd/2] (allocated on first use)
    size_t rope_rows_cap = 0;
    bf16* tok;                               // [M, O] final-layer output tokens
    bf16* mod;
    bf16 *vel, *ystate, *ymid;
    bf16* kbuf[3];                           // rk4: k2, k3, k4 (k1 = vel)
    bf16 *stage_z, *stage_cap;
    uint8_t* stage_mask;
    RopeSlot rope[2];
    int rope_next = 0;
    // caption state
    int cap_batch = 0, cap_T = 0;
    int cap_rows = 0;                        // caption rows held in yhat / kvy / vyt (= cap_batch, or n_cond + 1 in region mode)
    int cap_rows_req = 0, cap_rows_max = 0;  // caption-row capacity of the workspace: max(max_batch, what ndit_set_caption_regions asked for)
    // region-masked captions of the compositional model (ndit_set_caption_regions): n_cond region captions + the unconditional one
    int region_cond = 0, region_hs = 1, region_ws = 1;
    // plans
    int plan_M = 0, plan_B = 0, plan_N = 0, plan_T = 0;
    std::vector<GemmPlan> p_qkv, p_wo, p_w13, p_w2;
    GemmPlan p_final;                        // final_layer.linear: [M, D] x [O, D]^T + bias
    std::vector<AttnPlan> p_attn;
    bool attn_plans_valid = false;
    bool vt_ones_valid = false;
    // CUDA graphs of whole fixed-grid solves (ndit_sample): key = everything the captured launch sequence depends on
    struct SolveGraph {
        std::vector<float> grid;
        int batch = 0, height = 0, width = 0, method = 0, cap_T = 0, cap_rows = 0, region_cond = 0, region_hs = 0, region_ws = 0, with_traj = 0, attn_ref = 0, attn_tp = 0, pdl = 0, vt_epi = 0, moe_grouped = 0;
        ndit_step_params sp;
        cudaGraphExec_t exec = nullptr;
        bool capture_failed = false;         // this solve could not be captured: run it directly from now on
        int64_t launches = 0;
        uint64_t last_use = 0;
        RopeSlot rope_after[2];              // what a replay leaves in the two RoPE table slots
        int rope_next_after = 0;
    };
    std::vector<SolveGraph> graphs;
    uint64_t graph_clock = 0;
    int use_graph = 1;                       // option "graph" / NDIT_GRAPH: replay a captured graph for repeated solves
    cudaStream_t capture_stream = nullptr;   // capture happens here when the caller's stream is the legacy default stream (not capturable)
    int64_t graph_replays = 0;               // solves that ran as one graph launch (ndit_graph_replay_count)
    bf16* traj_buf = nullptr;                // internal trajectory buffer of graph-captured solves [traj_cap][count]
    size_t traj_cap_elems = 0;

    int fail(int code, const char* fmt, ...) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, sizeof(err), fmt, ap);
        va_end(ap);
        return code;
    }
};

static thread_local char g_create_err[512] = "";

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return h->fail(NDIT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define CKL(call)                                                                                             \
    do {                                                                                                      \
        CK(call);                                                                                             \
        h->launches++;                                                                                        \
    } while (0)

// kernel classes for the optional per-launch CUDA-event profile (ndit_profile_read)
enum { KC_GEMM_QKV = 0, KC_GEMM_WO, KC_GEMM_W13, KC_GEMM_W2, KC_ATTN, KC_ROWWISE, KC_COND, KC_COUNT };

static int prof_begin(ndit_engine* h, int cls, cudaStream_t s) {
    if (!h->profile) return 0;
    if (h->ev_used + 2 > h->ev_pool.size()) {
        for (int i = 0; i < 2; ++i) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            h->ev_pool.push_back(e);
        }
    }
    h->ev_class.push_back(cls);
    CK(cudaEventRecord(h->ev_pool[h->ev_used], s));
    return 0;
}
static int prof_end(ndit_engine* h, cudaStream_t s) {
    if (!h->profile) return 0;
    CK(cudaEventRecord(h->ev_pool[h->ev_used + 1], s));
    h->ev_used += 2;
    return 0;
}
#define PROF(cls, call)                                  \
    do {                                                 \
        if (int pe_ = prof_begin(h, cls, s)) return pe_; \
        CKL(call);                                       \
        if (int pe_ = prof_end(h, s)) return pe_;        \
    } while (0)

template <typename T>
static int dev_alloc(ndit_engine* h, T** p, size_t count, bool workspace = false) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
    if (e != cudaSuccess) return h->fail(NDIT_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
    e = cudaMemset(q, 0, count * sizeof(T) + 256);
    if (e != cudaSuccess) return h->fail(NDIT_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
    (workspace ? h->ws_allocs : h->allocs).push_back(q);
    if (!workspace) h->alloc_bytes.push_back(count * sizeof(T));
    *p = static_cast<T*>(q);
    return 0;
}
#define ALLOC(ptr, count)                                \
    do {                                                 \
        int r_ = dev_alloc(h, &(h->ptr), (size_t)(count)); \
        if (r_) return r_;                               \
    } while (0)
#define WALLOC(ptr, count)                                     \
    do {                                                       \
        int r_ = dev_alloc(h, &(h->ptr), (size_t)(count), true); \
        if (r_) return r_;                                     \
    } while (0)

extern "C" int ndit_abi_version(void) { return NDIT_ABI_VERSION; }

extern "C" const char* ndit_last_error(ndit_handle h) { return h ? h->err : g_create_err; }

// Workspace sized by cfg.max_batch / max_tokens / max_cap_len.  Separate from the weights so that ndit_reserve can grow it.
static int alloc_workspace(ndit_engine* h) {
    const ndit_config& c = h->cfg;
    const size_t D = h->D, L = h->L, F = h->F, C = h->C, cd = h->cd, KV = (size_t)h->Hkv * h->hd, NCH = h->NCH, S = h->S;
    h->Bmax = c.max_batch; h->Tmax = h->cls ? 0 : c.max_cap_len; h->Tpad_max = (h->Tmax + 7) / 8 * 8;
    h->Mmax = c.max_batch * c.max_tokens;
    const size_t M = h->Mmax, B = h->Bmax, T = h->Tmax;
    h->cap_rows_max = h->cap_rows_req > h->Bmax ? h->cap_rows_req : h->Bmax;
    const size_t Bc = h->cap_rows_max;       // caption rows (region mode may hold more captions than batch rows)
    WALLOC(X, M * D); WALLOC(u, M * D); WALLOC(qkv, M * h->Wq); WALLOC(attn, M * D); WALLOC(o, M * D); WALLOC(hbuf, M * F);
    WALLOC(vt, B * h->Hkv * h->vrows * ((size_t)c.max_tokens + 8));
    WALLOC(yhat, L * Bc * T * C); WALLOC(kvy, L * Bc * T * 2 * KV); WALLOC(vyt, L * Bc * h->Hkv * h->vrows * h->Tpad_max);
    WALLOC(ymask, Bc * T); WALLOC(pool, B * C); WALLOC(capemb, B * cd); WALLOC(tf, B * 256); WALLOC(trow, B + 8); WALLOC(kvlen, B + 8); WALLOC(h1, B * cd); WALLOC(sc, B * cd);
    WALLOC(mod, B * (L * NCH * D + h->FD * D)); WALLOC(tok, M * h->O);
    if (S > 1) {
        int emax = c.moe_space_experts > 2 ? c.moe_space_experts : 2;
        WALLOC(oE, (size_t)emax * M * D); WALLOC(wtok, M * 8); WALLOC(temb, B * cd); WALLOC(tlogits, B * L * 8);
        WALLOC(tsel, L * 2); WALLOC(tw, L * 2);
        if (c.moe_space_experts > 0) {
            const size_t R = 2 * M + (size_t)c.moe_space_experts * 256;
            WALLOC(u_perm, R * D); WALLOC(h_perm, R * F); WALLOC(o_perm, R * D); WALLOC(rt, 64); WALLOC(rpos, 2 * M);
        }
    }
    const size_t lat = B * c.in_channels * (size_t)c.max_tokens * 4;
    WALLOC(vel, lat); WALLOC(ystate, lat); WALLOC(ymid, lat); WALLOC(kbuf[0], lat); WALLOC(kbuf[1], lat); WALLOC(kbuf[2], lat); WALLOC(stage_z, lat); WALLOC(stage_cap, Bc * T * C); WALLOC(stage_mask, Bc * T);
    for (int i = 0; i < 2; ++i) {
        int r = dev_alloc(h, &h->rope[i].tab, (size_t)c.max_tokens * (h->hd / 2), true);
        h->rope[i].Hp = 0;
        if (r) return r;
    }
    h->plan_M = h->plan_B = h->plan_N = h->plan_T = 0;
    h->attn_plans_valid = false;
    h->vt_ones_valid = false;
    h->cap_batch = h->cap_T = h->cap_rows = 0;
    h->region_cond = 0; h->region_hs = h->region_ws = 1;
    for (auto& g : h->graphs) if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    h->graphs.clear();
    h->traj_buf = nullptr; h->traj_cap_elems = 0;
    return 0;
}

static void free_workspace(ndit_engine* h) {
    for (void* p : h->ws_allocs) cudaFree(p);
    h->ws_allocs.clear();
}

static int create_impl(ndit_engine* h) {
    const ndit_config& c = h->cfg;
    if (c.dim <= 0 || c.n_heads <= 0 || c.dim % c.n_heads != 0) return h->fail(NDIT_ERR_INVALID, "bad dim/n_heads");
    h->D = c.dim; h->L = c.n_layers; h->H = c.n_heads; h->Hkv = c.n_kv_heads > 0 ? c.n_kv_heads : c.n_heads;
    if (getenv("NDIT_ATTN_GEN")) h->attn_tp = atoi(getenv("NDIT_ATTN_GEN"));
    if (getenv("NDIT_VT_EPI")) h->vt_epi = atoi(getenv("NDIT_VT_EPI"));
    if (getenv("NDIT_MOE_GROUPED")) h->moe_grouped = atoi(getenv("NDIT_MOE_GROUPED"));
    if (getenv("NDIT_GRAPH")) h->use_graph = atoi(getenv("NDIT_GRAPH"));
    h->cls = c.num_classes > 0;
    h->flag = c.flag_dit != 0;
    if (h->cls && h->flag) return h->fail(NDIT_ERR_INVALID, "num_classes > 0 and flag_dit are mutually exclusive");
    h->FD = (h->cls || h->flag) ? 2 : 1;
    h->NCH = h->flag ? 6 : 4;
    if (c.moe_time_experts < 0 || c.moe_space_experts < 0 || c.moe_time_experts > 8 || c.moe_space_experts > 8 ||
        c.moe_time_experts == 1 || c.moe_space_experts == 1)
        return h->fail(NDIT_ERR_INVALID, "MoE expert counts must be 0 or 2..8");
    if ((c.moe_time_experts || c.moe_space_experts) && !h->cls)
        return h->fail(NDIT_ERR_INVALID, "the MoE FFN belongs to the class-conditional model (num_classes > 0)");
    if (c.moe_time_experts && c.moe_space_experts) {          // models2.py: time MoE then space MoE, 6-chunk adaLN
        h->NF = 2; h->NCH = 6;
        h->ffn_kind[0] = 1; h->ffn_E[0] = c.moe_time_experts; h->ffn_name[0] = "feed_forward_time"; h->ffn_norm_name[0] = "ffn_norm_time";
        h->ffn_kind[1] = 2; h->ffn_E[1] = c.moe_space_experts; h->ffn_name[1] = "feed_forward_space"; h->ffn_norm_name[1] = "ffn_norm_space";
    } else if (c.moe_time_experts) {                          // models.py
        h->ffn_kind[0] = 1; h->ffn_E[0] = c.moe_time_experts;
    } else if (c.moe_space_experts) {                         // models1.py
        h->ffn_kind[0] = 2; h->ffn_E[0] = c.moe_space_experts;
    }
    h->S = 0;
    for (int f = 0; f < h->NF; ++f) { h->ffn_slot0[f] = h->S; h->S += h->ffn_E[f] > 0 ? h->ffn_E[f] : 1; }
    h->hd = c.dim / c.n_heads; h->C = h->cls ? 0 : c.cap_feat_dim; h->cd = c.dim < 1024 ? c.dim : 1024;
    if (h->hd != 72 && h->hd != 48 && h->hd != 96) return h->fail(NDIT_ERR_INVALID, "head_dim must be 72, 48 or 96 (got %d)", h->hd);
    h->vrows = attn_vrows(h->hd);
    if (c.patch_size != 2 || c.in_channels < 2 || c.in_channels > 16 || (c.in_channels & 1))
        return h->fail(NDIT_ERR_INVALID, "patch_size 2 and an even in_channels in 2..16 only (got %d, %d)", c.patch_size, c.in_channels);
    if (h->H % h->Hkv != 0) return h->fail(NDIT_ERR_INVALID, "n_heads %% n_kv_heads != 0");
    if (c.max_batch < 2 || c.max_batch > 4 || (c.max_batch & 1)) return h->fail(NDIT_ERR_INVALID, "max_batch must be 2 or 4");
    if (h->C % 8 != 0 || h->D % 64 != 0) return h->fail(NDIT_ERR_INVALID, "dims must be multiples of 8/64");
    int hidden = static_cast<int>(2 * (4 * c.dim) / 3);                 // model.py:441-503 FeedForward
    h->F = c.ffn_dim > 0 ? c.ffn_dim : c.multiple_of * ((hidden + c.multiple_of - 1) / c.multiple_of);
    if (h->F % 128 != 0) return h->fail(NDIT_ERR_INVALID, "ffn dim must be a multiple of 128");
    const int out_ch = c.learn_sigma ? 2 * c.in_channels : c.in_channels;
    h->O = c.patch_size * c.patch_size * out_ch;
    h->Wq = (h->H + 2 * h->Hkv) * h->hd;
    CK(cudaGetDevice(&h->device));
    CK(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device));
    int cc_major = 0;
    CK(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, h->device));
    if (cc_major != 10) return h->fail(NDIT_ERR_INVALID, "needs an sm_100 device (compute capability %d.x found)", cc_major);

    const size_t D = h->D, L = h->L, F = h->F, C = h->C, cd = h->cd, KV = (size_t)h->Hkv * h->hd;
    ALLOC(Wx, D * 4 * c.in_channels); ALLOC(bx, D); ALLOC(Wt0, cd * 256); ALLOC(bt0, cd); ALLOC(Wt2, cd * cd); ALLOC(bt2, cd);
    ALLOC(capln_w, C); ALLOC(capln_b, C); ALLOC(Wcap, cd * C); ALLOC(bcap, cd);
    const size_t NCH = h->NCH;
    ALLOC(Wada, (L * NCH * D + h->FD * D) * cd); ALLOC(bada, L * NCH * D + h->FD * D);
    if (h->cls) ALLOC(Yemb, ((size_t)c.num_classes + 1) * cd);
    ALLOC(Wout, (size_t)h->O * D); ALLOC(bout, h->O); ALLOC(pad_token, D); ALLOC(eol_token, D);
    const size_t S = h->S, NF = h->NF;
    ALLOC(Wqkv, L * h->Wq * D); ALLOC(Wo, L * D * D); ALLOC(W13, L * S * 2 * F * D); ALLOC(W2, L * S * D * F);
    if (c.moe_time_experts) ALLOC(Wg_time, L * c.moe_time_experts * cd);
    if (c.moe_space_experts) ALLOC(Wg_space, L * c.moe_space_experts * D);
    ALLOC(Wkvy, L * 2 * KV * C);
    ALLOC(qn_w, L * D); ALLOC(qn_b, L * D); ALLOC(kn_w, L * KV); ALLOC(kn_b, L * KV); ALLOC(kyn_w, L * KV); ALLOC(kyn_b, L * KV);
    ALLOC(an1, L * D); ALLOC(an2, L * D); ALLOC(fn1, L * D); ALLOC(fn2, NF * L * D); ALLOC(yn, L * C);
    ALLOC(gate_raw, L * h->H); ALLOC(gate_tanh, L * h->H);

    if (int r = alloc_workspace(h)) return r;
    if (h->cls) {
        // weight-free pre-norms (PFRMSNorm, models.py:76-117) = RMSNorm with unit weight
        std::vector<uint16_t> ones(L * D, 0x3F80);
        CK(cudaMemcpy(h->an1, ones.data(), ones.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->fn1, ones.data(), ones.size() * 2, cudaMemcpyHostToDevice));
    }
    return 0;
}

extern "C" int ndit_create(const ndit_config* cfg, ndit_handle* out) {
    if (!cfg || !out) {
        snprintf(g_create_err, sizeof(g_create_err), "ndit_create: null argument");
        return NDIT_ERR_INVALID;
    }
    ndit_engine* h = new ndit_engine();
    h->cfg = *cfg;
    h->err[0] = 0;
    int r = create_impl(h);
    if (r) {
        snprintf(g_create_err, sizeof(g_create_err), "%s", h->err);
        for (void* p : h->allocs) cudaFree(p);
        free_workspace(h);
        delete h;
        *out = nullptr;
        return r;
    }
    *out = h;
    return NDIT_OK;
}

extern "C" int ndit_destroy(ndit_handle h) {
    if (!h) return NDIT_OK;
    cudaDeviceSynchronize();
    for (auto& g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    for (void* p : h->allocs) cudaFree(p);
    if (h->rope_rows) cudaFree(h->rope_rows);
    if (h->tap_buf) cudaFree(h->tap_buf);
    if (h->capture_stream) cudaStreamDestroy(h->capture_stream);
    free_workspace(h);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    delete h;
    return NDIT_OK;
}

extern "C" int ndit_reserve(ndit_handle h, int32_t max_tokens, int32_t max_cap_len, int32_t max_batch) {
    if (!h) return NDIT_ERR_INVALID;
    ndit_config& c = h->cfg;
    const int nt = max_tokens > c.max_tokens ? max_tokens : c.max_tokens;
    const int nc = max_cap_len > c.max_cap_len ? max_cap_len : c.max_cap_len;
    const int nb = max_batch > c.max_batch ? max_batch : c.max_batch;
    if (nt == c.max_tokens && nc == c.max_cap_len && nb == c.max_batch) return NDIT_OK;
    if (nb > 4 || (nb & 1)) return h->fail(NDIT_ERR_INVALID, "max_batch must be 2 or 4");
    CK(cudaDeviceSynchronize());
    const ndit_config old = c;
    free_workspace(h);
    c.max_tokens = nt; c.max_cap_len = nc; c.max_batch = nb;
    if (int r = alloc_workspace(h)) {          // out of memory: go back to the old size so the handle stays usable
        free_workspace(h);
        c = old;
        char msg[512];
        snprintf(msg, sizeof(msg), "%s", h->err);
        if (alloc_workspace(h)) return h->fail(NDIT_ERR_NOMEM, "ndit_reserve: could not restore the workspace after: %s", msg);
        return h->fail(r, "ndit_reserve(%d tokens, %d caption tokens, batch %d): %s", nt, nc, nb, msg);
    }
    return NDIT_OK;
}

extern "C" int64_t ndit_parameter_count(ndit_handle h) { return h ? h->n_params : 0; }
extern "C" int64_t ndit_launch_count(ndit_handle h) { return h ? h->launches : 0; }
extern "C" int64_t ndit_graph_replay_count(ndit_handle h) { return h ? h->graph_replays : 0; }

extern "C" int ndit_set_option(ndit_handle h, const char* name, int32_t value) {
    if (!h || !name) return NDIT_ERR_INVALID;
    if (!strcmp(name, "attn_ref")) { h->attn_ref = value; return NDIT_OK; }
    if (!strcmp(name, "graph")) { h->use_graph = value ? 1 : 0; return NDIT_OK; }
    if (!strcmp(name, "pdl")) { h->pdl = value ? 1 : 0; return NDIT_OK; }
    if (!strcmp(name, "vt_epi")) { h->vt_epi = value; return NDIT_OK; }
    if (!strcmp(name, "moe_grouped")) { h->moe_grouped = value; return NDIT_OK; }
    if (!strcmp(name, "tap_layer")) { h->tap_layer = value; return NDIT_OK; }
    if (!strcmp(name, "attn_gen") || !strcmp(name, "attn_tp")) { h->attn_tp = value; h->attn_plans_valid = false; return NDIT_OK; }
    if (!strcmp(name, "profile")) {
        h->profile = value;
        h->ev_used = 0;
        h->ev_class.clear();
        return NDIT_OK;
    }
    return h->fail(NDIT_ERR_INVALID, "unknown option %s", name);
}

// ------------------------------------------------------------------------------------ weights

static int place(ndit_engine* h, bf16* dst, size_t dst_ld, const void* src, int dtype, size_t rows, size_t cols,
                 size_t blk, size_t blk_stride, size_t row0, cudaStream_t s, bool fresh = true) {
    const size_t total = rows * cols;
    if (fresh) h->n_params += (int64_t)total;     // NextDiT.parameter_count: every state-dict tensor once
    const int grid = (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256);
    place_rows_kernel<<<grid, 256, 0, s>>>(dst, dst_ld, src, dtype == NDIT_F32, rows, cols, blk ? blk : rows,
                                           blk_stride, row0);
    CKL(cudaGetLastError());
    return 0;
}

extern "C" int ndit_set_weight(ndit_handle h, const char* key, const void* src, const int64_t* shape, int32_t ndim,
                               int32_t dtype, void* stream) {
    if (!h || !key || !src || !shape) return NDIT_ERR_INVALID;
    if (dtype != NDIT_BF16 && dtype != NDIT_F32) return h->fail(NDIT_ERR_INVALID, "%s: dtype must be bf16 or f32", key);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t D = h->D, F = h->F, C = h->C, cd = h->cd, KV = (size_t)h->Hkv * h->hd, Wq = h->Wq;
    const size_t r = ndim >= 1 ? (size_t)shape[0] : 1, c = ndim >= 2 ? (size_t)shape[1] : 1;
    auto want = [&](size_t er, size_t ec) -> int {
        const bool ok = (ec == 0) ? (ndim == 1 && r == er) : (ndim == 2 && r == er && c == ec);
        if (!ok) return h->fail(NDIT_ERR_INVALID, "%s: shape mismatch (got [%zu,%zu] ndim %d, want [%zu,%zu])", key, r, c, ndim, er, ec);
        return 0;
    };
#define VEC(name, dst, n)                                         \
    if (!strcmp(key, name)) {                                     \
        if (int e = want(n, 0)) return e;                         \
        const bool fresh_ = h->seen.insert(key).second;           \
        return place(h, dst, 1, src, dtype, n, 1, 0, 0, 0, s, fresh_); \
    }
#define MAT(name, dst, rr, cc)                                    \
    if (!strcmp(key, name)) {                                     \
        if (int e = want(rr, cc)) return e;                       \
        const bool fresh_ = h->seen.insert(key).second;           \
        return place(h, dst, cc, src, dtype, rr, cc, 0, 0, 0, s, fresh_); \
    }
    if (!h->cls) { VEC("pad_token", h->pad_token, D) }
    if (h->flag) { VEC("eol_token", h->eol_token, D) }
    MAT("x_embedder.weight", h->Wx, D, (size_t)4 * h->cfg.in_channels) VEC("x_embedder.bias", h->bx, D)
    MAT("t_embedder.mlp.0.weight", h->Wt0, cd, 256) VEC("t_embedder.mlp.0.bias", h->bt0, cd)
    MAT("t_embedder.mlp.2.weight", h->Wt2, cd, cd) VEC("t_embedder.mlp.2.bias", h->bt2, cd)
    if (!h->cls) {
        VEC("cap_embedder.0.weight", h->capln_w, C) VEC("cap_embedder.0.bias", h->capln_b, C)
        MAT("cap_embedder.1.weight", h->Wcap, cd, C) VEC("cap_embedder.1.bias", h->bcap, cd)
    } else {
        MAT("y_embedder.embedding_table.weight", h->Yemb, (size_t)h->cfg.num_classes + 1, cd)
    }
    MAT("final_layer.linear.weight", h->Wout, (size_t)h->O, D) VEC("final_layer.linear.bias", h->bout, (size_t)h->O)
    const size_t Lz = h->L, NCH = h->NCH;
    // Next-DiT T2I: [scale]; class-conditional and Flag-DiT: [shift | scale] (models.py:829-833, lumina_t2i model.py:655-656)
    MAT("final_layer.adaLN_modulation.1.weight", h->Wada + Lz * NCH * D * cd, (size_t)h->FD * D, cd)
    VEC("final_layer.adaLN_modulation.1.bias", h->bada + Lz * NCH * D, (size_t)h->FD * D)
#undef VEC
#undef MAT
    int li = -1, pos = 0;
    if (sscanf(key, "layers.%d.%n", &li, &pos) >= 1 && pos > 0 && li >= 0 && li < h->L) {
        const char* sub = key + pos;
        const size_t l = li;
        if (h->cfg.no_qk_norm && (!strncmp(sub, "attention.q_norm.", 17) || !strncmp(sub, "attention.k_norm.", 17) || !strncmp(sub, "attention.ky_norm.", 18)))
            return h->fail(NDIT_ERR_INVALID, "unexpected state-dict key for a qk_norm=False model: %s", key);
        // FFN sub-blocks: "<name>.w{1,2,3}.weight" (dense) or "<name>.experts.<j>.w{1,2,3}.weight" + "<name>.gate.weight" (MoE)
        for (int f = 0; f < h->NF; ++f) {
            const size_t nl = strlen(h->ffn_name[f]);
            if (strncmp(sub, h->ffn_name[f], nl) != 0 || sub[nl] != '.') continue;
            const char* rest = sub + nl + 1;
            const int E = h->ffn_E[f];
            size_t slot = h->ffn_slot0[f];
            if (E > 0) {
                if (!strcmp(rest, "gate.weight")) {
                    const bool time = h->ffn_kind[f] == 1;
                    const size_t gc = time ? cd : D;
                    if (int er = want((size_t)E, gc)) return er;
                    const bool fresh = h->seen.insert(key).second;
                    return place(h, (time ? h->Wg_time : h->Wg_space) + l * E * gc, gc, src, dtype, (size_t)E, gc, 0, 0, 0, s, fresh);
                }
                int ej = -1, p2 = 0;
                if (sscanf(rest, "experts.%d.%n", &ej, &p2) < 1 || p2 == 0 || ej < 0 || ej >= E)
                    return h->fail(NDIT_ERR_INVALID, "unexpected state-dict key: %s", key);
                rest += p2;
                slot += ej;
            }
            const size_t S_ = h->S;
            bf16* w13 = h->W13 + (l * S_ + slot) * 2 * F * D;
            bf16* w2 = h->W2 + (l * S_ + slot) * D * F;
            const bool fresh = h->seen.count(key) == 0;
            if (!strcmp(rest, "w1.weight")) { if (int er = want(F, D)) return er; h->seen.insert(key); return place(h, w13, D, src, dtype, F, D, 128, 256, 0, s, fresh); }
            if (!strcmp(rest, "w3.weight")) { if (int er = want(F, D)) return er; h->seen.insert(key); return place(h, w13, D, src, dtype, F, D, 128, 256, 128, s, fresh); }
            if (!strcmp(rest, "w2.weight")) { if (int er = want(D, F)) return er; h->seen.insert(key); return place(h, w2, F, src, dtype, D, F, 0, 0, 0, s, fresh); }
            return h->fail(NDIT_ERR_INVALID, "unexpected state-dict key: %s", key);
        }
        if (h->cls) {
            for (int f = 0; f < h->NF; ++f) {
                const std::string nk = std::string(h->ffn_norm_name[f]) + ".weight";
                if (nk == sub) {
                    if (int er = want(D, 0)) return er;
                    const bool fresh = h->seen.insert(key).second;
                    return place(h, h->fn2 + ((size_t)f * h->L + l) * D, 1, src, dtype, D, 1, 0, 0, 0, s, fresh);
                }
            }
        }
        struct Ent { const char* name; bf16* dst; size_t dst_ld, rows, cols, blk, blk_stride, row0; };
        const Ent ents[] = {
            {"attention.wq.weight", h->Wqkv + l * Wq * D, D, D, D, 0, 0, 0},
            {"attention.wk.weight", h->Wqkv + l * Wq * D, D, KV, D, 0, 0, D},
            {"attention.wv.weight", h->Wqkv + l * Wq * D, D, KV, D, 0, 0, D + KV},
            {"attention.wo.weight", h->Wo + l * D * D, D, D, D, 0, 0, 0},
            {"attention.wk_y.weight", h->Wkvy + l * 2 * KV * C, C, KV, C, 0, 0, 0},
            {"attention.wv_y.weight", h->Wkvy + l * 2 * KV * C, C, KV, C, 0, 0, KV},
            // (feed_forward.w1|w3|w2 are handled above: w1|w3 interleaved per 256-row block for the SwiGLU epilogue)
            {"adaLN_modulation.1.weight", h->Wada + l * NCH * D * cd, cd, NCH * D, cd, 0, 0, 0},
        };
        for (const Ent& e : ents) {
            if (h->cls && (!strcmp(e.name, "attention.wk_y.weight") || !strcmp(e.name, "attention.wv_y.weight"))) continue;
            if (!strcmp(sub, e.name)) {
                if (int er = want(e.rows, e.cols)) return er;
                const bool fresh = h->seen.insert(key).second;
                return place(h, e.dst, e.dst_ld, src, dtype, e.rows, e.cols, e.blk, e.blk_stride, e.row0, s, fresh);
            }
        }
        struct VEnt { const char* name; bf16* dst; size_t n; };
        const VEnt vents[] = {
            {"attention.gate", h->gate_raw + l * h->H, (size_t)h->H},
            {"attention.q_norm.weight", h->qn_w + l * D, D}, {"attention.q_norm.bias", h->qn_b + l * D, D},
            {"attention.k_norm.weight", h->kn_w + l * KV, KV}, {"attention.k_norm.bias", h->kn_b + l * KV, KV},
            {"attention.ky_norm.weight", h->kyn_w + l * KV, KV}, {"attention.ky_norm.bias", h->kyn_b + l * KV, KV},
            {"attention_norm1.weight", h->an1 + l * D, D}, {"attention_norm2.weight", h->an2 + l * D, D},
            {"ffn_norm1.weight", h->fn1 + l * D, D}, {"ffn_norm2.weight", h->fn2 + l * D, D},
            {"attention_y_norm.weight", h->yn + l * C, C},
            {"adaLN_modulation.1.bias", h->bada + l * 4 * D, 4 * D},
        };
        // class-conditional block (TransformerBlockSandwichNorm2, models.py:692-796): post-norms are called
        // attention_norm / ffn_norm, the pre-norms carry no weight, there is no caption branch
        const VEnt cls_vents[] = {
            {"attention.q_norm.weight", h->qn_w + l * D, D}, {"attention.q_norm.bias", h->qn_b + l * D, D},
            {"attention.k_norm.weight", h->kn_w + l * KV, KV}, {"attention.k_norm.bias", h->kn_b + l * KV, KV},
            {"attention_norm.weight", h->an2 + l * D, D},
            {"adaLN_modulation.1.bias", h->bada + l * NCH * D, NCH * D},
        };
        // Flag-DiT block (lumina_t2i model.py:505-622): one weighted RMSNorm in front of each sub-block, no post-norms
        const VEnt flag_vents[] = {
            {"attention.gate", h->gate_raw + l * h->H, (size_t)h->H},
            {"attention.q_norm.weight", h->qn_w + l * D, D}, {"attention.q_norm.bias", h->qn_b + l * D, D},
            {"attention.k_norm.weight", h->kn_w + l * KV, KV}, {"attention.k_norm.bias", h->kn_b + l * KV, KV},
            {"attention.ky_norm.weight", h->kyn_w + l * KV, KV}, {"attention.ky_norm.bias", h->kyn_b + l * KV, KV},
            {"attention_norm.weight", h->an1 + l * D, D}, {"ffn_norm.weight", h->fn1 + l * D, D},
            {"attention_y_norm.weight", h->yn + l * C, C},
            {"adaLN_modulation.1.bias", h->bada + l * 6 * D, 6 * D},
        };
        if (h->flag) {
            for (const VEnt& e : flag_vents) {
                if (!strcmp(sub, e.name)) {
                    if (int er = want(e.n, 0)) return er;
                    const bool fresh = h->seen.insert(key).second;
                    return place(h, e.dst, 1, src, dtype, e.n, 1, 0, 0, 0, s, fresh);
                }
            }
            return h->fail(NDIT_ERR_INVALID, "unexpected state-dict key: %s", key);
        }
        if (h->cls) {
            for (const VEnt& e : cls_vents) {
                if (!strcmp(sub, e.name)) {
                    if (int er = want(e.n, 0)) return er;
                    const bool fresh = h->seen.insert(key).second;
                    return place(h, e.dst, 1, src, dtype, e.n, 1, 0, 0, 0, s, fresh);
                }
            }
            return h->fail(NDIT_ERR_INVALID, "unexpected state-dict key: %s", key);
        }
        for (const VEnt& e : vents) {
            if (!strcmp(sub, e.name)) {
                if (int er = want(e.n, 0)) return er;
                const bool fresh = h->seen.insert(key).second;
                return place(h, e.dst, 1, src, dtype, e.n, 1, 0, 0, 0, s, fresh);
            }
        }
    }
    return h->fail(NDIT_ERR_INVALID, "unexpected state-dict key: %s", key);
}

static void expected_keys(const ndit_engine* h, std::vector<std::string>* out) {
    const char* common[] = {"x_embedder.weight", "x_embedder.bias", "t_embedder.mlp.0.weight", "t_embedder.mlp.0.bias",
                            "t_embedder.mlp.2.weight", "t_embedder.mlp.2.bias", "final_layer.linear.weight", "final_layer.linear.bias",
                            "final_layer.adaLN_modulation.1.weight", "final_layer.adaLN_modulation.1.bias"};
    for (const char* k : common) out->push_back(k);
    const char* t2i_top[] = {"pad_token", "cap_embedder.0.weight", "cap_embedder.0.bias", "cap_embedder.1.weight", "cap_embedder.1.bias"};
    if (h->cls) out->push_back("y_embedder.embedding_table.weight");
    else for (const char* k : t2i_top) out->push_back(k);
    if (h->flag) out->push_back("eol_token");
    const char* per[] = {"attention.wq.weight", "attention.wk.weight", "attention.wv.weight", "attention.wo.weight",
                         "attention.q_norm.weight", "attention.q_norm.bias", "attention.k_norm.weight", "attention.k_norm.bias",
                         "adaLN_modulation.1.weight", "adaLN_modulation.1.bias"};
    const char* per_t2i[] = {"attention.gate", "attention.wk_y.weight", "attention.wv_y.weight", "attention.ky_norm.weight",
                             "attention.ky_norm.bias", "attention_norm1.weight", "attention_norm2.weight", "ffn_norm1.weight",
                             "ffn_norm2.weight", "attention_y_norm.weight"};
    const char* per_cls[] = {"attention_norm.weight"};
    const char* per_flag[] = {"attention.gate", "attention.wk_y.weight", "attention.wv_y.weight", "attention.ky_norm.weight",
                              "attention.ky_norm.bias", "attention_norm.weight", "ffn_norm.weight", "attention_y_norm.weight"};
    for (int l = 0; l < h->L; ++l) {
        const std::string pre = "layers." + std::to_string(l) + ".";
        for (const char* k : per) if (!(h->cfg.no_qk_norm && strstr(k, "_norm."))) out->push_back(pre + k);
        for (int f = 0; f < h->NF; ++f) {
            const std::string fp = pre + h->ffn_name[f] + ".";
            const int E = h->ffn_E[f];
            for (int j = 0; j < (E > 0 ? E : 1); ++j) {
                const std::string ep = E > 0 ? fp + "experts." + std::to_string(j) + "." : fp;
                out->push_back(ep + "w1.weight"); out->push_back(ep + "w2.weight"); out->push_back(ep + "w3.weight");
            }
            if (E > 0) out->push_back(fp + "gate.weight");
            if (h->cls) out->push_back(pre + h->ffn_norm_name[f] + ".weight");
        }
        if (h->cls) for (const char* k : per_cls) out->push_back(pre + k);
        else if (h->flag) { for (const char* k : per_flag) if (!(h->cfg.no_qk_norm && strstr(k, "ky_norm."))) out->push_back(pre + k); }
        else { for (const char* k : per_t2i) if (!(h->cfg.no_qk_norm && strstr(k, "ky_norm."))) out->push_back(pre + k); }
    }
}

extern "C" int ndit_finalize_weights(ndit_handle h, void* stream) {
    if (!h) return NDIT_ERR_INVALID;
    std::vector<std::string> keys;
    expected_keys(h, &keys);
    for (const std::string& k : keys)
        if (!h->seen.count(k)) return h->fail(NDIT_ERR_STATE, "missing state-dict key (strict): %s", k.c_str());
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (h->cls) {
        h->finalized = true;
        return NDIT_OK;
    }
    // tanh(gate) per head, bf16 in / bf16 out (model.py:433)
    const size_t n = (size_t)h->L * h->H;
    std::vector<uint16_t> raw(n);
    std::vector<float> th(n);
    CK(cudaMemcpyAsync(raw.data(), h->gate_raw, n * 2, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (size_t i = 0; i < n; ++i) th[i] = host_bf16_round(tanhf(host_bf16_to_float(raw[i])));
    CK(cudaMemcpyAsync(h->gate_tanh, th.data(), n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    h->finalized = true;
    return NDIT_OK;
}

// ------------------------------------------------------------------------------------ packed weight files
// The engine's own weight layout (fused q|k|v, w1|w3 interleaved per 256 rows, packed adaLN matrices, expert stacks,
// precomputed tanh(gate)) written to / read from one flat file, so a cold start copies ~3.3 GB straight into place instead of
// re-packing the state dict with one kernel per tensor (the role of the reference's checkpoint tooling,
// lumina_next_t2i/entry_point.py:115-156, for this engine).  Layout: header | sizes[n] | buffers back to back.
struct PackedHeader {
    char magic[8];            // "NDITPK02"
    int32_t abi, config_bytes;
    ndit_config cfg;
    int64_t n_params;
    int32_t n_buffers, finalized;
};

extern "C" int ndit_save_packed(ndit_handle h, const char* path) {
    if (!h || !path) return NDIT_ERR_INVALID;
    if (!h->finalized) return h->fail(NDIT_ERR_STATE, "ndit_save_packed: weights not finalized");
    CK(cudaDeviceSynchronize());
    FILE* f = fopen(path, "wb");
    if (!f) return h->fail(NDIT_ERR_INVALID, "ndit_save_packed: cannot open %s for writing", path);
    PackedHeader hd;
    memset(&hd, 0, sizeof(hd));
    memcpy(hd.magic, "NDITPK02", 8);
    hd.abi = NDIT_ABI_VERSION; hd.config_bytes = (int32_t)sizeof(ndit_config); hd.cfg = h->cfg;
    hd.cfg.max_tokens = hd.cfg.max_cap_len = hd.cfg.max_batch = 0;      // workspace limits are not part of the weights
    hd.n_params = h->n_params; hd.n_buffers = (int32_t)h->allocs.size(); hd.finalized = 1;
    bool ok = fwrite(&hd, sizeof(hd), 1, f) == 1;
    std::vector<uint64_t> sizes(h->alloc_bytes.begin(), h->alloc_bytes.end());
    ok = ok && fwrite(sizes.data(), sizeof(uint64_t), sizes.size(), f) == sizes.size();
    const size_t CH = (size_t)64 << 20;
    void* bounce = nullptr;
    if (cudaMallocHost(&bounce, CH) != cudaSuccess) { fclose(f); cudaGetLastError(); return h->fail(NDIT_ERR_NOMEM, "ndit_save_packed: no pinned staging buffer"); }
    for (size_t i = 0; ok && i < h->allocs.size(); ++i) {
        for (size_t off = 0; ok && off < sizes[i]; off += CH) {
            const size_t n = sizes[i] - off < CH ? sizes[i] - off : CH;
            ok = cudaMemcpy(bounce, static_cast<char*>(h->allocs[i]) + off, n, cudaMemcpyDeviceToHost) == cudaSuccess && fwrite(bounce, 1, n, f) == n;
        }
    }
    cudaFreeHost(bounce);
    ok = (fclose(f) == 0) && ok;
    return ok ? NDIT_OK : h->fail(NDIT_ERR_CUDA, "ndit_save_packed: write to %s failed", path);
}

extern "C" int ndit_load_packed(ndit_handle h, const char* path) {
    if (!h || !path) return NDIT_ERR_INVALID;
    FILE* f = fopen(path, "rb");
    if (!f) return h->fail(NDIT_ERR_INVALID, "ndit_load_packed: cannot open %s", path);
    PackedHeader hd;
    std::vector<uint64_t> sizes;
    bool ok = fread(&hd, sizeof(hd), 1, f) == 1 && !memcmp(hd.magic, "NDITPK02", 8) && hd.abi == NDIT_ABI_VERSION &&
              hd.config_bytes == (int32_t)sizeof(ndit_config) && hd.n_buffers == (int32_t)h->allocs.size();
    if (ok) {
        ndit_config a = hd.cfg, b = h->cfg;
        a.max_tokens = a.max_cap_len = a.max_batch = b.max_tokens = b.max_cap_len = b.max_batch = 0;
        ok = !memcmp(&a, &b, sizeof(a));
    }
    if (ok) {
        sizes.resize(hd.n_buffers);
        ok = fread(sizes.data(), sizeof(uint64_t), sizes.size(), f) == sizes.size();
        for (size_t i = 0; ok && i < sizes.size(); ++i) ok = sizes[i] == h->alloc_bytes[i];
    }
    if (!ok) { fclose(f); return h->fail(NDIT_ERR_INVALID, "ndit_load_packed: %s is not a packed weight file of this architecture / ABI", path); }
    const size_t CH = (size_t)64 << 20;
    void* bounce[2] = {nullptr, nullptr};
    cudaStream_t cs = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    if (cudaMallocHost(&bounce[0], CH) != cudaSuccess || cudaMallocHost(&bounce[1], CH) != cudaSuccess || cudaStreamCreate(&cs) != cudaSuccess ||
        cudaEventCreate(&ev[0]) != cudaSuccess || cudaEventCreate(&ev[1]) != cudaSuccess) {
        fclose(f);
        cudaGetLastError();
        if (bounce[0]) cudaFreeHost(bounce[0]);
        if (bounce[1]) cudaFreeHost(bounce[1]);
        return h->fail(NDIT_ERR_NOMEM, "ndit_load_packed: no staging resources");
    }
    int cur = 0;        // disk read of chunk i+1 overlaps the H2D copy of chunk i
    bool used[2] = {false, false};
    for (size_t i = 0; ok && i < h->allocs.size(); ++i) {
        for (size_t off = 0; ok && off < sizes[i]; off += CH) {
            const size_t n = sizes[i] - off < CH ? sizes[i] - off : CH;
            if (used[cur]) ok = cudaEventSynchronize(ev[cur]) == cudaSuccess;
            ok = ok && fread(bounce[cur], 1, n, f) == n &&
                 cudaMemcpyAsync(static_cast<char*>(h->allocs[i]) + off, bounce[cur], n, cudaMemcpyHostToDevice, cs) == cudaSuccess &&
                 cudaEventRecord(ev[cur], cs) == cudaSuccess;
            used[cur] = true;
            cur ^= 1;
        }
    }
    ok = cudaStreamSynchronize(cs) == cudaSuccess && ok;
    cudaFreeHost(bounce[0]); cudaFreeHost(bounce[1]); cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]); cudaStreamDestroy(cs);
    fclose(f);
    if (!ok) return h->fail(NDIT_ERR_CUDA, "ndit_load_packed: reading %s failed", path);
    std::vector<std::string> keys;
    expected_keys(h, &keys);
    for (const std::string& k : keys) h->seen.insert(k);
    h->n_params = hd.n_params;
    h->finalized = true;
    return NDIT_OK;
}

// ------------------------------------------------------------------------------------ caption

// Caption-side work shared by ndit_set_caption and ndit_set_caption_regions: `rows` caption rows [rows, T, C] -> per layer
// attention_y_norm, wk_y | wv_y, ky_norm, V^T with its all-ones row.  pool_cap / pool_mask [pool_rows, pool_T]: what the adaLN
// conditioning pools over (model.py:847-850) - the caption rows themselves, or the global caption of the compositional model.
static int set_caption_impl(ndit_engine* h, const bf16* capb, const uint8_t* mask, int rows, int T, const bf16* pool_cap,
                            const uint8_t* pool_mask_u8, int pool_rows, int pool_T, cudaStream_t s) {
    const size_t C = h->C, KV = (size_t)h->Hkv * h->hd, L = h->L;
    const int M = rows * T;
    const int Tpad = (T + 7) / 8 * 8;
    mask_to_u8_kernel<<<(M + 255) / 256, 256, 0, s>>>(h->ymask, mask, M);
    CKL(cudaGetLastError());
    CKL(cond_prepare(0.f, nullptr, pool_cap ? pool_cap : capb, pool_cap ? pool_mask_u8 : h->ymask, h->capln_w, h->capln_b, h->tf, h->pool,
                     pool_rows, pool_T, (int)C, 1, s));
    CKL(gemv_rows(h->pool, h->Wcap, h->bcap, nullptr, h->capemb, nullptr, pool_rows, h->cd, (int)C, 0, POST_NONE, 0, 0, 0, s));
    CKL(rms_rows_layers(capb, h->yn, h->yhat, M, (int)C, (int)L, h->cfg.norm_eps, s));
    const size_t ys = (size_t)M * C, ks = (size_t)M * 2 * KV;
    for (size_t l = 0; l < L; ++l) {
        GemmPlan p;
        if (make_gemm_plan(&p, h->yhat + l * ys, (int)C, h->Wkvy + l * 2 * KV * C, h->kvy + l * ks, (int)(2 * KV), M,
                           (int)(2 * KV), (int)C, EPI_STORE, h->num_sms))
            return h->fail(NDIT_ERR_CUDA, "%s", tmap_last_error());
        CKL(gemm_bf16_tn(p, s));
    }
    if (!h->cfg.no_qk_norm) CKL(ln_rows(h->kvy, (int)(2 * KV), ks, h->kyn_w, h->kyn_b, KV, M, (int)KV, (int)L, s));
    const size_t vs = (size_t)rows * h->Hkv * h->vrows * Tpad;
    CK(cudaMemsetAsync(h->vyt, 0, L * vs * sizeof(bf16), s));
    CKL(transpose_v(h->kvy, (int)(2 * KV), (int)KV, ks, h->vyt, Tpad, vs, rows, T, h->Hkv, h->hd, h->vrows, (int)L, s));
    CKL(fill_ones_row(h->vyt, Tpad, vs, rows * h->Hkv, Tpad, h->hd, h->vrows, (int)L, s));
    return NDIT_OK;
}

extern "C" int ndit_set_caption(ndit_handle h, const void* cap, const uint8_t* mask, int32_t batch, int32_t T, void* stream) {
    if (!h || !cap || !mask) return NDIT_ERR_INVALID;
    if (h->cls) return h->fail(NDIT_ERR_STATE, "ndit_set_caption: this engine is class-conditional (use ndit_set_labels)");
    if (!h->finalized) return h->fail(NDIT_ERR_STATE, "weights not finalized");
    if (batch < 1 || batch > h->Bmax || T < 1 || T > h->Tmax) return h->fail(NDIT_ERR_INVALID, "caption batch/T out of range (%d,%d)", batch, T);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (int e = set_caption_impl(h, static_cast<const bf16*>(cap), mask, batch, T, nullptr, nullptr, batch, T, s)) return e;
    if (batch != h->cap_batch || batch != h->cap_rows || T != h->cap_T || h->region_cond != 0) h->attn_plans_valid = false;
    h->cap_batch = batch;
    h->cap_rows = batch;
    h->cap_T = T;
    h->region_cond = 0; h->region_hs = h->region_ws = 1;
    return NDIT_OK;
}

extern "C" int ndit_set_caption_regions(ndit_handle h, const void* cap, const uint8_t* mask, int32_t n_caps, int32_t T,
                                        const void* global_cap, const uint8_t* global_mask, int32_t global_T, int32_t h_split,
                                        int32_t w_split, void* stream) {
    if (!h || !cap || !mask || !global_cap || !global_mask) return NDIT_ERR_INVALID;
    if (h->cls || h->flag) return h->fail(NDIT_ERR_STATE, "ndit_set_caption_regions: text-conditioned Next-DiT only");
    if (!h->finalized) return h->fail(NDIT_ERR_STATE, "weights not finalized");
    if (h->hd != 72) return h->fail(NDIT_ERR_INVALID, "region-masked cross-attention is built for head_dim 72 (got %d)", h->hd);
    if (n_caps < 2 || n_caps > 64) return h->fail(NDIT_ERR_INVALID, "need 1..63 region captions + the unconditional one (got %d rows)", n_caps);
    if (T < 1 || T > h->Tmax || global_T < 1 || global_T > h->Tmax) return h->fail(NDIT_ERR_INVALID, "caption length out of range (%d, %d; max %d)", T, global_T, h->Tmax);
    if (h_split < 1 || w_split < 1) return h->fail(NDIT_ERR_INVALID, "h_split / w_split must be >= 1");
    // model.py:879-883 indexes region_mask[region_id] with region_id up to h_split * w_split - 1: beyond the caption rows it raises
    if (h_split * w_split - 1 >= n_caps) return h->fail(NDIT_ERR_INVALID, "region id %d >= %d caption rows (the reference raises IndexError)", h_split * w_split - 1, n_caps);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n_caps > h->cap_rows_max) {            // grow the caption buffers (weights stay; like ndit_reserve)
        CK(cudaDeviceSynchronize());
        const int old_req = h->cap_rows_req;
        free_workspace(h);
        h->cap_rows_req = n_caps;
        if (int r = alloc_workspace(h)) {
            free_workspace(h);
            h->cap_rows_req = old_req;
            char msg[512];
            snprintf(msg, sizeof(msg), "%s", h->err);
            if (alloc_workspace(h)) return h->fail(NDIT_ERR_NOMEM, "ndit_set_caption_regions: could not restore the workspace after: %s", msg);
            return h->fail(r, "ndit_set_caption_regions(%d caption rows): %s", n_caps, msg);
        }
    }
    // adaLN conditioning: the pooled GLOBAL caption, one row broadcast to cond and uncond (model.py:866-870: cap_emb [1, D])
    mask_to_u8_kernel<<<(global_T + 255) / 256, 256, 0, s>>>(h->stage_mask, global_mask, global_T);
    CKL(cudaGetLastError());
    if (int e = set_caption_impl(h, static_cast<const bf16*>(cap), mask, n_caps, T, static_cast<const bf16*>(global_cap), h->stage_mask, 1,
                                 global_T, s))
        return e;
    CK(cudaMemcpyAsync(h->capemb + h->cd, h->capemb, sizeof(float) * h->cd, cudaMemcpyDeviceToDevice, s));
    h->attn_plans_valid = false;
    h->cap_batch = 2;
    h->cap_rows = n_caps;
    h->cap_T = T;
    h->region_cond = n_caps - 1; h->region_hs = h_split; h->region_ws = w_split;
    return NDIT_OK;
}

extern "C" int ndit_set_labels(ndit_handle h, const int64_t* labels, int32_t batch, void* stream) {
    if (!h || !labels) return NDIT_ERR_INVALID;
    if (!h->cls) return h->fail(NDIT_ERR_STATE, "ndit_set_labels: this engine is caption-conditioned (num_classes == 0)");
    if (!h->finalized) return h->fail(NDIT_ERR_STATE, "weights not finalized");
    if (batch < 1 || batch > h->Bmax) return h->fail(NDIT_ERR_INVALID, "label batch out of range (%d)", batch);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CKL(gather_label_rows(h->Yemb, reinterpret_cast<const long long*>(labels), h->capemb, batch, h->cfg.num_classes + 1, h->cd, s));
    if (batch != h->cap_batch) h->attn_plans_valid = false;
    h->cap_batch = batch;
    h->cap_rows = batch;
    h->cap_T = 0;
    return NDIT_OK;
}

// ------------------------------------------------------------------------------------ forward

// Tensor maps of the fused attention kernel.  q/k are read straight from the fused qkv buffer [B*N, Wq] through 3-D maps
// (head_dim, head, token); the 64-wide box covers elements [0,64) (head_dim 48: 48..63 are out-of-bounds zeros), the
// 16-wide box elements [64,80) of head_dim 72.  V^T buffers are [group][vrows][tokens].
static int build_attn_maps(AttnPlan* a, const bf16* qkv, int Wq, const bf16* vt, const bf16* kvy, const bf16* vyt, int B, int N,
                           int T, int H, int Hkv, int hd, int bkv = 128, int cap_rows = 0) {
    const uint64_t Bc = cap_rows > 0 ? cap_rows : B;     // caption rows (region mode: more captions than batch rows)
    const int vrows = attn_vrows(hd), KV = Hkv * hd, Tpad = (T + 7) / 8 * 8, Npad = (N + 7) / 8 * 8;
    const uint64_t rs = (uint64_t)Wq * 2, M = (uint64_t)B * N;
    int e = 0;
    e |= make_tmap_3d(&a->tmQ64, qkv, hd, H, M, hd * 2, rs, 64, 1, 128, 128);
    e |= make_tmap_3d(&a->tmK64, qkv + (size_t)H * hd, hd, Hkv, M, hd * 2, rs, 64, 1, bkv, 128);
    e |= make_tmap_3d(&a->tmVt, vt, N, vrows, (uint64_t)B * Hkv, (uint64_t)Npad * 2, (uint64_t)Npad * vrows * 2, 64, vrows, 1, 128);
    if (hd > 64) {
        e |= make_tmap_3d(&a->tmQ16, qkv, hd, H, M, hd * 2, rs, 16, 1, 128, 32);
        e |= make_tmap_3d(&a->tmK16, qkv + (size_t)H * hd, hd, Hkv, M, hd * 2, rs, 16, 1, bkv, 32);
    }
    if (T > 0) {
        e |= make_tmap_3d(&a->tmKy64, kvy, hd, Hkv, Bc * T, hd * 2, (uint64_t)2 * KV * 2, 64, 1, bkv, 128);
        if (hd > 64) e |= make_tmap_3d(&a->tmKy16, kvy, hd, Hkv, Bc * T, hd * 2, (uint64_t)2 * KV * 2, 16, 1, bkv, 32);
        e |= make_tmap_3d(&a->tmVyt, vyt, Tpad, vrows, Bc * Hkv, (uint64_t)Tpad * 2, (uint64_t)Tpad * vrows * 2, 64, vrows, 1, 128);
    }
    a->B = B; a->N = N; a->T = T; a->H = H; a->Hkv = Hkv; a->hd = hd; a->bkv = bkv;
    return e;
}

static int ensure_plans(ndit_engine* h, int batch, int N) {
    const int M = batch * N;
    const size_t D = h->D, F = h->F, L = h->L, Wq = h->Wq, KV = (size_t)h->Hkv * h->hd, C = h->C;
    if (M != h->plan_M) {
        const size_t S = h->S;      // FFN weight slots per layer (experts); slot plans write to h->o unless the launch overrides C
        h->p_qkv.resize(L); h->p_wo.resize(L); h->p_w13.resize(L * S); h->p_w2.resize(L * S);
        for (size_t l = 0; l < L; ++l) {
            int e = 0;
            e |= make_gemm_plan(&h->p_qkv[l], h->u, (int)D, h->Wqkv + l * Wq * D, h->qkv, (int)Wq, M, (int)Wq, (int)D, EPI_STORE, h->num_sms);
            e |= make_gemm_plan(&h->p_wo[l], h->attn, (int)D, h->Wo + l * D * D, h->o, (int)D, M, (int)D, (int)D, EPI_STORE, h->num_sms);
            for (size_t sl = 0; sl < S; ++sl) {
                const size_t i = l * S + sl;
                e |= make_gemm_plan(&h->p_w13[i], h->u, (int)D, h->W13 + i * 2 * F * D, h->hbuf, (int)F, M, (int)(2 * F), (int)D, EPI_SWIGLU, h->num_sms);
                e |= make_gemm_plan(&h->p_w2[i], h->hbuf, (int)F, h->W2 + i * D * F, h->o, (int)D, M, (int)D, (int)F, EPI_STORE, h->num_sms);
            }
            if (e) return h->fail(NDIT_ERR_CUDA, "%s", tmap_last_error());
        }
        h->p_w13t.clear(); h->p_w2t.clear();
        for (int f = 0; f < h->NF; ++f) {
            if (h->ffn_kind[f] != 1) continue;
            const int E = h->ffn_E[f];
            h->p_w13t.resize(L * 2); h->p_w2t.resize(L * 2);
            for (size_t l = 0; l < L; ++l) {
                const size_t base = l * S + h->ffn_slot0[f];
                for (int k = 0; k < 2; ++k) {
                    GemmPlan& a = h->p_w13t[l * 2 + k];
                    GemmPlan& b = h->p_w2t[l * 2 + k];
                    int e = make_gemm_plan(&a, h->u, (int)D, h->W13 + base * 2 * F * D, h->hbuf, (int)F, M, (int)(2 * F), (int)D, EPI_SWIGLU,
                                           h->num_sms, 1, (int)(E * 2 * F));
                    e |= make_gemm_plan(&b, h->hbuf, (int)F, h->W2 + base * D * F, h->oE + (size_t)k * M * D, (int)D, M, (int)D, (int)F,
                                        EPI_STORE, h->num_sms, 1, (int)(E * D));
                    if (e) return h->fail(NDIT_ERR_CUDA, "%s", tmap_last_error());
                    a.w_row_off = h->tsel + l * 2 + k; a.w_row_mul = (int)(2 * F);
                    b.w_row_off = h->tsel + l * 2 + k; b.w_row_mul = (int)D;
                }
            }
        }
        h->p_w13g.clear(); h->p_w2g.clear();
        for (int f = 0; f < h->NF; ++f) {
            if (h->ffn_kind[f] != 2 || h->u_perm == nullptr) continue;
            // grouped GEMM over the gathered rows: the maps span the whole gathered buffer, every expert's launch covers at most
            // roundup(M, 256) rows of it, and which rows (offset, padded count) is read from device memory by the kernel
            const int E = h->ffn_E[f];
            const int R = 2 * M + E * 256, Mg = (M + 255) / 256 * 256;
            h->p_w13g.resize(L * E); h->p_w2g.resize(L * E);
            for (size_t l = 0; l < L; ++l) {
                const size_t base = l * S + h->ffn_slot0[f];
                for (int x = 0; x < E; ++x) {
                    GemmPlan& a = h->p_w13g[l * E + x];
                    GemmPlan& b = h->p_w2g[l * E + x];
                    int e = make_gemm_plan(&a, h->u_perm, (int)D, h->W13 + (base + x) * 2 * F * D, h->h_perm, (int)F, R, (int)(2 * F), (int)D, EPI_SWIGLU, h->num_sms);
                    e |= make_gemm_plan(&b, h->h_perm, (int)F, h->W2 + (base + x) * D * F, h->o_perm, (int)D, R, (int)D, (int)F, EPI_STORE, h->num_sms);
                    if (e) return h->fail(NDIT_ERR_CUDA, "%s", tmap_last_error());
                    a.M = Mg; b.M = Mg;
                    a.rows = GemmRowWin{h->rt + 8 + x, h->rt + 16 + x};
                    b.rows = a.rows;
                }
            }
        }
        if (make_gemm_plan(&h->p_final, h->u, (int)D, h->Wout, h->tok, h->O, M, h->O, (int)D, EPI_STORE, h->num_sms, 0))
            return h->fail(NDIT_ERR_CUDA, "%s", tmap_last_error());
        h->p_final.bias = h->bout;
        h->plan_M = M;
        h->attn_plans_valid = false;
    }
    if (!h->attn_plans_valid || h->plan_B != batch || h->plan_N != N || h->plan_T != h->cap_T) {
        const int T = h->cap_T, Tpad = (T + 7) / 8 * 8;
        h->p_attn.resize(L);
        for (size_t l = 0; l < L; ++l) {
            AttnPlan& a = h->p_attn[l];
            memset(&a, 0, sizeof(a));
            const size_t crows = h->cls ? (size_t)batch : (size_t)h->cap_rows;
            const bf16* kvy = h->kvy + l * crows * T * 2 * KV;
            const bf16* vyt = h->vyt + l * crows * h->Hkv * h->vrows * Tpad;
            const bool gen3 = attn_gen(h->attn_tp) == 3 && h->region_cond == 0;     // region-masked captions: first-generation kernel
            if (build_attn_maps(&a, h->qkv, (int)Wq, h->vt, kvy, vyt, batch, N, T, h->H, h->Hkv, h->hd,
                                gen3 ? attention_hr_bkv(h->hd) : 128, (int)crows))
                return h->fail(NDIT_ERR_CUDA, "%s", tmap_last_error());
            a.ymask = h->ymask;
            a.gate_tanh = h->gate_tanh + l * h->H;
            a.out = h->attn;
        }
        (void)D;
        (void)C;
        h->plan_B = batch; h->plan_N = N; h->plan_T = h->cap_T;
        h->attn_plans_valid = true;
        h->vt_ones_valid = false;       // V^T layout depends on (batch, N): rewrite its all-ones row
    }
    return 0;
}

static int get_rope(ndit_engine* h, int Hp, int Wp, float theta, float lin, cudaStream_t s, const float2** out) {
    for (int i = 0; i < 2; ++i) {
        RopeSlot& r = h->rope[i];
        if (r.Hp == Hp && r.Wp == Wp && r.theta == theta && r.lin == lin) { *out = r.tab; return 0; }
    }
    RopeSlot& r = h->rope[h->rope_next];
    h->rope_next ^= 1;
    CKL(rope_table(r.tab, Hp, Wp, h->hd, theta, lin, h->flag ? 1 : 0, s));
    r.Hp = Hp; r.Wp = Wp; r.theta = theta; r.lin = lin;
    *out = r.tab;
    return 0;
}

// Variable-resolution list input of NextDiT.forward (model.py:789-834): every row its own latent size; rows are padded with the
// learned pad token to the longest row, keys beyond a row's own tokens are masked in the self-attention.
struct ListArgs {
    const void* const* xs;      // host array of device pointers, bf16 [C, heights[i], widths[i]]
    const int32_t* heights;
    const int32_t* widths;
    void* const* outs;          // host array of device pointers, bf16 [C, heights[i], widths[i]]
};

// t_rows == nullptr: forward_with_cfg (cond/uncond pair rows, one timestep t, guidance in the unpatchify);
// t_rows != nullptr (host array of `batch` floats): plain NextDiT.forward - every row its own sample and timestep, no guidance;
// list != nullptr (plain only): variable-resolution rows (x, out, Hh, Ww unused).
static int forward_impl(ndit_engine* h, const bf16* x, float t, int batch, int Hh, int Ww, const ndit_step_params* sp,
                        bf16* out, cudaStream_t s, const float* t_rows = nullptr, const ListArgs* list = nullptr) {
    const bool plain = t_rows != nullptr;
    if (list != nullptr && (!plain || h->flag || h->cls))
        return h->fail(NDIT_ERR_INVALID, "list input is the plain forward of the text-conditioned Next-DiT");
    if (!h->finalized) return h->fail(NDIT_ERR_STATE, "weights not finalized");
    struct PdlGuard {       // per-launch CUDA events (profile mode) and overlapping launches do not mix
        int saved;
        explicit PdlGuard(int v) : saved(g_pdl) { g_pdl = v; }
        ~PdlGuard() { g_pdl = saved; }
    } pdl_guard(h->profile != 0 ? 0 : h->pdl);
    if (batch != h->cap_batch) return h->fail(NDIT_ERR_STATE, "caption not set for batch %d (have %d)", batch, h->cap_batch);
    if (!plain && (batch < 2 || (batch & 1) || batch > h->Bmax)) return h->fail(NDIT_ERR_INVALID, "batch must be even and <= %d", h->Bmax);
    if (plain && (batch < 1 || batch > h->Bmax)) return h->fail(NDIT_ERR_INVALID, "batch must be in 1..%d", h->Bmax);
    if (h->region_cond > 0 && (list != nullptr || batch != 2))
        return h->fail(NDIT_ERR_INVALID, "region-masked captions (ndit_set_caption_regions) drive one cond / uncond pair of rows (batch 2, tensor input)");
    if (plain) {
        t = t_rows[0];
        bool uniform = true;
        for (int b = 1; b < batch; ++b) uniform = uniform && t_rows[b] == t;
        if (!uniform && h->cfg.moe_time_experts > 0)
            return h->fail(NDIT_ERR_INVALID, "the time-gated mixture of experts selects one expert pair per call: all rows need the same timestep");
        CK(cudaMemcpyAsync(h->trow, t_rows, sizeof(float) * batch, cudaMemcpyHostToDevice, s));
    }
    int lens[64];
    if (list != nullptr) {
        if (batch > 64) return h->fail(NDIT_ERR_INVALID, "list input: at most 64 rows");
        Hh = 2; Ww = 0;             // N = the longest row (below); Hp / Wp are per row
        for (int i = 0; i < batch; ++i) {
            const int hi = list->heights[i], wi = list->widths[i];
            if ((hi & 1) || (wi & 1) || hi <= 0 || wi <= 0 || !list->xs[i] || !list->outs[i]) return h->fail(NDIT_ERR_INVALID, "list input: latent H/W must be even");
            if (hi / 2 > 384 || wi / 2 > 384) return h->fail(NDIT_ERR_INVALID, "rope table covers 384x384 patches (model.py:733)");
            lens[i] = (hi / 2) * (wi / 2);
            if (lens[i] > Ww) Ww = lens[i];
        }
        Ww *= 2;                    // so that N = (Hh / 2) * (Ww / 2) = the longest row
        CK(cudaMemcpyAsync(h->kvlen, lens, sizeof(int) * batch, cudaMemcpyHostToDevice, s));
    }
    if ((Hh & 1) || (Ww & 1) || Hh <= 0 || Ww <= 0) return h->fail(NDIT_ERR_INVALID, "latent H/W must be even");
    const int Hp = Hh / 2, Wp = Ww / 2, eol = h->flag ? 1 : 0;
    const int N = Hp * (Wp + eol), M = batch * N, Npad = (N + 7) / 8 * 8;     // Flag-DiT: one [eol] token per row of patches
    if (N > h->cfg.max_tokens) return h->fail(NDIT_ERR_INVALID, "%d tokens > max_tokens %d", N, h->cfg.max_tokens);
    if (!h->flag && list == nullptr && (Hp > 384 || Wp > 384)) return h->fail(NDIT_ERR_INVALID, "rope table covers 384x384 patches (model.py:733)");
    if (h->flag && N > 40000) return h->fail(NDIT_ERR_INVALID, "rope table covers 40000 tokens (lumina_t2i model.py:722-727)");
    if (int e = ensure_plans(h, batch, N)) return e;
    const int D = h->D, L = h->L, hd = h->hd;
    if (!h->vt_ones_valid) {
        CK(cudaMemsetAsync(h->vt, 0, (size_t)batch * h->Hkv * h->vrows * Npad * sizeof(bf16), s));
        CKL(fill_ones_row(h->vt, Npad, 0, batch * h->Hkv, N, hd, h->vrows, 1, s));
        h->vt_ones_valid = true;
    }
    const int NCH = h->NCH;
    const int mod_stride = L * NCH * D + h->FD * D;
    float lin, ntk;
    if (h->cls || h->flag || plain) {   // plain forward: whatever table self.freqs_cis holds (model.py:735,839) = these two factors
        // DiT_Llama.precompute_freqs_cis(rope_scaling_factor, ntk_factor) (models.py:977-1012; lumina_t2i model.py:925-960)
        lin = sp->scale_factor > 0.f ? sp->scale_factor : 1.0f;
        ntk = sp->ntk_factor > 0.f ? sp->ntk_factor : 1.0f;
    } else if (t < sp->scale_watershed) {   // time-aware RoPE scaling (model.py:944-952)
        lin = sp->scale_factor; ntk = 1.0f;
    } else {
        lin = 1.0f; ntk = sp->scale_factor;
    }
    const float theta = 10000.0f * ntk;
    const float2* rope = nullptr;
    int rope_rows_per_batch = N;       // ln_rope_qk indexes the table with (row % rope_rows_per_batch)
    if (list != nullptr) {
        // per-row tables: freqs_cis[:H/2, :W/2] of each image, padded by repeating its last position (model.py:796,817-824)
        const size_t need = (size_t)M * (hd / 2);
        if (need > h->rope_rows_cap) {
            if (h->rope_rows) cudaFree(h->rope_rows);
            h->rope_rows = nullptr; h->rope_rows_cap = 0;
            CK(cudaMalloc(&h->rope_rows, need * sizeof(float2)));
            h->rope_rows_cap = need;
        }
        for (int i = 0; i < batch; ++i) {
            float2* tab = h->rope_rows + (size_t)i * N * (hd / 2);
            CKL(rope_table(tab, list->heights[i] / 2, list->widths[i] / 2, hd, theta, lin, 0, s));
            CKL(broadcast_row(tab + (size_t)lens[i] * (hd / 2), tab + (size_t)(lens[i] - 1) * (hd / 2), N - lens[i], (hd / 2) * (int)sizeof(float2), s));
        }
        rope = h->rope_rows;
        rope_rows_per_batch = M;
    } else if (int e = get_rope(h, h->flag ? N : Hp, h->flag ? 1 : Wp, theta, lin, s, &rope)) return e;
    float scale_self;
    if (sp->proportional_attn) {
        if (sp->base_seqlen <= 1) return h->fail(NDIT_ERR_INVALID, "proportional_attn needs base_seqlen > 1");
        scale_self = (float)sqrt(log((double)N) / log((double)sp->base_seqlen) / (double)hd);   // model.py:373-376
    } else {
        scale_self = (float)sqrt(1.0 / (double)hd);
    }
    const float scale_cross = (float)(1.0 / sqrt((double)hd));
    // region rectangles of the compositional model: H // h_split // patch_size x W // w_split // patch_size tokens (model.py:875)
    AttnRegion region{0, 0, 0, 0, 0, 0};
    if (h->region_cond > 0) {
        region = AttnRegion{h->region_cond, Wp, Hp / h->region_hs, Wp / h->region_ws, h->region_hs, h->region_ws};
        if (region.hp < 1 || region.wp < 1) return h->fail(NDIT_ERR_INVALID, "%d x %d regions do not fit a %d x %d token grid", h->region_hs, h->region_ws, Hp, Wp);
    }

    if (list != nullptr) {
        for (int i = 0; i < batch; ++i) {
            bf16* Xi = h->X + (size_t)i * N * D;
            PROF(KC_ROWWISE, patch_embed(static_cast<const bf16*>(list->xs[i]), h->Wx, h->bx, nullptr, Xi, 1, 1, h->cfg.in_channels, list->heights[i], list->widths[i], D, s));
            PROF(KC_ROWWISE, broadcast_row(Xi + (size_t)lens[i] * D, h->pad_token, N - lens[i], D * (int)sizeof(bf16), s));
        }
    } else
    PROF(KC_ROWWISE, patch_embed(x, h->Wx, h->bx, h->flag ? h->eol_token : nullptr, h->X, batch, plain ? batch : batch / 2, h->cfg.in_channels, Hh, Ww, D, s));
    PROF(KC_COND, cond_prepare(t, plain ? h->trow : nullptr, nullptr, nullptr, nullptr, nullptr, h->tf, nullptr, batch, 0, 0, 0, s));
    PROF(KC_COND, gemv_rows(h->tf, h->Wt0, h->bt0, nullptr, h->h1, nullptr, batch, h->cd, 256, 0, POST_SILU, 0, 0, 0, s));
    // sc = bf16(silu(c)), c = bf16(temb + cap_emb)
    PROF(KC_COND, gemv_rows(h->h1, h->Wt2, h->bt2, h->capemb, h->sc, nullptr, batch, h->cd, h->cd, 0, POST_SILU, 0, 0, 0, s));
    PROF(KC_COND, gemv_rows(h->sc, h->Wada, h->bada, nullptr, nullptr, h->mod, batch, mod_stride, h->cd, 0, POST_ADALN, D, L * NCH,
                            h->flag ? ADALN_FLAG : (h->cls ? ADALN_CLASS : ADALN_NEXT), s));
    // time-gated MoE (Next-DiT-MoE models.py:459-477): the gate sees only the timestep embedding, so one pair of experts
    // serves the whole batch in every layer.  Logits for all layers in one GEMV; the top-2 selection stays on the device
    // (moe_time_select) and the expert GEMMs read "which expert" from device memory: no host round trip, graph-capturable.
    const int Et = h->cfg.moe_time_experts;
    if (Et > 0) {
        PROF(KC_COND, gemv_rows(h->h1, h->Wt2, h->bt2, nullptr, h->temb, nullptr, batch, h->cd, h->cd, 0, POST_NONE, 0, 0, 0, s));
        PROF(KC_COND, gemv_rows(h->temb, h->Wg_time, nullptr, nullptr, h->tlogits, nullptr, batch, L * Et, h->cd, 0, POST_NONE, 0, 0, 0, s));
        PROF(KC_COND, moe_time_select(h->tlogits, L, Et, h->tsel, h->tw, s));
    }
    // per-layer modulation chunks (offsets in units of D): Next-DiT [1+scale_msa, tanh gate_msa, 1+scale_mlp, tanh gate_mlp];
    // Flag-DiT [shift_msa, 1+scale_msa, gate_msa, shift_mlp, 1+scale_mlp, gate_mlp]
    const int o_sc1 = h->flag ? 1 : 0, o_g1 = h->flag ? 2 : 1, o_sc2 = h->flag ? 4 : 2, o_g2 = h->flag ? 5 : 3;
    const bf16* mod0 = h->mod;
    PROF(KC_ROWWISE, resid_rms_mod(h->X, nullptr, nullptr, nullptr, h->an1, mod0 + (size_t)o_sc1 * D, h->flag ? mod0 : nullptr, h->u, M, N, D,
                                   mod_stride, h->cfg.norm_eps, s));
    for (int l = 0; l < L; ++l) {
        const bf16* ml = h->mod + (size_t)l * NCH * D;
        const bf16* mn = ml + (size_t)NCH * D;      // next layer's chunks
        // value heads go straight to the V^T buffer from the GEMM epilogue (the debug attention reads V from the qkv buffer)
        const bool vt_fused = h->vt_epi && !h->attn_ref;
        h->p_qkv[l].vt = vt_fused ? GemmVtOut{h->vt, (h->H + h->Hkv) * hd, hd, h->Hkv, h->vrows, Npad, N} : GemmVtOut{};
        PROF(KC_GEMM_QKV, gemm_bf16_tn(h->p_qkv[l], s));
        if (h->cfg.no_qk_norm)      // qk_norm=False: q_norm / k_norm are Identity (model.py:219), only the rotary embedding remains
            PROF(KC_ROWWISE, rope_qk(h->qkv, h->Wq, rope, M, rope_rows_per_batch, h->H, h->Hkv, hd, s));
        else
        PROF(KC_ROWWISE, ln_rope_qk(h->qkv, h->Wq, h->qn_w + (size_t)l * D, h->qn_b + (size_t)l * D, h->kn_w + (size_t)l * h->Hkv * hd,
                       h->kn_b + (size_t)l * h->Hkv * hd, rope, M, rope_rows_per_batch, h->H, h->Hkv, hd, s));
        if (h->attn_ref) {
            PROF(KC_ATTN, attention_ref(h->qkv, h->Wq, h->kvy + (size_t)l * h->cap_rows * h->cap_T * 2 * h->Hkv * hd, 2 * h->Hkv * hd, h->ymask,
                              h->gate_tanh + (size_t)l * h->H, h->attn, batch, N, h->cap_T, h->H, h->Hkv, hd, scale_self,
                              scale_cross, s, list != nullptr ? h->kvlen : nullptr, region));
        } else {
            if (!vt_fused)
                PROF(KC_ROWWISE, transpose_v(h->qkv, h->Wq, (h->H + h->Hkv) * hd, 0, h->vt, Npad, 0, batch, N, h->Hkv, hd, h->vrows, 1, s));
            AttnPlan& a = h->p_attn[l];
            a.scale_self = scale_self;
            a.scale_cross = scale_cross;
            a.kv_len = list != nullptr ? h->kvlen : nullptr;
            a.region = region;
            PROF(KC_ATTN, (attn_gen(h->attn_tp) == 3 && region.n_cond == 0) ? attention_fused_hr(a, s) : attention_fused(a, s));
        }
        PROF(KC_GEMM_WO, gemm_bf16_tn(h->p_wo[l], s));
        PROF(KC_ROWWISE, resid_rms_mod(h->X, h->o, h->flag ? nullptr : h->an2 + (size_t)l * D, ml + (size_t)o_g1 * D, h->fn1 + (size_t)l * D,
                                       ml + (size_t)o_sc2 * D, h->flag ? ml + 3 * (size_t)D : nullptr, h->u, M, N, D, mod_stride,
                                       h->cfg.norm_eps, s));
        for (int f = 0; f < h->NF; ++f) {
            // ---- FFN sub-block f -> h->o
            const size_t base = (size_t)l * h->S + h->ffn_slot0[f];
            const size_t MD = (size_t)M * D;
            if (h->ffn_kind[f] == 0) {
                PROF(KC_GEMM_W13, gemm_bf16_tn(h->p_w13[base], s));
                PROF(KC_GEMM_W2, gemm_bf16_tn(h->p_w2[base], s));
            } else if (h->ffn_kind[f] == 1) {           // time-gated: the two selected experts, ascending index
                for (int k = 0; k < 2; ++k) {
                    PROF(KC_GEMM_W13, gemm_bf16_tn(h->p_w13t[(size_t)l * 2 + k], s));
                    PROF(KC_GEMM_W2, gemm_bf16_tn(h->p_w2t[(size_t)l * 2 + k], s));
                }
                PROF(KC_ROWWISE, moe_combine(h->oE, MD, 2, nullptr, h->tw + (size_t)l * 2, h->o, M, D, s));
            } else {                                    // token-gated: every expert runs densely, the gate weights select
                const int E = h->ffn_E[f];
                PROF(KC_ROWWISE, moe_space_gate(h->u, h->Wg_space + (size_t)l * E * D, h->wtok, M, D, E, s));
                if (h->moe_grouped && !h->p_w13g.empty()) {
                    // every expert only sees the tokens that selected it (models1.py:471-476): gather, grouped GEMMs, routed combine
                    int* rt = h->rt;
                    PROF(KC_ROWWISE, moe_route(h->u, h->wtok, rt, rt + 8, rt + 16, rt + 32, h->rpos, h->u_perm, M, D, E, s));
                    for (int e = 0; e < E; ++e) {
                        PROF(KC_GEMM_W13, gemm_bf16_tn(h->p_w13g[(size_t)l * E + e], s));
                        PROF(KC_GEMM_W2, gemm_bf16_tn(h->p_w2g[(size_t)l * E + e], s));
                    }
                    PROF(KC_ROWWISE, moe_combine_routed(h->o_perm, h->rpos, h->wtok, h->o, M, D, E, s));
                } else {
                for (int e = 0; e < E; ++e) {
                    PROF(KC_GEMM_W13, gemm_bf16_tn(h->p_w13[base + e], s));
                    GemmPlan p2 = h->p_w2[base + e];
                    p2.C = h->oE + e * MD;
                    PROF(KC_GEMM_W2, gemm_bf16_tn(p2, s));
                }
                PROF(KC_ROWWISE, moe_combine(h->oE, MD, E, h->wtok, nullptr, h->o, M, D, s));
                }
            }
            // ---- gated (post-normed) residual, then the pre-norm + modulation of whatever runs next
            const bf16* post = h->flag ? nullptr : h->fn2 + ((size_t)f * L + l) * D;
            const bf16* gch = h->flag ? ml + (size_t)o_g2 * D : ml + (size_t)(3 + 2 * f) * D;
            if (f + 1 < h->NF) {
                PROF(KC_ROWWISE, resid_rms_mod(h->X, h->o, post, gch, h->fn1 + (size_t)l * D, ml + (size_t)(4 + 2 * f) * D, nullptr, h->u, M, N, D,
                                               mod_stride, h->cfg.norm_eps, s));
            } else if (l + 1 < L) {
                PROF(KC_ROWWISE, resid_rms_mod(h->X, h->o, post, gch, h->an1 + (size_t)(l + 1) * D, mn + (size_t)o_sc1 * D, h->flag ? mn : nullptr,
                                               h->u, M, N, D, mod_stride, h->cfg.norm_eps, s));
            } else {
                // final adaLN: Next-DiT T2I [scale]; class-conditional and Flag-DiT [shift | scale]
                const bf16* fin = h->mod + (size_t)L * NCH * D;
                const bool fsh = h->cls || h->flag;
                PROF(KC_ROWWISE, final_norm(h->X, h->o, post, gch, fsh ? fin + D : fin, fsh ? fin : nullptr, h->u, M, N, D, mod_stride,
                                            h->cfg.norm_eps, s, h->tap_layer == l ? h->X : nullptr));
                PROF(KC_ROWWISE, gemm_bf16_tn(h->p_final, s));
            }
        }
        if (l == h->tap_layer) {    // debug tap: the residual stream X after block l (the last block's X is complete as well:
                                    // final_norm reads it, the post-norm of the second sub-block is already applied)
            if (h->tap_rows < (size_t)M) {
                if (h->tap_buf) cudaFree(h->tap_buf);
                h->tap_buf = nullptr; h->tap_rows = 0;
                CK(cudaMalloc(&h->tap_buf, (size_t)M * D * sizeof(bf16)));
                h->tap_rows = M;
            }
            CK(cudaMemcpyAsync(h->tap_buf, h->X, (size_t)M * D * sizeof(bf16), cudaMemcpyDeviceToDevice, s));
        }
    }
    if (list != nullptr) {
        for (int i = 0; i < batch; ++i)
            PROF(KC_ROWWISE, unpatchify_plain(h->tok + (size_t)i * N * h->O, static_cast<bf16*>(list->outs[i]), 1, h->cfg.in_channels, list->heights[i],
                                              list->widths[i], h->O, 0, s));
    } else if (plain) PROF(KC_ROWWISE, unpatchify_plain(h->tok, out, batch, h->cfg.in_channels, Hh, Ww, h->O, eol, s));
    else PROF(KC_ROWWISE, unpatchify_cfg(h->tok, out, batch / 2, h->cfg.in_channels, Hh, Ww, h->O, sp->cfg_scale, eol, s));
    return NDIT_OK;
}

extern "C" int ndit_forward_cfg(ndit_handle h, const void* x, float t, int32_t batch, int32_t height, int32_t width,
                                const ndit_step_params* sp, void* out, void* stream) {
    if (!h || !x || !sp || !out) return NDIT_ERR_INVALID;
    return forward_impl(h, static_cast<const bf16*>(x), t, batch, height, width, sp, static_cast<bf16*>(out),
                        static_cast<cudaStream_t>(stream));
}

extern "C" int ndit_forward(ndit_handle h, const void* x, const float* t_host, int32_t batch, int32_t height, int32_t width,
                            const ndit_step_params* sp, void* out, void* stream) {
    if (!h || !x || !t_host || !sp || !out) return NDIT_ERR_INVALID;
    return forward_impl(h, static_cast<const bf16*>(x), 0.f, batch, height, width, sp, static_cast<bf16*>(out),
                        static_cast<cudaStream_t>(stream), t_host);
}

extern "C" int ndit_debug_read_residual(ndit_handle h, void* out_dev, int64_t rows, void* stream) {
    if (!h || !out_dev) return NDIT_ERR_INVALID;
    if (!h->tap_buf || rows <= 0 || (size_t)rows > h->tap_rows) return h->fail(NDIT_ERR_STATE, "no tap recorded (option tap_layer, then a forward)");
    cudaError_t e = cudaMemcpyAsync(out_dev, h->tap_buf, (size_t)rows * h->D * sizeof(bf16), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? NDIT_OK : h->fail(NDIT_ERR_CUDA, "ndit_debug_read_residual: %s", cudaGetErrorString(e));
}

extern "C" int ndit_forward_list(ndit_handle h, const void* const* x_dev, const int32_t* heights, const int32_t* widths, const float* t_host,
                                 int32_t batch, const ndit_step_params* sp, void* const* out_dev, void* stream) {
    if (!h || !x_dev || !heights || !widths || !t_host || !sp || !out_dev) return NDIT_ERR_INVALID;
    const ListArgs la{x_dev, heights, widths, out_dev};
    return forward_impl(h, nullptr, 0.f, batch, 0, 0, sp, nullptr, static_cast<cudaStream_t>(stream), t_host, &la);
}

// The fixed-grid solve on the engine's own state buffer (h->ystate); trajectory rows 1.. go to `tr` when non-null.
static int sample_body(ndit_engine* h, int batch, int height, int width, const float* grid, int n_grid, int method,
                       const ndit_step_params* sp, bf16* tr, cudaStream_t s) {
    const size_t count = (size_t)batch * h->cfg.in_channels * height * width;
    bf16* y = h->ystate;
    for (int i = 0; i + 1 < n_grid; ++i) {
        const float t0 = grid[i], t1 = grid[i + 1];
        const float dt = t1 - t0;
        // torchdiffeq semantics with a bf16 state: the model sees t cast to the state dtype, and the 0-dim
        // fp32 tensors dt / dt/2 are cast to bf16 by type promotion when multiplied with the bf16 velocity.
        const float dt_b = host_bf16_round(dt);
        if (method == NDIT_EULER) {
            if (int e = forward_impl(h, y, host_bf16_round(t0), batch, height, width, sp, h->vel, s)) return e;
            CKL(axpy_bf16(y, y, h->vel, dt_b, count, s));
        } else if (method == NDIT_MIDPOINT) {
            const float half_dt = 0.5f * dt;
            if (int e = forward_impl(h, y, host_bf16_round(t0), batch, height, width, sp, h->vel, s)) return e;
            CKL(axpy_bf16(h->ymid, y, h->vel, host_bf16_round(half_dt), count, s));
            if (int e = forward_impl(h, h->ymid, host_bf16_round(t0 + half_dt), batch, height, width, sp, h->vel, s)) return e;
            CKL(axpy_bf16(y, y, h->vel, dt_b, count, s));
        } else {
            // torchdiffeq rk4_alt_step_func (3/8 rule): stage times t0 + dt/3, t0 + 2 dt/3, t1 are fp32 scalars, cast to
            // the state dtype when the model is called
            bf16 *k1 = h->vel, *k2 = h->kbuf[0], *k3 = h->kbuf[1], *k4 = h->kbuf[2];
            const float third = 1.0f / 3.0f, two_thirds = 2.0f / 3.0f;
            if (int e = forward_impl(h, y, host_bf16_round(t0), batch, height, width, sp, k1, s)) return e;
            CKL(rk4_stage(1, h->ymid, y, k1, k2, k3, k4, dt_b, count, s));
            if (int e = forward_impl(h, h->ymid, host_bf16_round(t0 + dt * third), batch, height, width, sp, k2, s)) return e;
            CKL(rk4_stage(2, h->ymid, y, k1, k2, k3, k4, dt_b, count, s));
            if (int e = forward_impl(h, h->ymid, host_bf16_round(t0 + dt * two_thirds), batch, height, width, sp, k3, s)) return e;
            CKL(rk4_stage(3, h->ymid, y, k1, k2, k3, k4, dt_b, count, s));
            if (int e = forward_impl(h, h->ymid, host_bf16_round(t1), batch, height, width, sp, k4, s)) return e;
            CKL(rk4_stage(4, y, y, k1, k2, k3, k4, dt_b, count, s));
        }
        if (tr) CK(cudaMemcpyAsync(tr + (size_t)(i + 1) * count, y, count * 2, cudaMemcpyDeviceToDevice, s));
    }
    return NDIT_OK;
}

// Replays the CUDA graph of this solve (captured the second time the same solve is requested; the first one runs directly and
// leaves every lazily initialised piece - plans, kernel attributes, RoPE tables - in place): ~20 k kernel launches of a
// 30-point solve become one graph launch, which is what the launch-bound small configurations need (class-conditional 600M:
// 151 launches in 2.2 ms per model call).  Returns 1 when the solve ran through a graph, 0 when the caller must run it
// directly, an NDIT_ERR_* code (< 0) on error.
static int sample_graph(ndit_engine* h, int batch, int height, int width, const float* grid, int n_grid, int method,
                        const ndit_step_params* sp, bool with_traj, cudaStream_t s) {
    if (!h->use_graph || h->profile) return 0;
    const size_t count = (size_t)batch * h->cfg.in_channels * height * width;
    if (with_traj) {
        const size_t need = (size_t)n_grid * count;
        if (need > h->traj_cap_elems) {
            if (need * 2 > ((size_t)1 << 30)) return 0;
            bf16* nb = nullptr;
            if (cudaMalloc(&nb, need * 2) != cudaSuccess) { cudaGetLastError(); return 0; }
            for (auto& g : h->graphs) if (g.exec && g.with_traj) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }   // they point at the old buffer
            if (h->traj_buf) {
                cudaStreamSynchronize(s);
                for (size_t i = 0; i < h->ws_allocs.size(); ++i) if (h->ws_allocs[i] == h->traj_buf) { h->ws_allocs.erase(h->ws_allocs.begin() + i); break; }
                cudaFree(h->traj_buf);
            }
            h->traj_buf = nb; h->traj_cap_elems = need; h->ws_allocs.push_back(nb);
        }
    }
    ndit_engine::SolveGraph* hit = nullptr;
    for (auto& g : h->graphs) {
        if (g.batch == batch && g.height == height && g.width == width && g.method == method && g.cap_T == h->cap_T &&
            g.cap_rows == h->cap_rows && g.region_cond == h->region_cond && g.region_hs == h->region_hs && g.region_ws == h->region_ws &&
            g.with_traj == (int)with_traj && g.attn_ref == h->attn_ref && g.attn_tp == h->attn_tp && g.pdl == h->pdl && g.vt_epi == h->vt_epi && g.moe_grouped == h->moe_grouped &&
            (int)g.grid.size() == n_grid && !memcmp(g.grid.data(), grid, n_grid * sizeof(float)) && !memcmp(&g.sp, sp, sizeof(*sp))) {
            hit = &g;
            break;
        }
    }
    if (!hit) {          // first sight of this solve: remember it, run it directly
        if (h->graphs.size() < 4) h->graphs.emplace_back();
        ndit_engine::SolveGraph* slot = &h->graphs[0];
        for (auto& g : h->graphs) { if (g.last_use < slot->last_use) slot = &g; }
        if (slot->exec) { cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
        slot->grid.assign(grid, grid + n_grid);
        slot->batch = batch; slot->height = height; slot->width = width; slot->method = method; slot->cap_T = h->cap_T;
        slot->cap_rows = h->cap_rows; slot->region_cond = h->region_cond; slot->region_hs = h->region_hs; slot->region_ws = h->region_ws;
        slot->with_traj = with_traj; slot->attn_ref = h->attn_ref; slot->attn_tp = h->attn_tp; slot->pdl = h->pdl; slot->vt_epi = h->vt_epi; slot->moe_grouped = h->moe_grouped; slot->sp = *sp;
        slot->launches = 0;
        slot->capture_failed = false;
        slot->last_use = ++h->graph_clock;
        return 0;
    }
    if (hit->capture_failed) return 0;
    if (!hit->exec) {
        const int eol = h->flag ? 1 : 0;
        if (int e = ensure_plans(h, batch, (height / 2) * (width / 2 + eol))) return e;
        cudaGraph_t graph = nullptr;
        // a graph must be self-contained: it (re)builds the V^T ones rows and its RoPE tables itself, because solves of other
        // shapes may run between two replays
        h->vt_ones_valid = false;
        for (auto& r : h->rope) r.Hp = 0;
        // The legacy default stream (what PyTorch hands over unless the caller switched streams) and the per-thread default stream cannot
        // be captured (cudaStreamBeginCapture fails with cudaErrorStreamCaptureUnsupported): record the launch sequence on an engine-owned
        // non-blocking stream instead - nothing executes during capture - and launch the instantiated graph on the caller's stream.
        cudaStream_t cs = s;
        if (s == nullptr || s == cudaStreamLegacy || s == cudaStreamPerThread) {
            if (!h->capture_stream && cudaStreamCreateWithFlags(&h->capture_stream, cudaStreamNonBlocking) != cudaSuccess) {
                cudaGetLastError();
                h->capture_stream = nullptr;
                hit->capture_failed = true;
                return 0;
            }
            cs = h->capture_stream;
        }
        auto give_up = [&]() {               // nothing recorded during a failed capture has run: the caller runs the solve directly
            cudaGetLastError();
            h->vt_ones_valid = false;
            for (auto& r : h->rope) r.Hp = 0;
            hit->capture_failed = true;
            return 0;
        };
        if (cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed) != cudaSuccess) return give_up();
        const int64_t l0 = h->launches;
        const int rc = sample_body(h, batch, height, width, grid, n_grid, method, sp, with_traj ? h->traj_buf : nullptr, cs);
        const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
        const int64_t captured = h->launches - l0;
        h->launches = l0;
        if (rc != NDIT_OK || ce != cudaSuccess || graph == nullptr) {
            if (graph) cudaGraphDestroy(graph);
            return give_up();                // a genuine error (not a capture restriction) shows up again in the direct run
        }
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) return give_up();
        hit->exec = exec; hit->launches = captured;
        hit->rope_after[0] = h->rope[0]; hit->rope_after[1] = h->rope[1]; hit->rope_next_after = h->rope_next;
    }
    hit->last_use = ++h->graph_clock;
    if (cudaGraphLaunch(hit->exec, s) != cudaSuccess) return h->fail(NDIT_ERR_CUDA, "cudaGraphLaunch of the captured solve failed");
    h->launches += hit->launches;
    h->graph_replays += 1;
    h->rope[0] = hit->rope_after[0]; h->rope[1] = hit->rope_after[1]; h->rope_next = hit->rope_next_after;
    h->vt_ones_valid = (h->plan_B == batch && h->plan_N == (height / 2) * (width / 2 + (h->flag ? 1 : 0)));
    return 1;
}

extern "C" int ndit_sample(ndit_handle h, const void* z, int32_t batch, int32_t height, int32_t width, const float* grid,
                           int32_t n_grid, int32_t method, const ndit_step_params* sp, void* traj, void* final_out,
                           void* stream) {
    if (!h || !z || !grid || !sp || !final_out) return NDIT_ERR_INVALID;
    if (n_grid < 2) return h->fail(NDIT_ERR_INVALID, "need at least 2 grid points");
    if (method != NDIT_EULER && method != NDIT_MIDPOINT && method != NDIT_RK4) return h->fail(NDIT_ERR_INVALID, "method must be euler, midpoint or rk4");
    if (!h->finalized) return h->fail(NDIT_ERR_STATE, "weights not finalized");
    if (batch != h->cap_batch) return h->fail(NDIT_ERR_STATE, "caption not set for batch %d (have %d)", batch, h->cap_batch);
    if (batch < 2 || (batch & 1) || batch > h->Bmax) return h->fail(NDIT_ERR_INVALID, "batch must be even and <= %d", h->Bmax);
    if ((height & 1) || (width & 1) || height <= 0 || width <= 0) return h->fail(NDIT_ERR_INVALID, "latent H/W must be even");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t count = (size_t)batch * h->cfg.in_channels * height * width;
    if (count > (size_t)h->Bmax * h->cfg.in_channels * h->cfg.max_tokens * 4) return h->fail(NDIT_ERR_INVALID, "latent too large");
    bf16* y = h->ystate;
    bf16* tr = static_cast<bf16*>(traj);
    CK(cudaMemcpyAsync(y, z, count * 2, cudaMemcpyDeviceToDevice, s));
    if (tr) CK(cudaMemcpyAsync(tr, z, count * 2, cudaMemcpyDeviceToDevice, s));
    const int g = sample_graph(h, batch, height, width, grid, n_grid, method, sp, tr != nullptr, s);
    if (g < 0) return g;
    if (g == 1) {
        if (tr) CK(cudaMemcpyAsync(tr + count, h->traj_buf + count, (size_t)(n_grid - 1) * count * 2, cudaMemcpyDeviceToDevice, s));
    } else {
        if (int e = sample_body(h, batch, height, width, grid, n_grid, method, sp, tr, s)) return e;
    }
    if (final_out != y) CK(cudaMemcpyAsync(final_out, y, count * 2, cudaMemcpyDeviceToDevice, s));
    return NDIT_OK;
}

extern "C" int ndit_sample_sde(ndit_handle h, const void* z, int32_t batch, int32_t height, int32_t width, int32_t n_steps, int32_t method,
                               const ndit_sde_point* pts, float dt, float sqrt_dt, float half_dt, const void* noise, const ndit_step_params* sp,
                               void* traj, void* stream) {
    if (!h || !z || !pts || !noise || !sp || !traj) return NDIT_ERR_INVALID;
    if (n_steps < 1 || (method != 0 && method != 1)) return h->fail(NDIT_ERR_INVALID, "ndit_sample_sde: n_steps >= 1, method 0 (Euler-Maruyama) or 1 (Heun)");
    if (!h->finalized) return h->fail(NDIT_ERR_STATE, "weights not finalized");
    if (batch != h->cap_batch) return h->fail(NDIT_ERR_STATE, "caption not set for batch %d (have %d)", batch, h->cap_batch);
    if (batch < 2 || (batch & 1) || batch > h->Bmax) return h->fail(NDIT_ERR_INVALID, "batch must be even and <= %d", h->Bmax);
    if ((height & 1) || (width & 1) || height <= 0 || width <= 0) return h->fail(NDIT_ERR_INVALID, "latent H/W must be even");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t count = (size_t)batch * h->cfg.in_channels * height * width;
    if (count > (size_t)h->Bmax * h->cfg.in_channels * h->cfg.max_tokens * 4) return h->fail(NDIT_ERR_INVALID, "latent too large");
    bf16* y = h->ystate;
    bf16* tr = static_cast<bf16*>(traj);
    const bf16* nz = static_cast<const bf16*>(noise);
    CK(cudaMemcpyAsync(y, z, count * 2, cudaMemcpyDeviceToDevice, s));
    for (int i = 0; i < n_steps; ++i) {
        const bf16* w = nz + (size_t)i * count;
        if (method == 0) {                       // Euler-Maruyama (integrators.py:33-47)
            const ndit_sde_point& p = pts[i];
            if (int e = forward_impl(h, y, p.t, batch, height, width, sp, h->vel, s)) return e;
            CKL(sde_step(0, y, nullptr, y, h->vel, w, nullptr, nullptr, p.ratio, p.var, p.diffusion, p.sqrt_2diffusion, dt, sqrt_dt, half_dt, count, s));
        } else {                                 // Heun (integrators.py:49-66): points 2i (t) and 2i+1 (t + dt)
            const ndit_sde_point& p = pts[2 * i];
            const ndit_sde_point& q = pts[2 * i + 1];
            bf16 *xhat = h->ymid, *k1 = h->kbuf[0], *xp = h->kbuf[1];
            CKL(sde_step(1, xhat, nullptr, y, nullptr, w, nullptr, nullptr, p.ratio, p.var, p.diffusion, p.sqrt_2diffusion, dt, sqrt_dt, half_dt, count, s));
            if (int e = forward_impl(h, xhat, p.t, batch, height, width, sp, h->vel, s)) return e;
            CKL(sde_step(2, xp, k1, xhat, h->vel, nullptr, nullptr, nullptr, p.ratio, p.var, p.diffusion, p.sqrt_2diffusion, dt, sqrt_dt, half_dt, count, s));
            if (int e = forward_impl(h, xp, q.t, batch, height, width, sp, h->vel, s)) return e;
            CKL(sde_step(3, y, nullptr, xp, h->vel, nullptr, xhat, k1, q.ratio, q.var, q.diffusion, q.sqrt_2diffusion, dt, sqrt_dt, half_dt, count, s));
        }
        CK(cudaMemcpyAsync(tr + (size_t)i * count, y, count * 2, cudaMemcpyDeviceToDevice, s));
    }
    return NDIT_OK;
}

extern "C" int ndit_sample_host(ndit_handle h, const void* z_host, const void* cap_host, const uint8_t* mask_host,
                                int32_t batch, int32_t height, int32_t width, int32_t T, const float* grid, int32_t n_grid,
                                int32_t method, const ndit_step_params* sp, void* final_host, void* stream) {
    if (!h || !z_host || !cap_host || !mask_host || !final_host) return NDIT_ERR_INVALID;
    if (h->cls) return h->fail(NDIT_ERR_INVALID, "ndit_sample_host is the text-conditioned entry point; class-conditional engines use ndit_set_labels + ndit_sample");
    if (batch < 2 || batch > h->Bmax || T < 1 || T > h->Tmax) return h->fail(NDIT_ERR_INVALID, "batch/T out of range");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t count = (size_t)batch * h->cfg.in_channels * height * width;
    if (count > (size_t)h->Bmax * h->cfg.in_channels * h->cfg.max_tokens * 4) return h->fail(NDIT_ERR_INVALID, "latent too large");
    CK(cudaMemcpyAsync(h->stage_z, z_host, count * 2, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->stage_cap, cap_host, (size_t)batch * T * h->C * 2, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->stage_mask, mask_host, (size_t)batch * T, cudaMemcpyHostToDevice, s));
    if (int e = ndit_set_caption(h, h->stage_cap, h->stage_mask, batch, T, stream)) return e;
    if (int e = ndit_sample(h, h->stage_z, batch, height, width, grid, n_grid, method, sp, nullptr, h->stage_z, stream)) return e;
    CK(cudaMemcpyAsync(final_host, h->stage_z, count * 2, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return NDIT_OK;
}

extern "C" int ndit_profile_read(ndit_handle h, float* ms_out, int64_t* count_out, int32_t n_classes) {
    if (!h || !ms_out || !count_out) return NDIT_ERR_INVALID;
    for (int i = 0; i < n_classes; ++i) { ms_out[i] = 0.f; count_out[i] = 0; }
    CK(cudaDeviceSynchronize());
    for (size_t i = 0; i < h->ev_class.size() && 2 * i + 1 < h->ev_used; ++i) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]));
        const int c = h->ev_class[i];
        if (c < n_classes) { ms_out[c] += ms; count_out[c] += 1; }
    }
    h->ev_used = 0;
    h->ev_class.clear();
    return NDIT_OK;
}

// ------------------------------------------------------------------------------------ single ops

// errors of the handle-less entry points are reported through ndit_last_error(NULL), like ndit_create's
static int op_fail(int code, const char* what, cudaError_t e) {
    snprintf(g_create_err, sizeof(g_create_err), "%s: %s", what, e == cudaSuccess ? tmap_last_error() : cudaGetErrorString(e));
    return code;
}
static int op_num_sms() {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

extern "C" int ndit_op_gemm(const void* A, const void* W, void* C, int32_t M, int32_t N, int32_t K, int32_t swiglu, void* stream) {
    GemmPlan p;
    const int ldc = swiglu ? N / 2 : N;
    if (make_gemm_plan(&p, static_cast<const bf16*>(A), K, static_cast<const bf16*>(W), static_cast<bf16*>(C), ldc, M, N, K,
                       swiglu ? EPI_SWIGLU : EPI_STORE, op_num_sms()))
        return op_fail(NDIT_ERR_CUDA, "ndit_op_gemm tensor map", cudaSuccess);
    cudaError_t e = gemm_bf16_tn(p, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? NDIT_OK : op_fail(NDIT_ERR_CUDA, "ndit_op_gemm", e);
}

extern "C" int ndit_op_gemm_bench(const void* A, const void* W, void* C, int32_t M, int32_t N, int32_t K, int32_t swiglu,
                                  int32_t allow_pair, int32_t iters, float* ms_out, void* stream) {
    if (iters <= 0 || !ms_out) return NDIT_ERR_INVALID;
    GemmPlan p;
    const int ldc = swiglu ? N / 2 : N;
    if (make_gemm_plan(&p, static_cast<const bf16*>(A), K, static_cast<const bf16*>(W), static_cast<bf16*>(C), ldc, M, N, K,
                       swiglu ? EPI_SWIGLU : EPI_STORE, op_num_sms(), allow_pair))
        return op_fail(NDIT_ERR_CUDA, "ndit_op_gemm_bench tensor map", cudaSuccess);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = gemm_bf16_tn(p, s);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, s);
    for (int i = 0; i < iters && e == cudaSuccess; ++i) e = gemm_bf16_tn(p, s);
    cudaEventRecord(e1, s);
    cudaError_t e2 = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_out = ms / iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e == cudaSuccess) e = e2;
    return e == cudaSuccess ? (p.pair ? 1 : NDIT_OK) : op_fail(NDIT_ERR_CUDA, "ndit_op_gemm_bench", e);
}

extern "C" int ndit_op_ln_rope(void* qkv, const void* qw, const void* qb, const void* kw, const void* kb, int32_t batch,
                               int32_t Hp, int32_t Wp, int32_t H, int32_t Hkv, int32_t hd, float theta, float linear_factor,
                               int32_t one_d, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float2* tab = nullptr;
    cudaError_t e = cudaMalloc(&tab, (size_t)Hp * Wp * (hd / 2) * sizeof(float2));
    if (e != cudaSuccess) return op_fail(NDIT_ERR_NOMEM, "ndit_op_ln_rope", e);
    e = rope_table(tab, Hp, Wp, hd, theta, linear_factor, one_d, s);
    if (e == cudaSuccess)
        e = ln_rope_qk(static_cast<bf16*>(qkv), (H + 2 * Hkv) * hd, static_cast<const bf16*>(qw), static_cast<const bf16*>(qb),
                       static_cast<const bf16*>(kw), static_cast<const bf16*>(kb), tab, batch * Hp * Wp, Hp * Wp, H, Hkv, hd, s);
    cudaError_t e2 = cudaStreamSynchronize(s);
    cudaFree(tab);
    if (e == cudaSuccess) e = e2;
    return e == cudaSuccess ? NDIT_OK : op_fail(NDIT_ERR_CUDA, "ndit_op_ln_rope", e);
}

static int op_attention_impl(const void* qkv_, const void* kvy_, const uint8_t* ymask, const float* gate_tanh, void* out,
                             int32_t B, int32_t N, int32_t T, int32_t H, int32_t Hkv, float scale_self, float scale_cross,
                             int32_t use_ref, void* stream, int bench_iters, float* bench_ms, int hd) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (hd != 72 && hd != 48 && hd != 96) return op_fail(NDIT_ERR_INVALID, "ndit_op_attention: head_dim must be 72, 48 or 96", cudaErrorInvalidValue);
    const int vrows = attn_vrows(hd);
    const bf16* qkv = static_cast<const bf16*>(qkv_);
    const bf16* kvy = static_cast<const bf16*>(kvy_);
    const int Wq = (H + 2 * Hkv) * hd, KV = Hkv * hd;
    if (use_ref == 1) {
        cudaError_t e = attention_ref(qkv, Wq, kvy, 2 * KV, ymask, gate_tanh, static_cast<bf16*>(out), B, N, T, H, Hkv, hd,
                                      scale_self, scale_cross, s);
        return e == cudaSuccess ? NDIT_OK : op_fail(NDIT_ERR_CUDA, "ndit_op_attention(ref)", e);
    }
    const int Tpad = (T + 7) / 8 * 8, Npad = (N + 7) / 8 * 8;
    bf16 *vt = nullptr, *vyt = nullptr;
    const size_t vt_elems = (size_t)B * Hkv * vrows * Npad, vyt_elems = (size_t)B * Hkv * vrows * Tpad;
    cudaError_t e = cudaMalloc(&vt, vt_elems * 2);
    if (e == cudaSuccess) e = cudaMalloc(&vyt, vyt_elems * 2 + 256);
    if (e != cudaSuccess) return op_fail(NDIT_ERR_NOMEM, "ndit_op_attention", e);
    cudaMemsetAsync(vt, 0, vt_elems * 2, s);
    cudaMemsetAsync(vyt, 0, vyt_elems * 2, s);
    e = transpose_v(qkv, Wq, (H + Hkv) * hd, 0, vt, Npad, 0, B, N, Hkv, hd, vrows, 1, s);
    if (e == cudaSuccess && T > 0) e = transpose_v(kvy, 2 * KV, KV, 0, vyt, Tpad, 0, B, T, Hkv, hd, vrows, 1, s);
    if (e == cudaSuccess) e = fill_ones_row(vt, Npad, 0, B * Hkv, N, hd, vrows, 1, s);
    if (e == cudaSuccess && T > 0) e = fill_ones_row(vyt, Tpad, 0, B * Hkv, Tpad, hd, vrows, 1, s);
    AttnPlan a;
    memset(&a, 0, sizeof(a));
    // use_ref: 0 default kernel (NDIT_ATTN_GEN overrides), 1 CUDA-core reference, 2 half-row / tensor-memory-P kernel
    // (attention_hr_tcgen05.cu), 3 first-generation kernel (attention_tcgen05.cu)
    static const int gen_env = getenv("NDIT_ATTN_GEN") ? atoi(getenv("NDIT_ATTN_GEN")) : 0;
    const int gen = use_ref == 2 ? 3 : (use_ref == 3 ? 1 : attn_gen(gen_env));
    const int te = build_attn_maps(&a, qkv, Wq, vt, kvy, vyt, B, N, T, H, Hkv, hd, gen == 3 ? attention_hr_bkv(hd) : 128);
    auto run = [&](const AttnPlan& pl) { return gen == 3 ? attention_fused_hr(pl, s) : attention_fused(pl, s); };
    int rc = NDIT_OK;
    if (te) rc = op_fail(NDIT_ERR_CUDA, "ndit_op_attention tensor map", cudaSuccess);
    if (!te && e == cudaSuccess) {
        a.ymask = ymask; a.gate_tanh = gate_tanh; a.out = static_cast<bf16*>(out);
        a.scale_self = scale_self; a.scale_cross = scale_cross;
        e = run(a);
        if (bench_iters > 0 && bench_ms && e == cudaSuccess) {      // micro-benchmark: average of `bench_iters` launches
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = run(a);
            cudaEventRecord(e0, s);
            for (int i = 0; i < bench_iters && e == cudaSuccess; ++i) e = run(a);
            cudaEventRecord(e1, s);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            *bench_ms = ms / bench_iters;
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
        }
    }
    cudaError_t e2 = cudaStreamSynchronize(s);
    cudaFree(vt);
    cudaFree(vyt);
    if (e == cudaSuccess) e = e2;
    if (rc == NDIT_OK && e != cudaSuccess) rc = op_fail(NDIT_ERR_CUDA, "ndit_op_attention", e);
    return rc;
}

extern "C" int ndit_op_attention(const void* qkv_, const void* kvy_, const uint8_t* ymask, const float* gate_tanh, void* out,
                                 int32_t B, int32_t N, int32_t T, int32_t H, int32_t Hkv, float scale_self, float scale_cross,
                                 int32_t use_ref, void* stream) {
    return op_attention_impl(qkv_, kvy_, ymask, gate_tanh, out, B, N, T, H, Hkv, scale_self, scale_cross, use_ref, stream, 0, nullptr, 72);
}

extern "C" int ndit_op_attention_hd(const void* qkv_, const void* kvy_, const uint8_t* ymask, const float* gate_tanh, void* out,
                                    int32_t B, int32_t N, int32_t T, int32_t H, int32_t Hkv, int32_t hd, float scale_self,
                                    float scale_cross, int32_t use_ref, void* stream) {
    return op_attention_impl(qkv_, kvy_, ymask, gate_tanh, out, B, N, T, H, Hkv, scale_self, scale_cross, use_ref, stream, 0, nullptr, hd);
}

extern "C" int ndit_op_attention_bench(const void* qkv_, const void* kvy_, const uint8_t* ymask, const float* gate_tanh, void* out,
                                       int32_t B, int32_t N, int32_t T, int32_t H, int32_t Hkv, float scale_self, float scale_cross,
                                       int32_t iters, float* ms_out, void* stream) {
    if (iters <= 0 || !ms_out) return NDIT_ERR_INVALID;
    return op_attention_impl(qkv_, kvy_, ymask, gate_tanh, out, B, N, T, H, Hkv, scale_self, scale_cross, 0, stream, iters, ms_out, 72);
}

extern "C" int ndit_op_moe_gate(const void* u, const void* Wg, void* wtok, int32_t M, int32_t D, int32_t E, void* stream) {
    cudaError_t e = moe_space_gate(static_cast<const bf16*>(u), static_cast<const bf16*>(Wg), static_cast<bf16*>(wtok), M, D, E,
                                   static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? NDIT_OK : op_fail(NDIT_ERR_CUDA, "ndit_op_moe_gate", e);
}

extern "C" int ndit_op_resid_rms_mod(void* X, const void* o, const void* w_post, const void* tanh_g, const void* w_pre,
                                     const void* onepls, const void* shift, void* u, int32_t M, int32_t rows_per_batch, int32_t D,
                                     float eps, void* stream) {
    cudaError_t e = resid_rms_mod(static_cast<bf16*>(X), static_cast<const bf16*>(o), static_cast<const bf16*>(w_post),
                                  static_cast<const bf16*>(tanh_g), static_cast<const bf16*>(w_pre),
                                  static_cast<const bf16*>(onepls), static_cast<const bf16*>(shift), static_cast<bf16*>(u), M,
                                  rows_per_batch, D, D, eps,
                                  static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? NDIT_OK : op_fail(NDIT_ERR_CUDA, "ndit_op_resid_rms_mod", e);
}
