// Fused image self-attention (+ optional gated caption cross-attention) for one DiT block; head_dim 72, 48 or 96.
//
// Replaces, in lumina_next_t2i/models/model.py: _upad_input/flash_attn_varlen_func/pad_input (:387-404),
// the GQA repeat + masked SDPA over the caption tokens (:421-432), the tanh(gate) scale and add (:433-434);
// and the dense flash_attn_func of the class-conditional Next-DiT (Next-DiT-ImageNet/models/models.py:389), T = 0.
//
//   out[b,n,h,:] = bf16( bf16(softmax(q k^T * s_self) v) + bf16(tanh(gate_h) * bf16(softmax(q ky^T / sqrt(hd) + mask) vy)) )
//
// Design (sm_100a, one CTA per (batch, head, 256 query rows), 384 threads, 1 CTA / SM):
//   warp  0     TMA producer: Q tiles once, then a 3-stage ring of K / V^T tiles (self blocks, then caption blocks)
//   warps 1,2   MMA issuers (one per 128-row query tile):  S = Q K^T (tcgen05.mma 128x128x16; for head_dim 72: 4 k-steps
//               from a 128B-swizzled [rows x 64] tile + 1 from a 32B-swizzled [rows x 16] tile, the zero padding 72 -> 80
//               comes from TMA out-of-bounds fill; head_dim 48: 3 k-steps; head_dim 96: 4 + 2), then O += P V (128 x HDP x 16, 8 k-steps,
//               P from shared memory).  One issuer per tile with plain blocking waits, so the two tiles are free to run
//               half a softmax period apart (tile B is started late on purpose).
//   warps 4-7   softmax warpgroup for query tile A (rows 0..127): one thread per row (TMEM lane)
//   warps 8-11  softmax warpgroup for query tile B (rows 128..255)
// Registers are re-balanced with setmaxnreg (control warps 56, softmax warps 208: 128*56 + 256*208 must stay within
// the 384*168 the CTA was launched with - asking for more makes setmaxnreg.inc wait forever) so a softmax thread keeps
// its whole 128-column row of S in registers (one TMEM read pass) and S_x is handed back to the tensor core before the
// exponentials start: Q K^T of the next block overlaps the softmax of the current one.  O accumulates in TMEM across
// blocks; it is rescaled in place (tcgen05.ld / st) only when a row maximum grows by more than 2^8 (lazy rescale: a
// stale reference maximum is exact after the final division by the row sum).  V^T carries an all-ones row (index HD),
// so column HD of O is the softmax row sum, accumulated by the tensor core from the same bf16 P it multiplies with V.
// P is converted to bf16 by truncation (one PRMT per pair; cvt.rn shares the XU pipe with MUFU.EX2); the -2^-9 mean
// bias cancels in O / rowsum.
// Measured alternatives that were slower and are not kept (profiles/r01_attention_microbench_phase_timing.txt):
// polynomial exp2 on the FMA pipe for 1/8..3/8 of the columns, hand-pipelined MUFU/pack ordering with volatile asm,
// two softmax threads per row (640-thread CTA), an explicit XU token between the two tiles, a speculative (stale) reference
// maximum with the row maximum accumulated inside the exp loop (480 vs 421 us: the up-front max pass is what keeps the two
// tiles half a period apart), P kept in tensor memory (attention_tp_tcgen05.cu, 445-466 us), a second buffer for the first half
// of P in the dead Q tile so that the softmax never waits for P V before it starts writing (449 vs 409 us: same story - every
// change that lets a softmax warpgroup run ahead destroys the half-period offset between the two tiles), and enforcing that
// offset with "half of my exponentials are issued" barriers between the two warpgroups (440 us, 443 us with the extra buffer).
// Round 2, also measured on this kernel and not kept (tools/attn_bench.py, config-2 shape, 410-414 us baseline):
//   * tcgen05.mma / TMA issued under elect.sync instead of `lane == 0` (back-to-back UTCHMMA instead of a ~45-cycle ELECT / BRA.U.ANY
//     loop per instruction, see ptx.cuh): 464 us - the MMAs of a tile complete earlier, the two softmax warpgroups drift out of
//     their half-period offset (third-generation kernel attention_hr_tcgen05.cu keeps elect.sync, there it is worth 90 us);
//   * software-pipelining the exponentials by basic block (packs / stores of chunk c-1 and the FFMAs of chunk c+1 in the block of
//     chunk c's 32 MUFU.EX2, behind always-true branches ptxas cannot fold - it schedules inside basic blocks only, and left alone
//     puts every PRMT 17 cycles behind its own MUFU pair): the SASS then issues one MUFU every 8 cycles with everything else in
//     the gaps, but the kernel takes 479 us - same mechanism, the exp phase of one tile now overlaps the other tile's;
//   * making the PRMT selector depend on the chunk's last exponential: ptxas simply computes that exponential first;
//   * the two tiles taking strict turns on the XU pipe (mbarrier token), alone and combined with the two changes above:
//     460 us (token only), 444 (+ basic-block pipelining), 446 (+ elect.sync) against 423 us on the same box: with the exponentials
//     of a tile down to ~1070 cycles per block the other tile is NOT ready when the token arrives - its own chain of barrier
//     waits, TMEM loads, row maximum and P hand-over takes ~2200 cycles per block, and that chain, not the exp loop, is the bound
//     (profiles/r02_attention_third_generation_log.txt).
// Q is a TENSOR-MEMORY operand (head_dim 72 / 48): copied once per CTA from its TMA tile into free TMEM columns, so Q K^T
// reads only K from shared memory (416.7 -> 412.9 us on the config-2 shape; -14 % shared-memory traffic).
// TMEM columns: S_A [0,128) S_B [128,256) O_A [256,256+HDP) Q_A [256+HDP, ..+40) O_B [384,384+HDP) Q_B [384+HDP, ..+40).
#include <math.h>
#include <stdlib.h>

#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace ndit {

constexpr int AT_BQ = 128;            // rows per query tile
constexpr int AT_BKV = 128;           // kv rows per block
constexpr int AT_THREADS = 384;       // warps 0-3 control (TMA, MMA tile A, MMA tile B, idle), 4-7 softmax A, 8-11 softmax B

constexpr int AT_Q64_BYTES = AT_BQ * 128;        // 16 KB  [rows x 64] bf16, 128B swizzle
constexpr int AT_Q16_BYTES = AT_BQ * 32;         //  4 KB  [rows x 16] bf16, 32B swizzle (head_dim 72: one, head_dim 96: two)
constexpr int AT_PHALF_BYTES = AT_BQ * 128;      // 16 KB
constexpr int AT_PTILE_BYTES = 2 * AT_PHALF_BYTES;

constexpr uint32_t AT_TM_S = 0, AT_TM_O = 256;   // + X*128
constexpr float AT_RESCALE_LOG2 = 8.0f;          // lazy rescale: keep a stale reference max while exp2 args stay <= 8

template <int HD>
struct AttnDims {
    static_assert(HD == 72 || HD == 48 || HD == 96,
                  "head_dim 72 (Lumina-Next-T2I 2B / Next-DiT 2B), 48 (Next-DiT 600M) or 96 (Flag-DiT 5B)");
    static constexpr int NK64 = HD >= 64 ? 4 : HD / 16;     // Q K^T k-steps from the 64-wide chunk
    static constexpr int N16 = HD > 64 ? (HD - 64 + 15) / 16 : 0;   // + k-steps from 16-wide chunks (elements 64.., zero padded)
    static constexpr int HDP = attn_vrows(HD);              // P V width = V^T rows: head_dim + ones row, padded to 16
    static constexpr int STAGES = HD > 80 ? 2 : 3;          // K / V^T ring depth (shared-memory budget)
    // shared-memory layout
    static constexpr int QTILE_BYTES = AT_Q64_BYTES + N16 * AT_Q16_BYTES;
    static constexpr int KTILE_BYTES = QTILE_BYTES;
    static constexpr int VHALF_BYTES = HDP * 128;           // HDP V^T rows x 64 kv
    static constexpr int VTILE_BYTES = 2 * VHALF_BYTES;
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = OFF_Q + 2 * QTILE_BYTES;
    static constexpr int OFF_V = OFF_K + STAGES * KTILE_BYTES;
    static constexpr int OFF_P = OFF_V + STAGES * VTILE_BYTES;
    static constexpr int OFF_BAR = OFF_P + 2 * AT_PTILE_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
    static_assert(OFF_K % 1024 == 0 && OFF_V % 1024 == 0 && OFF_P % 1024 == 0 && VHALF_BYTES % 1024 == 0, "swizzle atoms need 1 KB alignment");
    static_assert(SMEM_BYTES <= 227 * 1024, "attention smem budget");
    static constexpr int QK_TX = QTILE_BYTES;
    static constexpr int V_TX = VTILE_BYTES;
    // Q as a TENSOR-MEMORY operand: Q never changes during the KV loop, so after its TMA load each softmax thread copies
    // its row (packed bf16 pairs, one 32-bit column per pair) into the free TMEM columns next to O_x, and S = Q K^T is issued
    // with the A operand in TMEM.  The tensor core then fetches only K from shared memory for Q K^T: -40 KB of the 288 KB
    // of shared-memory traffic per pair of blocks.  Needs (64 + 16 N16) / 2 free columns per tile: head_dim 72 and 48.
#ifdef AT_NO_Q_TMEM
    static constexpr bool Q_TMEM = false;
#else
    static constexpr bool Q_TMEM = HDP + (64 + 16 * N16) / 2 <= 128;
#endif
    static constexpr uint32_t TM_Q = 256 + HDP;             // + x * 128 (inside the 128-column slot of O_x)
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16_trunc(float lo, float hi) {
    return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632);
}

#ifdef AT_TIMING
__device__ long long g_at_timing[2][64][8];
#define AT_STAMP(k)                                                                                     \
    do {                                                                                                \
        if (blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 0 && qd == 0 && lane == 0 && jj < 64)  \
            g_at_timing[x][jj][k] = clock64();                                                          \
    } while (0)
#else
#define AT_STAMP(k)
#endif

// REGION (compositional generation, lumina_next_compositional_generation/models/model.py:421-446, 872-887): the caption segment holds
// n_cond region captions followed by the unconditional one.  Batch row 0 (cond) attends, per query token, to the ONE caption whose region
// contains the token (region_id = (h_split + 1) * (w_split + 1) - 1 of its rectangle, :879; the reference sums the per-caption SDPA
// outputs, of which at most one is non-zero for a token, every other caption being fully masked -> NaN -> nan_to_num -> 0); batch row 1
// (uncond) attends to the last caption from every token (region_mask[-1] = 1, :885).  All captions of a batch row run as ONE softmax
// segment over the concatenated keys with a per-row mask, which is the same thing because exactly one caption survives the mask.
template <int HD, bool REGION>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_fused_kernel(const __grid_constant__ CUtensorMap tmQ64, const __grid_constant__ CUtensorMap tmQ16,
                       const __grid_constant__ CUtensorMap tmK64, const __grid_constant__ CUtensorMap tmK16,
                       const __grid_constant__ CUtensorMap tmVt, const __grid_constant__ CUtensorMap tmKy64,
                       const __grid_constant__ CUtensorMap tmKy16, const __grid_constant__ CUtensorMap tmVyt,
                       const uint8_t* __restrict__ ymask, const float* __restrict__ gate_tanh, bf16* __restrict__ out,
                       int N, int T, int H, int Hkv, float sl2_self, float sl2_cross, const int* __restrict__ kv_len, const AttnRegion reg) {
    using Dm = AttnDims<HD>;
    constexpr int HDP = Dm::HDP;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = sbase + Dm::OFF_BAR;
    auto q_full = [&](int x) { return bar0 + 8u * (0 + x); };
    auto k_full = [&](int s) { return bar0 + 8u * (2 + s); };
    auto v_full = [&](int s) { return bar0 + 8u * (2 + Dm::STAGES + s); };
    auto kv_empty = [&](int s) { return bar0 + 8u * (2 + 2 * Dm::STAGES + s); };
    constexpr int BB = 2 + 3 * Dm::STAGES;
    auto s_full = [&](int x) { return bar0 + 8u * (BB + 0 + x); };   // S_x = Q K^T landed in TMEM
    auto s_free = [&](int x) { return bar0 + 8u * (BB + 2 + x); };   // softmax x holds S_x in registers
    auto p_full = [&](int x) { return bar0 + 8u * (BB + 4 + x); };   // P_x in smem (and O_x rescaled if needed)
    auto o_full = [&](int x) { return bar0 + 8u * (BB + 6 + x); };   // O_x += P_x V finished (P_x smem reusable)
    const uint32_t stagger_bar = bar0 + 8u * (BB + 8);                // tile B starts ~half a softmax period after tile A
    auto q_ready = [&](int x) { return bar0 + 8u * (BB + 9 + x); };  // Q_x copied to tensor memory
    const uint32_t tmem_ptr_addr = bar0 + 8u * (BB + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (2 * AT_BQ);
    const int h = blockIdx.y, b = blockIdx.z;
    const int g = h / (H / Hkv);
    // variable-resolution list input: this batch row has only Nv valid image tokens (keys beyond are masked; the padded query
    // rows are computed like any other and dropped by the caller).  Read before pdl_wait: written by a host copy, not a kernel.
    const int Nv = kv_len != nullptr ? kv_len[b] : N;
    const int n_self = (Nv + AT_BKV - 1) / AT_BKV;
    const int n_tb = (T + AT_BKV - 1) / AT_BKV;          // kv blocks per caption; 0 for the class-conditional model
    // REGION: captions [cap0, cap0 + n_caps) belong to this batch row (cond: the region captions, uncond: the last one)
    const int cap0 = REGION ? (b == 0 ? 0 : reg.n_cond) : b;
    const int n_caps = REGION ? (b == 0 ? reg.n_cond : 1) : 1;
    const int n_cross = n_tb * n_caps;
    const int n_total = n_self + n_cross;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ64); tma_prefetch_desc(&tmK64); tma_prefetch_desc(&tmVt);
        if (Dm::N16 > 0) { tma_prefetch_desc(&tmQ16); tma_prefetch_desc(&tmK16); }
        if (n_cross > 0) { tma_prefetch_desc(&tmKy64); tma_prefetch_desc(&tmVyt); if (Dm::N16 > 0) tma_prefetch_desc(&tmKy16); }
        for (int x = 0; x < 2; ++x) {
            mbar_init(q_full(x), 1);
            mbar_init(s_full(x), 1);
            mbar_init(s_free(x), 4);
            mbar_init(p_full(x), 4);
            mbar_init(o_full(x), 1);
        }
        for (int s = 0; s < Dm::STAGES; ++s) {
            mbar_init(k_full(s), 1);
            mbar_init(v_full(s), 1);
            mbar_init(kv_empty(s), 2);
        }
        mbar_init(stagger_bar, 4);
        mbar_init(q_ready(0), 4);
        mbar_init(q_ready(1), 4);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_addr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
    pdl_trigger();
    pdl_wait();         // q/k/v (and the output buffer) belong to the previous kernels until they have completed

    if (warp < 4) {
        setmaxnreg_dec<56>();
        if (warp == 0) {
            // ================================================================= TMA producer
            if (lane == 0) {
                for (int x = 0; x < 2; ++x) {
                    const uint32_t dst = sbase + Dm::OFF_Q + x * Dm::QTILE_BYTES;
                    mbar_expect_tx(q_full(x), Dm::QK_TX);
                    tma_load_3d(dst, &tmQ64, q_full(x), 0, h, b * N + q0 + x * AT_BQ);
#pragma unroll
                    for (int c = 0; c < Dm::N16; ++c)
                        tma_load_3d(dst + AT_Q64_BYTES + c * AT_Q16_BYTES, &tmQ16, q_full(x), 64 + 16 * c, h, b * N + q0 + x * AT_BQ);
                }
            }
            __syncwarp();
            int s = 0;
            uint32_t ph = 0;
            for (int jj = 0; jj < n_total; ++jj) {
                mbar_wait(kv_empty(s), ph ^ 1);
                if (lane == 0) {
                    const uint32_t kd = sbase + Dm::OFF_K + s * Dm::KTILE_BYTES;
                    const uint32_t vd = sbase + Dm::OFF_V + s * Dm::VTILE_BYTES;
                    mbar_expect_tx(k_full(s), Dm::QK_TX);
                    mbar_expect_tx(v_full(s), Dm::V_TX);
                    if (jj < n_self) {
                        const int kv0 = jj * AT_BKV;
                        tma_load_3d(kd, &tmK64, k_full(s), 0, g, b * N + kv0);
#pragma unroll
                        for (int c = 0; c < Dm::N16; ++c)
                            tma_load_3d(kd + AT_Q64_BYTES + c * AT_Q16_BYTES, &tmK16, k_full(s), 64 + 16 * c, g, b * N + kv0);
                        tma_load_3d(vd, &tmVt, v_full(s), kv0, 0, b * Hkv + g);
                        tma_load_3d(vd + Dm::VHALF_BYTES, &tmVt, v_full(s), kv0 + 64, 0, b * Hkv + g);
                    } else {
                        const int jc = jj - n_self;
                        const int cap = REGION ? cap0 + jc / n_tb : b;                  // caption row of this block
                        const int kv0 = (REGION ? jc % n_tb : jc) * AT_BKV;
                        tma_load_3d(kd, &tmKy64, k_full(s), 0, g, cap * T + kv0);
#pragma unroll
                        for (int c = 0; c < Dm::N16; ++c)
                            tma_load_3d(kd + AT_Q64_BYTES + c * AT_Q16_BYTES, &tmKy16, k_full(s), 64 + 16 * c, g, cap * T + kv0);
                        tma_load_3d(vd, &tmVyt, v_full(s), kv0, 0, cap * Hkv + g);
                        tma_load_3d(vd + Dm::VHALF_BYTES, &tmVyt, v_full(s), kv0 + 64, 0, cap * Hkv + g);
                    }
                }
                __syncwarp();
                if (++s == Dm::STAGES) { s = 0; ph ^= 1; }
            }
        } else if (warp == 1 || warp == 2) {
            // ================================================================= MMA issuers: warp 1 -> tile A, warp 2 -> tile B
            const int x = warp - 1;
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, AT_BKV);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, HDP);
            auto issue_qk = [&](int jj) {
                const int s = jj % Dm::STAGES;
                mbar_wait(k_full(s), (jj / Dm::STAGES) & 1);
                mbar_wait(s_free(x), (jj & 1) ^ 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t qa = sbase + Dm::OFF_Q + x * Dm::QTILE_BYTES;
                    const uint32_t ka = sbase + Dm::OFF_K + s * Dm::KTILE_BYTES;
                    const uint64_t dq = make_smem_desc_kmajor(qa, 1024, UMMA_SW128);
                    const uint64_t dk = make_smem_desc_kmajor(ka, 1024, UMMA_SW128);
                    const uint32_t d = tmem_base + AT_TM_S + x * 128;
                    const uint32_t tq = tmem_base + Dm::TM_Q + x * 128;
#pragma unroll
                    for (int k = 0; k < Dm::NK64; ++k) {
                        if (Dm::Q_TMEM) umma_ts(d, tq + 8 * k, dk + 2 * k, idesc_qk, k != 0);
                        else umma_ss(d, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
                    }
#pragma unroll
                    for (int c = 0; c < Dm::N16; ++c) {
                        const uint64_t dq16 = make_smem_desc_kmajor(qa + AT_Q64_BYTES + c * AT_Q16_BYTES, 256, UMMA_SW32);
                        const uint64_t dk16 = make_smem_desc_kmajor(ka + AT_Q64_BYTES + c * AT_Q16_BYTES, 256, UMMA_SW32);
                        if (Dm::Q_TMEM) umma_ts(d, tq + 32 + 8 * c, dk16, idesc_qk, 1);
                        else umma_ss(d, dq16, dk16, idesc_qk, 1);
                    }
                    umma_commit(s_full(x));
                }
                __syncwarp();
            };
            auto issue_pv = [&](int jj) {
                const int s = jj % Dm::STAGES;
                const uint32_t acc0 = (jj != 0 && jj != n_self) ? 1u : 0u;   // new softmax segment -> fresh accumulator
                mbar_wait(v_full(s), (jj / Dm::STAGES) & 1);
                mbar_wait(p_full(x), jj & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t pa = sbase + Dm::OFF_P + x * AT_PTILE_BYTES;
                    const uint32_t va = sbase + Dm::OFF_V + s * Dm::VTILE_BYTES;
                    const uint32_t d = tmem_base + AT_TM_O + x * 128;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint64_t dp = make_smem_desc_kmajor(pa + (k >> 2) * AT_PHALF_BYTES, 1024, UMMA_SW128) + 2 * (k & 3);
                        const uint64_t dv = make_smem_desc_kmajor(va + (k >> 2) * Dm::VHALF_BYTES, 1024, UMMA_SW128) + 2 * (k & 3);
                        umma_ss(d, dp, dv, idesc_pv, acc0 | (k != 0));
                    }
                    umma_commit(o_full(x));
                    umma_commit(kv_empty(s));                   // second arrival (other tile) releases the K / V^T stage
                }
                __syncwarp();
            };
            mbar_wait(Dm::Q_TMEM ? q_ready(x) : q_full(x), 0);
            issue_qk(0);
            for (int jj = 0; jj < n_total; ++jj) {
                // S_x(jj+1) is issued as soon as softmax x holds S_x(jj) in registers, well ahead of P_x(jj) V
                if (jj + 1 < n_total) issue_qk(jj + 1);
                issue_pv(jj);
            }
        }
    } else {
        // ================================================================= softmax warpgroups (one thread per row)
        setmaxnreg_inc<208>();
        const int x = (warp >> 2) - 1;           // query tile
        const int qd = warp & 3;                 // TMEM lane quadrant
        const int r = qd * 32 + lane;            // row inside the tile
        const int qrow = q0 + x * AT_BQ + r;     // token index in this batch element
        const uint32_t lane_sel = static_cast<uint32_t>(qd * 32) << 16;
        const uint32_t ts = tmem_base + lane_sel + AT_TM_S + x * 128;
        const uint32_t to = tmem_base + lane_sel + AT_TM_O + x * 128;
        const uint32_t pbase = sbase + Dm::OFF_P + x * AT_PTILE_BYTES + r * 128;
        const uint32_t rsw = static_cast<uint32_t>(r & 7);

        if (Dm::Q_TMEM) {
            // ---- Q_x row r: shared memory (TMA swizzle layouts) -> tensor memory columns TM_Q .. (lane = row)
            const uint32_t tq = tmem_base + lane_sel + Dm::TM_Q + x * 128;
            const uint32_t qs = sbase + Dm::OFF_Q + x * Dm::QTILE_BYTES;
            mbar_wait(q_full(x), 0);
#pragma unroll
            for (int j = 0; j < 8; j += 2) {                 // 64-wide chunk: 16-byte piece j sits at (j ^ (r & 7)) (128B swizzle)
                uint32_t v[8];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const uint32_t addr = qs + r * 128 + ((static_cast<uint32_t>(j + u) ^ rsw) << 4);
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(v[4 * u]), "=r"(v[4 * u + 1]), "=r"(v[4 * u + 2]), "=r"(v[4 * u + 3]) : "r"(addr));
                }
                tmem_st_32x32b_x8(tq + 4 * j, v);
            }
#pragma unroll
            for (int c = 0; c < Dm::N16; ++c) {              // 16-wide chunks: piece j at (j ^ ((r >> 2) & 1)) (32B swizzle)
                uint32_t v[8];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const uint32_t addr = qs + AT_Q64_BYTES + c * AT_Q16_BYTES + r * 32 + ((static_cast<uint32_t>(u) ^ ((r >> 2) & 1u)) << 4);
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(v[4 * u]), "=r"(v[4 * u + 1]), "=r"(v[4 * u + 2]), "=r"(v[4 * u + 3]) : "r"(addr));
                }
                tmem_st_32x32b_x8(tq + 32 + 8 * c, v);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_ready(x));
        }

        uint32_t o_self[HD / 2];
#pragma unroll
        for (int i = 0; i < HD / 2; ++i) o_self[i] = 0u;
        float m_ref = -INFINITY;
        // REGION: the one caption row this query token attends to (-1: none - a token outside every rectangle, or whose region id
        // is not a cond caption: the reference then adds nan_to_num(NaN) = 0)
        int my_cap = b;
        if constexpr (REGION) {
            if (b == 0) {
                const int hs = (qrow / reg.Wp) / reg.hp, ws = (qrow % reg.Wp) / reg.wp;
                const int rid = (hs + 1) * (ws + 1) - 1;
                my_cap = (hs < reg.hs && ws < reg.ws && rid < reg.n_cond) ? rid : -1;
            } else {
                my_cap = reg.n_cond;
            }
        }

        // O_x[:, 0:HD] / rowsum (column HD) -> packed bf16.  combine: dst = bf16(dst + bf16(gt * bf16(value)))
        auto read_o = [&](uint32_t* dst, bool combine, float gt) {
            uint32_t l8[8];
            tmem_ld_32x32b_x8(to + HD, l8);       // HD is a multiple of 8
            tmem_ld_wait();
            float inv = 1.0f / __uint_as_float(l8[0]);
            if constexpr (REGION) { if (__uint_as_float(l8[0]) == 0.f) inv = 0.f; }   // no valid key: nan_to_num(softmax of all -inf) = 0
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {
                uint32_t v[8];
                tmem_ld_32x32b_x8(to + c * 8, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float a = __uint_as_float(v[2 * i]) * inv, cc = __uint_as_float(v[2 * i + 1]) * inv;
                    if (!combine) dst[c * 4 + i] = pack_bf16(a, cc);
                    else {
                        const float2 sv = unpack_bf16(dst[c * 4 + i]);
                        dst[c * 4 + i] = pack_bf16(sv.x + bf16_round(gt * bf16_round(a)), sv.y + bf16_round(gt * bf16_round(cc)));
                    }
                }
            }
        };

#ifndef AT_NO_STAGGER
        if (x == 1) mbar_wait(stagger_bar, 0);   // anti-phase the two tiles: one is in its exp phase while the other is not
#endif
        for (int jj = 0; jj < n_total; ++jj) {
            const bool cross = jj >= n_self;
            const bool first = (jj == 0) || (jj == n_self);
            if (jj == n_self) {
                // ---- segment switch: O_x holds the complete self-attention numerator -> bf16(O / l) into registers
                mbar_wait(o_full(x), (jj - 1) & 1);
                tc_fence_after();
                read_o(o_self, false, 0.f);
                m_ref = -INFINITY;
            }
            const float sl2 = cross ? sl2_cross : sl2_self;
            // validity words for this block's 128 kv columns
            uint32_t vw[4];
            if (!cross) {
                const int nvalid = Nv - jj * AT_BKV;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int rem = nvalid - c * 32;
                    vw[c] = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
                }
            } else {
                const int jc = jj - n_self;
                const int cap = REGION ? cap0 + jc / n_tb : b;
                const int t0 = (REGION ? jc % n_tb : jc) * AT_BKV;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int t = t0 + c * 32 + lane;
                    const bool ok = (t < T) && (ymask[cap * T + t] != 0);
                    vw[c] = __ballot_sync(0xffffffffu, ok);
                    if constexpr (REGION) { if (cap != my_cap) vw[c] = 0u; }     // per row: another region's caption
                }
            }
            const bool all_valid = (vw[0] & vw[1] & vw[2] & vw[3]) == 0xffffffffu;

            // ---- the whole 128-column row of S into registers, then hand S_x back to the tensor core
            AT_STAMP(0);
            mbar_wait(s_full(x), jj & 1);
            tc_fence_after();
            AT_STAMP(1);
            uint32_t sreg[128];
            tmem_ld_32x32b_x32(ts, sreg);
            tmem_ld_32x32b_x32(ts + 32, sreg + 32);
            tmem_ld_32x32b_x32(ts + 64, sreg + 64);
            tmem_ld_32x32b_x32(ts + 96, sreg + 96);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free(x));
            AT_STAMP(2);

            // ---- row max
            float mb;
            if (all_valid) {
                float m0 = __uint_as_float(sreg[0]), m1 = __uint_as_float(sreg[1]), m2 = __uint_as_float(sreg[2]), m3 = __uint_as_float(sreg[3]);
#pragma unroll
                for (int i = 4; i < 128; i += 4) {
                    m0 = fmaxf(m0, __uint_as_float(sreg[i]));
                    m1 = fmaxf(m1, __uint_as_float(sreg[i + 1]));
                    m2 = fmaxf(m2, __uint_as_float(sreg[i + 2]));
                    m3 = fmaxf(m3, __uint_as_float(sreg[i + 3]));
                }
                mb = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            } else {
                mb = -INFINITY;
#pragma unroll
                for (int i = 0; i < 128; ++i) {
                    if (!((vw[i >> 5] >> (i & 31)) & 1u)) sreg[i] = 0xff800000u;   // -inf
                    mb = fmaxf(mb, __uint_as_float(sreg[i]));
                }
            }
            const float m_new = fmaxf(m_ref, mb);
            // REGION, caption segment: a row may meet its first valid key in any block of the segment (its caption need not be the
            // first one); until then m_ref stays -inf ("unseen"), O_x of the row is all zeros and needs no rescale
            const bool unseen = REGION && cross && m_ref == -INFINITY;
            if (first && !(REGION && cross)) {
                m_ref = (m_new == -INFINITY) ? 0.f : m_new;
            } else {
                // lazy rescale (warp-uniform decision because tcgen05.ld/st are warp-collective)
                const bool need = !unseen && (m_new - m_ref) * sl2 > AT_RESCALE_LOG2;
                if (!first && __any_sync(0xffffffffu, need)) {
                    const float alpha = unseen ? 1.0f : ex2_approx((m_ref - m_new) * sl2);   // scales O and the row sum (column HD)
                    mbar_wait(o_full(x), (jj - 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int c = 0; c < HDP / 16; ++c) {
                        uint32_t v[16];
                        tmem_ld_32x32b_x16(to + c * 16, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                        tmem_st_32x32b_x16(to + c * 16, v);
                    }
                    tmem_st_wait();
                    m_ref = m_new;
                }
                if (unseen) m_ref = m_new;      // still -inf if this block held no valid key for the row either
            }
            const float moff = (REGION && m_ref == -INFINITY) ? 0.f : m_ref * sl2;
            AT_STAMP(3);
            if (jj == 0 && x == 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(stagger_bar);
            }
            if (jj > 0) mbar_wait(o_full(x), (jj - 1) & 1);   // P_x smem reusable: P_x(jj-1) V has completed
            AT_STAMP(4);
            // ---- p = exp2(s*sl2 - moff) -> P_x in smem (bf16, K-major 128B swizzle).  Masked entries hold -inf -> 0.
            // The row sum is not accumulated here (all-ones row of V^T, see the header).
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float pe[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) pe[i] = ex2_approx(fmaf(__uint_as_float(sreg[c * 32 + i]), sl2, -moff));
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const uint32_t addr = pbase + (c >> 1) * AT_PHALF_BYTES + ((static_cast<uint32_t>((c & 1) * 4 + ch) ^ rsw) << 4);
                    const uint32_t a0 = pack_bf16_trunc(pe[ch * 8 + 0], pe[ch * 8 + 1]), a1 = pack_bf16_trunc(pe[ch * 8 + 2], pe[ch * 8 + 3]);
                    const uint32_t a2 = pack_bf16_trunc(pe[ch * 8 + 4], pe[ch * 8 + 5]), a3 = pack_bf16_trunc(pe[ch * 8 + 6], pe[ch * 8 + 7]);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a0), "r"(a1), "r"(a2), "r"(a3) : "memory");
                }
            }
            AT_STAMP(5);
            // P_x written (and O_x rescaled) -> visible to the tensor core
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full(x));
            AT_STAMP(6);
        }
        // ---- epilogue: out = bf16(self + bf16(tanh(gate) * bf16(cross)))   (no caption segment: out = bf16(self))
        mbar_wait(o_full(x), (n_total - 1) & 1);
        tc_fence_after();
        if (n_cross > 0) read_o(o_self, true, gate_tanh[h]);
        else read_o(o_self, false, 0.f);
        if (qrow < N) {
            bf16* dst = out + (static_cast<size_t>(b) * N + qrow) * (static_cast<size_t>(H) * HD) + h * HD;
#pragma unroll
            for (int i = 0; i < HD / 8; ++i)
                *reinterpret_cast<uint4*>(dst + i * 8) = make_uint4(o_self[4 * i], o_self[4 * i + 1], o_self[4 * i + 2], o_self[4 * i + 3]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

#ifdef AT_TIMING
extern "C" int ndit_debug_attn_timing(long long* out) {   // [2][64][8] clock64 stamps of CTA (1,3,0), quadrant-0 lane 0
    return cudaMemcpyFromSymbol(out, g_at_timing, sizeof(g_at_timing)) == cudaSuccess ? 0 : -1;
}
#endif

template <int HD, bool REGION>
static cudaError_t launch_attention(const AttnPlan& p, cudaStream_t stream) {
    auto kern = attention_fused_kernel<HD, REGION>;
    static PerDeviceFlag flags;
    bool& configured = flags.here();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnDims<HD>::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const float log2e = 1.4426950408889634f;
    const dim3 grid((p.N + 2 * AT_BQ - 1) / (2 * AT_BQ), p.H, p.B);
    return launch_k(kern, grid, dim3(AT_THREADS), AttnDims<HD>::SMEM_BYTES, stream, p.tmQ64, p.tmQ16, p.tmK64, p.tmK16, p.tmVt, p.tmKy64, p.tmKy16, p.tmVyt, p.ymask,
                                                      p.gate_tanh, p.out, p.N, p.T, p.H, p.Hkv, p.scale_self * log2e,
                                                      p.scale_cross * log2e, p.kv_len, p.region);
}

cudaError_t attention_fused(const AttnPlan& p, cudaStream_t stream) {
    if (p.T < 0 || p.N <= 0) return cudaErrorInvalidValue;
    if (p.region.n_cond > 0) {
        // region-masked cross-attention of the compositional model: one cond / uncond pair, text-conditioned Next-DiT (head_dim 72)
        const AttnRegion& r = p.region;
        if (p.B != 2 || p.T <= 0 || p.hd != 72 || p.kv_len != nullptr || r.Wp <= 0 || r.hp <= 0 || r.wp <= 0 || r.hs <= 0 || r.ws <= 0)
            return cudaErrorInvalidValue;
        return launch_attention<72, true>(p, stream);
    }
    if (p.hd == 72) return launch_attention<72, false>(p, stream);
    if (p.hd == 48) return launch_attention<48, false>(p, stream);
    if (p.hd == 96) return launch_attention<96, false>(p, stream);
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------
// Slow CUDA-core reference of the same fused op (debug only; selected with NDIT_ATTN=ref).
// One block per (query row, head, batch); plain two-pass softmax in fp32.
__global__ void attention_ref_kernel(const bf16* __restrict__ qkv, int ld_qkv, const bf16* __restrict__ kvy, int ld_kvy,
                                     const uint8_t* __restrict__ ymask, const float* __restrict__ gate_tanh,
                                     bf16* __restrict__ out, int N, int T, int H, int Hkv, int hd, float scale_self,
                                     float scale_cross, const int* __restrict__ kv_len, const AttnRegion reg) {
    extern __shared__ float sh[];            // scores [max(N,T)] + q [hd] + red[32]
    const int n = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    // caption row of this query token: its own batch row, or (region mode, see attention_fused_kernel) the caption of its region
    int cap = b;
    if (reg.n_cond > 0) {
        if (b == 0) {
            const int hs = (n / reg.Wp) / reg.hp, ws = (n % reg.Wp) / reg.wp;
            const int rid = (hs + 1) * (ws + 1) - 1;
            cap = (hs < reg.hs && ws < reg.ws && rid < reg.n_cond) ? rid : -1;
        } else {
            cap = reg.n_cond;
        }
    }
    const int g = h / (H / Hkv);
    const int L = N > T ? N : T;
    float* sc = sh;
    float* qv = sh + L;
    float* red = qv + hd;
    const bf16* qp = qkv + (static_cast<size_t>(b) * N + n) * ld_qkv + h * hd;
    for (int d = threadIdx.x; d < hd; d += blockDim.x) qv[d] = __bfloat162float(qp[d]);
    __syncthreads();
    float result[2] = {0.f, 0.f};            // this thread owns output dims threadIdx.x (< hd) for both segments
    for (int seg = 0; seg < 2; ++seg) {
        const int len = seg == 0 ? (kv_len != nullptr ? kv_len[b] : N) : T;
        if (len == 0 || (seg == 1 && cap < 0)) continue;   // no caption segment / no caption for this token (uniform for the whole block)
        const float scale = seg == 0 ? scale_self : scale_cross;
        float mx = -INFINITY;
        for (int k = threadIdx.x; k < len; k += blockDim.x) {
            const bf16* kp = seg == 0 ? qkv + (static_cast<size_t>(b) * N + k) * ld_qkv + H * hd + g * hd
                                      : kvy + (static_cast<size_t>(cap) * T + k) * ld_kvy + g * hd;
            float s = 0.f;
            for (int d = 0; d < hd; ++d) s += qv[d] * __bfloat162float(kp[d]);
            s *= scale;
            if (seg == 1 && ymask[cap * T + k] == 0) s = -INFINITY;
            sc[k] = s;
            mx = fmaxf(mx, s);
        }
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
        __syncthreads();
        mx = -INFINITY;
        for (int w = 0; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
        __syncthreads();
        float sum = 0.f;
        for (int k = threadIdx.x; k < len; k += blockDim.x) {
            const float p = __expf(sc[k] - mx);
            sum += p;
            sc[k] = bf16_round(p);           // flash kernels feed bf16 probabilities to the P.V product
        }
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
        __syncthreads();
        sum = 0.f;
        for (int w = 0; w < (blockDim.x >> 5); ++w) sum += red[w];
        __syncthreads();
        if (threadIdx.x < hd) {
            const int d = threadIdx.x;
            float acc = 0.f;
            for (int k = 0; k < len; ++k) {
                const bf16* vp = seg == 0 ? qkv + (static_cast<size_t>(b) * N + k) * ld_qkv + (H + Hkv) * hd + g * hd
                                          : kvy + (static_cast<size_t>(cap) * T + k) * ld_kvy + Hkv * hd + g * hd;
                acc += sc[k] * __bfloat162float(vp[d]);
            }
            result[seg] = sum > 0.f ? bf16_round(acc / sum) : (reg.n_cond > 0 ? 0.f : bf16_round(acc / sum));   // region mode: nan_to_num
        }
        __syncthreads();
    }
    if (threadIdx.x < hd) {
        const float r = T > 0 ? result[0] + bf16_round(gate_tanh[h] * result[1]) : result[0];
        out[(static_cast<size_t>(b) * N + n) * (static_cast<size_t>(H) * hd) + h * hd + threadIdx.x] = __float2bfloat16_rn(r);
    }
}

cudaError_t attention_ref(const bf16* qkv, int ld_qkv, const bf16* kvy, int ld_kvy, const uint8_t* ymask,
                          const float* gate_tanh, bf16* out, int B, int N, int T, int H, int Hkv, int hd,
                          float scale_self, float scale_cross, cudaStream_t stream, const int* kv_len, AttnRegion region) {
    if (hd > 128) return cudaErrorInvalidValue;
    if (region.n_cond > 0 && (B != 2 || T <= 0 || region.Wp <= 0 || region.hp <= 0 || region.wp <= 0)) return cudaErrorInvalidValue;
    const int L = N > T ? N : T;
    const size_t sh = (L + hd + 32) * sizeof(float);
    static size_t configured_sz[64] = {};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    size_t& configured = configured_sz[dev_ & 63];
    if (sh > 48 * 1024 && sh > configured) {
        cudaError_t e = cudaFuncSetAttribute(attention_ref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
        if (e != cudaSuccess) return e;
        configured = sh;
    }
    attention_ref_kernel<<<dim3(N, H, B), 128, sh, stream>>>(qkv, ld_qkv, kvy, ld_kvy, ymask, gate_tanh, out, N, T, H, Hkv,
                                                           hd, scale_self, scale_cross, kv_len, region);
    return cudaGetLastError();
}

}  // namespace ndit
