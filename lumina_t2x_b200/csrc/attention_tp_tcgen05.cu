// Fused image self-attention (+ optional gated caption cross-attention), second generation: P stays in TENSOR MEMORY.
//
// Same op, interface and warp roles as attention_tcgen05.cu (see the header there for the reference lines it replaces).
// What changed and why: the first kernel wrote P (bf16) to shared memory and let O += P V read it back as the A operand.
// Counting bytes per pair of 128 x 128 blocks, shared memory then moves 288 KB (Q K^T operands 80, P V operands 104,
// P stores 64, TMA fills 40) = 2304 cycles at 128 B/clk/SM - MORE than the 2048 cycles the exponentials need on the XU
// pipe; ncu agrees (l1tex data pipe: 43 % tensor-core reads + 18 % st.shared, XU 65 %).  Here a softmax thread writes its
// row of P straight into tensor memory (tcgen05.st, packed bf16 pairs, lane = row) and O += P V is issued with the A
// operand in TMEM (tcgen05.mma [d], [a], b-desc): no P stores, no P reads, 145 KB per block pair.
// TMEM has 512 columns; S (fp32) + O + P for two 128-row query tiles only fit with a narrower KV block:
//   head_dim 72:  BKV = 112:  S 2 x 112 | O 2 x 80 | P 2 x 64 (56 used)  = 512 columns
// (N = 4096 -> 37 blocks instead of 32; the ragged last block is masked like any partial block.)
//
// STATUS: correct (same parity tests as the production kernel, ndit_op_attention use_ref = 2 / engine option "attn_tp" /
// NDIT_ATTN_TP=1) but NOT faster yet, so it is not the default.  B200, config-2 shape (2 x 32 heads x 4096 x 4096, hd 72):
//   production kernel (P through shared memory, BKV 128)                          420 us   759 TFLOP/s
//   this kernel, softmax warpgroups free-running                                  474-514 us  (they fall into lockstep: XU 94 % busy
//                                                                                 inside the exp phase, idle in every other phase)
//   this kernel, warpgroups taking turns on the XU pipe (default here)            445-466 us  (exp phase of ONE warp per scheduler:
//                                                                                 11.6 cycles / element instead of 8: ptxas puts the
//                                                                                 PRMT pack ~2 MUFU slots behind its MUFU and the
//                                                                                 in-order warp eats the XU latency)
// What it established for the next round: A-from-TMEM works with a one-thread-per-row tcgen05.st of packed bf16 pairs;
// accumulator / operand column offsets need no power-of-two alignment (S at 112, P at 448); the row maximum does not have
// to precede the exponentials (speculative reference, see the softmax loop); shared-memory traffic is no longer the bound
// (wait for S drops from 260 to 135 cycles).  Missing: enough independent work per scheduler while one warpgroup is in
// its non-exp phases (3 query tiles per CTA, or exp2 partly on the FMA pipe).
//
// V^T tiles are fetched as two 64-wide boxes, the second one starting at kv0 + BKV - 64, so that every P V k-step
// (16 kv positions) is a plain 32-byte step inside a 128-byte swizzle row; the overlap is simply not used.
#include <math.h>
#include <stdlib.h>

#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace ndit {

namespace {

constexpr int TP_BQ = 128;            // rows per query tile
constexpr int TP_THREADS = 384;       // warps 0-3 control (TMA, MMA tile A, MMA tile B, idle), 4-7 softmax A, 8-11 softmax B
constexpr float TP_RESCALE_LOG2 = 8.0f;

template <int HD, int BKV>
struct TpDims {
    static_assert(BKV % 16 == 0 && BKV > 64 && BKV <= 128, "kv block");
    static constexpr int NK64 = HD >= 64 ? 4 : HD / 16;
    static constexpr int N16 = HD > 64 ? (HD - 64 + 15) / 16 : 0;
    static constexpr int HDP = attn_vrows(HD);
    static constexpr int NPV = BKV / 16;                     // P V k-steps
    static constexpr int V1_OFF = BKV - 64;                  // kv offset of the second V^T box
    static constexpr int STAGES = 4;
    // tensor memory columns
    static constexpr int PCOLS = (BKV / 2 + 15) / 16 * 16;
    static constexpr uint32_t TM_S = 0;                      // + x * BKV
    static constexpr uint32_t TM_O = 2 * BKV;                // + x * HDP
    static constexpr uint32_t TM_P = 2 * BKV + 2 * HDP;      // + x * PCOLS
    static_assert(2 * BKV + 2 * HDP + 2 * PCOLS <= 512, "TMEM budget");
    // shared memory
    static constexpr int Q64_BYTES = TP_BQ * 128, Q16_BYTES = TP_BQ * 32;
    static constexpr int QTILE_BYTES = Q64_BYTES + N16 * Q16_BYTES;
    static constexpr int K64_BYTES = BKV * 128, K16_BYTES = BKV * 32;
    static constexpr int KTILE_BYTES = (K64_BYTES + N16 * K16_BYTES + 1023) / 1024 * 1024;
    static constexpr int VBOX_BYTES = HDP * 128;
    static constexpr int VTILE_BYTES = 2 * VBOX_BYTES;
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = OFF_Q + 2 * QTILE_BYTES;
    static constexpr int OFF_V = OFF_K + STAGES * KTILE_BYTES;
    static constexpr int OFF_BAR = OFF_V + STAGES * VTILE_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
    static_assert(OFF_K % 1024 == 0 && OFF_V % 1024 == 0 && VBOX_BYTES % 1024 == 0 && K64_BYTES % 256 == 0, "swizzle alignment");
    static_assert(SMEM_BYTES <= 227 * 1024, "attention smem budget");
    static constexpr int Q_TX = QTILE_BYTES;
    static constexpr int K_TX = K64_BYTES + N16 * K16_BYTES;
    static constexpr int V_TX = VTILE_BYTES;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16_trunc(float lo, float hi) {
    return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632);
}

}  // namespace

#ifdef TP_TIMING
__device__ long long g_tp_timing[2][64][8];
#define TP_STAMP(k)                                                                                     \
    do {                                                                                                \
        if (blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 0 && qd == 0 && lane == 0 && jj < 64)  \
            g_tp_timing[x][jj][k] = clock64();                                                          \
    } while (0)
extern "C" int ndit_debug_attn_timing(long long* out) {
    return cudaMemcpyFromSymbol(out, g_tp_timing, sizeof(g_tp_timing)) == cudaSuccess ? 0 : -1;
}
#else
#define TP_STAMP(k)
#endif

template <int HD, int BKV>
__global__ void __launch_bounds__(TP_THREADS, 1)
attention_tp_kernel(const __grid_constant__ CUtensorMap tmQ64, const __grid_constant__ CUtensorMap tmQ16,
                    const __grid_constant__ CUtensorMap tmK64, const __grid_constant__ CUtensorMap tmK16,
                    const __grid_constant__ CUtensorMap tmVt, const __grid_constant__ CUtensorMap tmKy64,
                    const __grid_constant__ CUtensorMap tmKy16, const __grid_constant__ CUtensorMap tmVyt,
                    const uint8_t* __restrict__ ymask, const float* __restrict__ gate_tanh, bf16* __restrict__ out,
                    int N, int T, int H, int Hkv, float sl2_self, float sl2_cross) {
    using Dm = TpDims<HD, BKV>;
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
    auto p_full = [&](int x) { return bar0 + 8u * (BB + 4 + x); };   // P_x in TMEM (and O_x rescaled if needed)
    auto o_full = [&](int x) { return bar0 + 8u * (BB + 6 + x); };   // O_x += P_x V finished (P_x reusable)
    // exp_done(x): softmax warpgroup x has issued the exponentials of one more block.  The two warpgroups take turns on
    // the XU pipe (A0 B0 A1 B1 ...): while one runs its 112 MUFU.EX2 per thread at full rate, the other does its
    // barrier waits, TMEM loads, row maximum and TMEM stores - instead of both fighting for the XU and then both idling.
    auto exp_done = [&](int x) { return bar0 + 8u * (BB + 8 + x); };
    const uint32_t tmem_ptr_addr = bar0 + 8u * (BB + 10);
    static_assert(8 * (BB + 11) <= 256, "barrier area");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (2 * TP_BQ);
    const int h = blockIdx.y, b = blockIdx.z;
    const int g = h / (H / Hkv);
    const int n_self = (N + BKV - 1) / BKV;
    const int n_cross = (T + BKV - 1) / BKV;
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
        mbar_init(exp_done(0), 4);
        mbar_init(exp_done(1), 4);
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

    if (warp < 4) {
        setmaxnreg_dec<56>();
        if (warp == 0) {
            // ================================================================= TMA producer
            if (lane == 0) {
                for (int x = 0; x < 2; ++x) {
                    const uint32_t dst = sbase + Dm::OFF_Q + x * Dm::QTILE_BYTES;
                    mbar_expect_tx(q_full(x), Dm::Q_TX);
                    tma_load_3d(dst, &tmQ64, q_full(x), 0, h, b * N + q0 + x * TP_BQ);
#pragma unroll
                    for (int c = 0; c < Dm::N16; ++c)
                        tma_load_3d(dst + Dm::Q64_BYTES + c * Dm::Q16_BYTES, &tmQ16, q_full(x), 64 + 16 * c, h, b * N + q0 + x * TP_BQ);
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
                    mbar_expect_tx(k_full(s), Dm::K_TX);
                    mbar_expect_tx(v_full(s), Dm::V_TX);
                    const bool self = jj < n_self;
                    const int kv0 = (self ? jj : jj - n_self) * BKV;
                    const int tok0 = (self ? b * N : b * T) + kv0;
                    const CUtensorMap* mk64 = self ? &tmK64 : &tmKy64;
                    const CUtensorMap* mk16 = self ? &tmK16 : &tmKy16;
                    const CUtensorMap* mv = self ? &tmVt : &tmVyt;
                    tma_load_3d(kd, mk64, k_full(s), 0, g, tok0);
#pragma unroll
                    for (int c = 0; c < Dm::N16; ++c)
                        tma_load_3d(kd + Dm::K64_BYTES + c * Dm::K16_BYTES, mk16, k_full(s), 64 + 16 * c, g, tok0);
                    tma_load_3d(vd, mv, v_full(s), kv0, 0, b * Hkv + g);
                    tma_load_3d(vd + Dm::VBOX_BYTES, mv, v_full(s), kv0 + Dm::V1_OFF, 0, b * Hkv + g);
                }
                __syncwarp();
                if (++s == Dm::STAGES) { s = 0; ph ^= 1; }
            }
        } else if (warp == 1 || warp == 2) {
            // ================================================================= MMA issuers: warp 1 -> tile A, warp 2 -> tile B
            const int x = warp - 1;
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, BKV);
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
                    const uint32_t d = tmem_base + Dm::TM_S + x * BKV;
#pragma unroll
                    for (int k = 0; k < Dm::NK64; ++k) umma_ss(d, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
#pragma unroll
                    for (int c = 0; c < Dm::N16; ++c) {
                        const uint64_t dq16 = make_smem_desc_kmajor(qa + Dm::Q64_BYTES + c * Dm::Q16_BYTES, 256, UMMA_SW32);
                        const uint64_t dk16 = make_smem_desc_kmajor(ka + Dm::K64_BYTES + c * Dm::K16_BYTES, 256, UMMA_SW32);
                        umma_ss(d, dq16, dk16, idesc_qk, 1);
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
                    const uint32_t va = sbase + Dm::OFF_V + s * Dm::VTILE_BYTES;
                    const uint32_t d = tmem_base + Dm::TM_O + x * HDP;
                    const uint32_t pa = tmem_base + Dm::TM_P + x * Dm::PCOLS;
#pragma unroll
                    for (int k = 0; k < Dm::NPV; ++k) {
                        // kv positions 16k .. 16k+15: box 0 holds 0..63, box 1 holds V1_OFF..V1_OFF+63
                        const int box = k < 4 ? 0 : 1;
                        const int kk = k < 4 ? k : k - Dm::V1_OFF / 16;
                        const uint64_t dv = make_smem_desc_kmajor(va + box * Dm::VBOX_BYTES, 1024, UMMA_SW128) + 2 * kk;
                        umma_ts(d, pa + 8 * k, dv, idesc_pv, acc0 | (k != 0));
                    }
                    umma_commit(o_full(x));
                    umma_commit(kv_empty(s));                   // second arrival (other tile) releases the K / V^T stage
                }
                __syncwarp();
            };
            mbar_wait(q_full(x), 0);
            issue_qk(0);
            for (int jj = 0; jj < n_total; ++jj) {
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
        const int qrow = q0 + x * TP_BQ + r;     // token index in this batch element
        const uint32_t lane_sel = static_cast<uint32_t>(qd * 32) << 16;
        const uint32_t ts = tmem_base + lane_sel + Dm::TM_S + x * BKV;
        const uint32_t to = tmem_base + lane_sel + Dm::TM_O + x * HDP;
        const uint32_t tp = tmem_base + lane_sel + Dm::TM_P + x * Dm::PCOLS;

        uint32_t o_self[HD / 2];
#pragma unroll
        for (int i = 0; i < HD / 2; ++i) o_self[i] = 0u;
        float m_ref = -INFINITY;     // exponent reference of the running segment (may trail the true row maximum)
        float m_pend = -INFINITY;    // max(m_ref, row maximum of every block seen so far)

        // O_x[:, 0:HD] / rowsum (column HD) -> packed bf16.  combine: dst = bf16(dst + bf16(gt * bf16(value)))
        auto read_o = [&](uint32_t* dst, bool combine, float gt) {
            uint32_t l8[8];
            tmem_ld_32x32b_x8(to + HD, l8);
            tmem_ld_wait();
            const float inv = 1.0f / __uint_as_float(l8[0]);
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

        for (int jj = 0; jj < n_total; ++jj) {
            const bool cross = jj >= n_self;
            const bool first = (jj == 0) || (jj == n_self);
            if (jj == n_self) {
                mbar_wait(o_full(x), (jj - 1) & 1);
                tc_fence_after();
                read_o(o_self, false, 0.f);
                m_ref = m_pend = -INFINITY;
            }
            const float sl2 = cross ? sl2_cross : sl2_self;
            // validity bits of this block's BKV kv columns, 32 per word (the last word may be partial)
            constexpr int NW = (BKV + 31) / 32;
            uint32_t vw[NW];
            bool all_valid = true;
#pragma unroll
            for (int c = 0; c < NW; ++c) {
                const int width = (BKV - c * 32) >= 32 ? 32 : (BKV - c * 32);
                const uint32_t full = width == 32 ? 0xffffffffu : ((1u << width) - 1u);
                if (!cross) {
                    const int rem = N - jj * BKV - c * 32;
                    vw[c] = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
                } else {
                    const int t = (jj - n_self) * BKV + c * 32 + lane;
                    const bool ok = (lane < width) && (t < T) && (ymask[b * T + t] != 0);
                    vw[c] = __ballot_sync(0xffffffffu, ok);
                }
                vw[c] &= full;
                all_valid = all_valid && (vw[c] == full);
            }

            // ---- the whole row of S into registers, then hand S_x back to the tensor core
            TP_STAMP(0);
            mbar_wait(s_full(x), jj & 1);
            tc_fence_after();
            TP_STAMP(1);
            uint32_t sreg[BKV];
#pragma unroll
            for (int c = 0; c < BKV / 32; ++c) tmem_ld_32x32b_x32(ts + c * 32, sreg + c * 32);
            if (BKV % 32 == 16) tmem_ld_32x32b_x16(ts + (BKV / 32) * 32, sreg + (BKV / 32) * 32);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free(x));
            TP_STAMP(2);

            // ---- reference maximum.  The exponent reference m_ref may be STALE: any reference is exact after the final
            // division by the row sum (which uses the same reference), as long as the exponentials stay in range.  So the
            // row maximum of this block is not needed before its exponentials: it is accumulated inside the exp loop
            // (FMNMX in the issue slots the XU leaves free) and only decides whether O is rescaled BEFORE THE NEXT block.
            // Only the first block of a segment (no reference yet) and masked blocks take the up-front pass.
            if (!all_valid) {
#pragma unroll
                for (int i = 0; i < BKV; ++i)
                    if (!((vw[i >> 5] >> (i & 31)) & 1u)) sreg[i] = 0xff800000u;   // -inf
            }
            if (first) {
                float mb = -INFINITY;
#pragma unroll
                for (int i = 0; i < BKV; ++i) mb = fmaxf(mb, __uint_as_float(sreg[i]));
                m_ref = (mb == -INFINITY) ? 0.f : mb;
                m_pend = m_ref;
            } else if (__any_sync(0xffffffffu, (m_pend - m_ref) * sl2 > TP_RESCALE_LOG2)) {
                // lazy rescale decided by the previous block's maximum (warp-uniform: tcgen05.ld/st are warp-collective)
                const float alpha = ex2_approx((m_ref - m_pend) * sl2);
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
                m_ref = m_pend;
            }
            TP_STAMP(3);
            if (jj > 0) {                       // P_x reusable: P_x(jj-1) V has completed
                mbar_wait(o_full(x), (jj - 1) & 1);
                tc_fence_after();
            }
#ifndef TP_NO_XU_TURNS
            if (x == 0) { if (jj > 0) mbar_wait(exp_done(1), (jj - 1) & 1); }
            else mbar_wait(exp_done(0), jj & 1);
#endif
            TP_STAMP(4);
            // ---- p = exp2(s*sl2 - moff) -> P_x in TMEM: column c holds the bf16 pair (kv 2c, kv 2c+1) of this row.
            // The store of chunk c is issued after the exponentials of chunk c+1 (its operands are long ready by then,
            // so the warp-collective tcgen05.st never stalls the MUFU stream behind it).
            float mx0 = -INFINITY, mx1 = -INFINITY;
            // Issue order is pinned with volatile asm: the pack (PRMT) of a pair of exponentials is issued one chunk
            // (16 MUFU slots, >= 128 cycles) after the MUFUs that produce it.  Left to ptxas, the PRMT lands a few slots
            // behind its MUFU and the in-order warp stalls on the XU latency - fatal now that the two softmax warpgroups
            // take turns and a single warp per scheduler has to keep the XU pipe fed.
            auto exp_block = [&](float moff) {
                float e[2][16];
#pragma unroll
                for (int c = 0; c <= BKV / 16; ++c) {
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (c < BKV / 16) {
                            const float s0 = __uint_as_float(sreg[c * 16 + 2 * i]), s1 = __uint_as_float(sreg[c * 16 + 2 * i + 1]);
                            mx0 = fmaxf(mx0, s0);
                            mx1 = fmaxf(mx1, s1);
                            const float a0 = fmaf(s0, sl2, -moff), a1 = fmaf(s1, sl2, -moff);
                            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[c & 1][2 * i]) : "f"(a0));
                            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[c & 1][2 * i + 1]) : "f"(a1));
                        }
                        if (c > 0)
                            asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(pk[i])
                                         : "r"(__float_as_uint(e[(c - 1) & 1][2 * i])), "r"(__float_as_uint(e[(c - 1) & 1][2 * i + 1])));
                    }
                    if (c > 0) tmem_st_32x32b_x8(tp + (c - 1) * 8, pk);
                }
            };
            float moff = m_ref * sl2;
            asm volatile("" : "+f"(moff));       // pins the exponentials behind the barrier waits above (they all depend on moff)
#pragma unroll 1
            for (int pass = 0;; ++pass) {
                exp_block(moff);
                m_pend = fmaxf(m_ref, fmaxf(mx0, mx1));
                // range guard (practically never taken): the stale reference is more than 2^64 below this block's maximum
                if (pass == 1 || !__any_sync(0xffffffffu, (m_pend - m_ref) * sl2 > 64.0f)) break;
                const float alpha = ex2_approx((m_ref - m_pend) * sl2);
                if (!first) {                   // O_x is complete up to block jj-1 (o_full waited above): rescale it now
#pragma unroll 1
                    for (int c = 0; c < HDP / 16; ++c) {
                        uint32_t v[16];
                        tmem_ld_32x32b_x16(to + c * 16, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                        tmem_st_32x32b_x16(to + c * 16, v);
                    }
                }
                m_ref = m_pend;
                tmem_st_wait();
                moff = m_ref * sl2;             // redo this block's P against the new reference
            }
#ifndef TP_NO_XU_TURNS
            __syncwarp();
            if (lane == 0) mbar_arrive(exp_done(x));
#endif
            TP_STAMP(5);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full(x));
            TP_STAMP(6);
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

template <int HD, int BKV>
static cudaError_t launch_attention_tp(const AttnPlan& p, cudaStream_t stream) {
    auto kern = attention_tp_kernel<HD, BKV>;
    static PerDeviceFlag flags;
    bool& configured = flags.here();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TpDims<HD, BKV>::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const float log2e = 1.4426950408889634f;
    const dim3 grid((p.N + 2 * TP_BQ - 1) / (2 * TP_BQ), p.H, p.B);
    kern<<<grid, TP_THREADS, TpDims<HD, BKV>::SMEM_BYTES, stream>>>(p.tmQ64, p.tmQ16, p.tmK64, p.tmK16, p.tmVt, p.tmKy64, p.tmKy16,
                                                                    p.tmVyt, p.ymask, p.gate_tanh, p.out, p.N, p.T, p.H, p.Hkv,
                                                                    p.scale_self * log2e, p.scale_cross * log2e);
    return cudaGetLastError();
}

int attention_tp_bkv(int hd) { return hd == 72 ? 112 : 0; }

cudaError_t attention_fused_tp(const AttnPlan& p, cudaStream_t stream) {
    if (p.T < 0 || p.N <= 0 || p.bkv != attention_tp_bkv(p.hd)) return cudaErrorInvalidValue;
    if (p.hd == 72) return launch_attention_tp<72, 112>(p, stream);
    return cudaErrorInvalidValue;
}

}  // namespace ndit
