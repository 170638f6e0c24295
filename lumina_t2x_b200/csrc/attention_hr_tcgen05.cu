// Fused image self-attention (+ optional gated caption cross-attention), third generation: P stays in TENSOR MEMORY
// (aliased onto S) and every 128-column row of S is shared by TWO softmax threads ("half rows").
//
// Same op, interface and tensor maps as attention_tcgen05.cu (see the header there for the reference lines it replaces:
// lumina_next_t2i/models/model.py:387-434, Next-DiT-ImageNet/models/models.py:389).
//
// Why: what bounds this op on B200 is the exponential (MUFU.EX2: 16 / clk / SM -> 2048 cycles per pair of 128 x 128 blocks
// against 1280 cycles of tensor-core work at head_dim 72).  The first kernel reached 62-67 % XU utilisation for two
// reasons, both visible in its SASS / ncu capture (profiles/r01_ncu_attention_full_final.txt):
//   (1) P went through shared memory: 248 KB of shared-memory traffic per block pair = 1940 cycles at 128 B/clk - as much
//       as the exponentials themselves;
//   (2) ONE softmax warp per scheduler and tile: ptxas places the PRMT that packs a pair of exponentials 17 cycles behind
//       the second MUFU (volatile asm does not pin SASS order), the XU result takes ~23, and the in-order warp eats the
//       difference on every pair (11.6 instead of 8 cycles per element); the other tile's warp cannot fill the holes
//       because it has to be in its non-exp phases at that time or both idle together later.
// Here:
//   * P is written with tcgen05.st as packed bf16 pairs into its own tensor-memory columns and O += P V takes its A operand
//     from tensor memory: no P stores to / reads from shared memory.
//   * 16 softmax warps (4 per scheduler): a row of tile x is handled by the two threads (x, half 0/1, row), half of the
//     block's columns each, so two warps per scheduler run the exponentials of a tile together and a second pair belongs
//     to the other tile: the XU latency of one warp is hidden by the others.
//   * S, P and O of both tiles must fit the 512 TMEM columns: 2 BKV + 2 (BKV / 2) + 2 HDP <= 512 -> kv blocks of 112 rows at
//     head_dim 72 (37 blocks for 4096 tokens, the ragged last one is masked), 128 at head_dim 48, 96 at head_dim 96.
//     (First attempt, kept for the record: BKV 128 with P aliased onto S.  Q K^T of block j+1 then has to queue behind
//     P V of block j and a tile waited 1250 cycles per block for its next S - 3360 cycles per block pair, 471 us.)
//     With S and P apart, Q K^T of block j+1 is issued as soon as the softmax threads hold S(j) in registers and
//     P V of block j only has to finish before P(j+1) is stored, one block later.
//   * the row maximum is exchanged between the two half-row threads through shared memory + a 64-thread named barrier;
//     O is rescaled lazily (reference maximum may trail by 2^8) by both partners, half of the columns each.
//   * the row sum still comes from the tensor core (all-ones row HD of V^T).
//   * ONE MMA issuer for both tiles, fixed cyclic order (Q K^T A, P V A, Q K^T B, P V B): two issuers interleave on the
//     tensor pipe and pull the two tiles into lockstep (XU saturated, then idle while all 16 warps load / reduce).  Tile B
//     starts when tile A has stored its first P.
// Warp roles (640 threads, 1 CTA / SM): 0 TMA producer, 1 MMA issuer, 2-3 idle, 4-19 softmax:
// tile = (w-4)>>3, half = ((w-4)>>2)&1, TMEM lane quadrant = w&3.
// TMEM columns: S_A [0,BKV) S_B [BKV,2BKV) O_A, O_B [2BKV + x HDP, ..) P_A, P_B [2BKV + 2HDP + x PCOLS, ..).
// V^T tiles are fetched as two 64-wide boxes, the second one starting at kv0 + BKV - 64, so that every P V k-step
// (16 kv positions) is a plain 32-byte step inside a 128-byte swizzle row; the overlap is simply not used.
#include <math.h>
#include <stdlib.h>

#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace ndit {

int attention_hr_bkv(int hd) { return hd == 72 ? 112 : (hd == 48 ? 128 : (hd == 96 ? 96 : 0)); }

namespace {

#ifndef HR_SPIN_WAIT
#define mbar_wait mbar_wait_long     // every barrier wait of this kernel suspends in hardware instead of polling
#endif

constexpr int HR_BQ = 128;            // rows per query tile
constexpr int HR_THREADS = 640;
constexpr int HR_Q64_BYTES = HR_BQ * 128;        // [rows x 64] bf16, 128B swizzle
constexpr int HR_Q16_BYTES = HR_BQ * 32;         // [rows x 16] bf16, 32B swizzle
constexpr float HR_RESCALE_LOG2 = 8.0f;

template <int HD>
struct HrDims {
    static_assert(HD == 72 || HD == 48 || HD == 96, "head_dim 72, 48 or 96");
    static constexpr int BKV = HD == 72 ? 112 : (HD == 48 ? 128 : 96);   // == attention_hr_bkv(HD)
    static constexpr int HC = BKV / 2;                       // S columns per half-row thread
    static constexpr int PC = HC / 2;                        // P columns (bf16 pairs) per half-row thread
    static constexpr int NK64 = HD >= 64 ? 4 : HD / 16;
    static constexpr int N16 = HD > 64 ? (HD - 64 + 15) / 16 : 0;
    static constexpr int HDP = attn_vrows(HD);
    static constexpr int NPV = BKV / 16;                     // P V k-steps
    static constexpr int V1_OFF = BKV - 64;                  // kv offset of the second V^T box
    static constexpr int STAGES = HD > 80 ? 3 : 4;
    // tensor memory columns
    static constexpr int PCOLS = (BKV / 2 + 15) / 16 * 16;
    static constexpr uint32_t TM_S = 0;                      // + x * BKV
    static constexpr uint32_t TM_O = 2 * BKV;                // + x * HDP
    static constexpr uint32_t TM_P = 2 * BKV + 2 * HDP;      // + x * PCOLS
    static_assert(2 * BKV + 2 * HDP + 2 * PCOLS <= 512, "TMEM budget");
    // shared memory
    static constexpr int QTILE_BYTES = HR_Q64_BYTES + N16 * HR_Q16_BYTES;
    static constexpr int K64_BYTES = BKV * 128, K16_BYTES = BKV * 32;
    static constexpr int KTILE_BYTES = (K64_BYTES + N16 * K16_BYTES + 1023) / 1024 * 1024;
    static constexpr int VBOX_BYTES = HDP * 128;
    static constexpr int VTILE_BYTES = 2 * VBOX_BYTES;
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = OFF_Q + 2 * QTILE_BYTES;
    static constexpr int OFF_V = OFF_K + STAGES * KTILE_BYTES;
    static constexpr int OFF_MX = OFF_V + STAGES * VTILE_BYTES;    // row-maximum exchange: [parity 2][tile 2][half 2][128] fp32
    static constexpr int OFF_BAR = OFF_MX + 2 * 2 * 2 * 128 * 4;
    static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
    static_assert(OFF_K % 1024 == 0 && OFF_V % 1024 == 0 && VBOX_BYTES % 1024 == 0 && K64_BYTES % 256 == 0, "swizzle alignment");
    static_assert(SMEM_BYTES <= 227 * 1024, "attention smem budget");
    static constexpr int Q_TX = QTILE_BYTES;
    static constexpr int K_TX = K64_BYTES + N16 * K16_BYTES;
    static constexpr int V_TX = VTILE_BYTES;
    static constexpr int OC = HD / 2;                        // output columns per half-row thread
    static constexpr int RC = HDP / 2;                       // O columns rescaled per half-row thread
    static_assert(HC % 8 == 0 && PC % 4 == 0 && OC % 4 == 0 && RC % 8 == 0, "column split");
};

__device__ __forceinline__ float hr_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t hr_pack_trunc(float lo, float hi) {
    return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632);
}
__device__ __forceinline__ void hr_tmem_st_x4(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
// N consecutive 32-bit columns of this warp's 32 lanes -> registers (N a multiple of 8), no wait.  (HR_LD_WAIT_EACH waits
// for every tcgen05.ld before the next one is issued - tried because LDTM and MUFU share the sub-partition's MIO queue;
// 434 vs 426 us, not kept.)
template <int N>
__device__ __forceinline__ void hr_ld_cols(uint32_t taddr, uint32_t* v) {
#ifdef HR_LD_WAIT_EACH
    if constexpr (N >= 32) { tmem_ld_32x32b_x32(taddr, v); tmem_ld_wait(); hr_ld_cols<N - 32>(taddr + 32, v + 32); }
    else if constexpr (N >= 16) { tmem_ld_32x32b_x16(taddr, v); tmem_ld_wait(); hr_ld_cols<N - 16>(taddr + 16, v + 16); }
    else if constexpr (N >= 8) { tmem_ld_32x32b_x8(taddr, v); tmem_ld_wait(); hr_ld_cols<N - 8>(taddr + 8, v + 8); }
#else
    if constexpr (N >= 32) { tmem_ld_32x32b_x32(taddr, v); hr_ld_cols<N - 32>(taddr + 32, v + 32); }
    else if constexpr (N >= 16) { tmem_ld_32x32b_x16(taddr, v); hr_ld_cols<N - 16>(taddr + 16, v + 16); }
    else if constexpr (N >= 8) { tmem_ld_32x32b_x8(taddr, v); hr_ld_cols<N - 8>(taddr + 8, v + 8); }
#endif
}
// registers -> N consecutive columns (N a multiple of 4), no wait
template <int N>
__device__ __forceinline__ void hr_st_cols(uint32_t taddr, const uint32_t* v) {
    if constexpr (N >= 16) { tmem_st_32x32b_x16(taddr, v); hr_st_cols<N - 16>(taddr + 16, v + 16); }
    else if constexpr (N >= 8) { tmem_st_32x32b_x8(taddr, v); hr_st_cols<N - 8>(taddr + 8, v + 8); }
    else if constexpr (N >= 4) { hr_tmem_st_x4(taddr, v); hr_st_cols<N - 4>(taddr + 4, v + 4); }
}

}  // namespace

#ifdef HR_TIMING
__device__ long long g_hr_timing[3][64][8];      // [tile 0 / tile 1 / MMA warp][block][stamp]
#define HR_STAMP(k)                                                                                               \
    do {                                                                                                          \
        if (blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 0 && qd == 0 && hf == 0 && lane == 0 && jj < 64) \
            g_hr_timing[x][jj][k] = clock64();                                                                    \
    } while (0)
#define HR_STAMP_MMA(k)                                                                                           \
    do {                                                                                                          \
        if (blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 0 && lane == 0 && jj < 64)                       \
            g_hr_timing[2][jj][k] = clock64();                                                                    \
    } while (0)
extern "C" int ndit_debug_attn_hr_timing(long long* out) {   // clock64 stamps of CTA (1,3,0): softmax half 0 / quadrant 0 / lane 0, MMA warp
    return cudaMemcpyFromSymbol(out, g_hr_timing, sizeof(g_hr_timing)) == cudaSuccess ? 0 : -1;
}
#else
#define HR_STAMP(k)
#define HR_STAMP_MMA(k)
#endif

template <int HD>
__global__ void __launch_bounds__(HR_THREADS, 1)
attention_hr_kernel(const __grid_constant__ CUtensorMap tmQ64, const __grid_constant__ CUtensorMap tmQ16,
                    const __grid_constant__ CUtensorMap tmK64, const __grid_constant__ CUtensorMap tmK16,
                    const __grid_constant__ CUtensorMap tmVt, const __grid_constant__ CUtensorMap tmKy64,
                    const __grid_constant__ CUtensorMap tmKy16, const __grid_constant__ CUtensorMap tmVyt,
                    const uint8_t* __restrict__ ymask, const float* __restrict__ gate_tanh, bf16* __restrict__ out,
                    int N, int T, int H, int Hkv, float sl2_self, float sl2_cross, const int* __restrict__ kv_len) {
    using Dm = HrDims<HD>;
    constexpr int HDP = Dm::HDP, BKV = Dm::BKV, HC = Dm::HC, PC = Dm::PC;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* const sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t bar0 = sbase + Dm::OFF_BAR;
    auto q_full = [&](int x) { return bar0 + 8u * (0 + x); };
    auto k_full = [&](int s) { return bar0 + 8u * (2 + s); };
    auto v_full = [&](int s) { return bar0 + 8u * (2 + Dm::STAGES + s); };
    auto kv_empty = [&](int s) { return bar0 + 8u * (2 + 2 * Dm::STAGES + s); };
    constexpr int BB = 2 + 3 * Dm::STAGES;
    auto s_full = [&](int x) { return bar0 + 8u * (BB + 0 + x); };    // S_x(j) = Q K^T landed in TMEM
    auto s_free = [&](int x) { return bar0 + 8u * (BB + 2 + x); };    // the 8 softmax warps of tile x hold S_x(j) in registers
    auto p_full = [&](int x) { return bar0 + 8u * (BB + 4 + x); };    // P_x(j) in TMEM (and O_x rescaled if needed): 8 warps
    auto pv_done = [&](int x) { return bar0 + 8u * (BB + 6 + x); };   // O_x += P_x(j) V finished (P_x reusable, O_x complete up to j)
    const uint32_t stagger_bar = bar0 + 8u * (BB + 8);                 // tile A has stored its first P: tile B may start
    auto xchg = [&](int x, int qd) { return bar0 + 8u * (BB + 9 + x * 4 + qd); };   // both half-row warps posted their row maxima
    // exps_done(x, qd) (HR_XU_TURNS builds only): the two warps of tile x on scheduler qd have issued the exponentials of one
    // more block; the two tiles then take turns on each scheduler's XU pipe (A0 B0 A1 B1 ...).  Measured 430 vs 422 us
    // free-running: with the XU saturated the MIO queue it shares with LDTM / STTM / mbarrier / LDS is always full, every such
    // instruction of the waiting tile queues ~100 cycles, and its ~12 dependent MIO round trips per block (s_full, S load,
    // maxima exchange, pv_done, P store, arrives) take as long as the other tile's exponentials either way.
    auto exps_done = [&](int x, int qd) { return bar0 + 8u * (BB + 17 + x * 4 + qd); };
    const uint32_t tmem_ptr_addr = bar0 + 8u * (BB + 25);
    static_assert(8 * (BB + 26) <= 512, "barrier area");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (2 * HR_BQ);
    const int h = blockIdx.y, b = blockIdx.z;
    const int g = h / (H / Hkv);
    const int Nv = kv_len != nullptr ? kv_len[b] : N;    // valid image tokens of this batch row (variable-resolution list input)
    const int n_self = (Nv + BKV - 1) / BKV;
    const int n_cross = (T + BKV - 1) / BKV;      // 0 for the class-conditional model
    const int n_total = n_self + n_cross;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ64); tma_prefetch_desc(&tmK64); tma_prefetch_desc(&tmVt);
        if (Dm::N16 > 0) { tma_prefetch_desc(&tmQ16); tma_prefetch_desc(&tmK16); }
        if (n_cross > 0) { tma_prefetch_desc(&tmKy64); tma_prefetch_desc(&tmVyt); if (Dm::N16 > 0) tma_prefetch_desc(&tmKy16); }
        for (int x = 0; x < 2; ++x) {
            mbar_init(q_full(x), 1);
            mbar_init(s_full(x), 1);
            mbar_init(s_free(x), 8);
            mbar_init(p_full(x), 8);
            mbar_init(pv_done(x), 1);
        }
        for (int s = 0; s < Dm::STAGES; ++s) {
            mbar_init(k_full(s), 1);
            mbar_init(v_full(s), 1);
            mbar_init(kv_empty(s), 1);
        }
        mbar_init(stagger_bar, 8);
        for (int i = 0; i < 8; ++i) { mbar_init(xchg(i >> 2, i & 3), 2); mbar_init(exps_done(i >> 2, i & 3), 2); }
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
        setmaxnreg_dec<48>();
        if (warp == 0) {
            // ================================================================= TMA producer
            if (elect_one_sync()) {
                for (int x = 0; x < 2; ++x) {
                    const uint32_t dst = sbase + Dm::OFF_Q + x * Dm::QTILE_BYTES;
                    mbar_expect_tx(q_full(x), Dm::Q_TX);
                    tma_load_3d(dst, &tmQ64, q_full(x), 0, h, b * N + q0 + x * HR_BQ);
#pragma unroll
                    for (int c = 0; c < Dm::N16; ++c)
                        tma_load_3d(dst + HR_Q64_BYTES + c * HR_Q16_BYTES, &tmQ16, q_full(x), 64 + 16 * c, h, b * N + q0 + x * HR_BQ);
                }
            }
            __syncwarp();
            int s = 0;
            uint32_t ph = 0;
            for (int jj = 0; jj < n_total; ++jj) {
                mbar_wait(kv_empty(s), ph ^ 1);
                if (elect_one_sync()) {
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
        } else if (warp == 1) {
            // ================================================================= MMA issuer (both tiles, fixed cyclic order)
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, BKV);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, HDP);
            auto issue_qk = [&](int x, int jj) {
                const int s = jj % Dm::STAGES;
                mbar_wait(k_full(s), (jj / Dm::STAGES) & 1);
                if (jj > 0) mbar_wait(s_free(x), (jj - 1) & 1);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t qa = sbase + Dm::OFF_Q + x * Dm::QTILE_BYTES;
                    const uint32_t ka = sbase + Dm::OFF_K + s * Dm::KTILE_BYTES;
                    const uint64_t dq = make_smem_desc_kmajor(qa, 1024, UMMA_SW128);
                    const uint64_t dk = make_smem_desc_kmajor(ka, 1024, UMMA_SW128);
                    const uint32_t d = tmem_base + Dm::TM_S + x * BKV;
#pragma unroll
                    for (int k = 0; k < Dm::NK64; ++k) umma_ss(d, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
#pragma unroll
                    for (int c = 0; c < Dm::N16; ++c) {
                        const uint64_t dq16 = make_smem_desc_kmajor(qa + HR_Q64_BYTES + c * HR_Q16_BYTES, 256, UMMA_SW32);
                        const uint64_t dk16 = make_smem_desc_kmajor(ka + Dm::K64_BYTES + c * Dm::K16_BYTES, 256, UMMA_SW32);
                        umma_ss(d, dq16, dk16, idesc_qk, 1);
                    }
                    umma_commit(s_full(x));
                }
                __syncwarp();
            };
            auto issue_pv = [&](int x, int jj) {
                const int s = jj % Dm::STAGES;
                const uint32_t acc0 = (jj != 0 && jj != n_self) ? 1u : 0u;   // new softmax segment -> fresh accumulator
                mbar_wait(v_full(s), (jj / Dm::STAGES) & 1);
                mbar_wait(p_full(x), jj & 1);
                tc_fence_after();
                HR_STAMP_MMA(2 * x);
                if (elect_one_sync()) {
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
                    umma_commit(pv_done(x));
                    if (x == 1) umma_commit(kv_empty(s));       // every MMA that reads this K / V^T stage has been issued before
                }
                __syncwarp();
                HR_STAMP_MMA(2 * x + 1);
            };
            for (int x = 0; x < 2; ++x) {
                mbar_wait(q_full(x), 0);
                issue_qk(x, 0);
            }
            for (int jj = 0; jj < n_total; ++jj) {
                for (int x = 0; x < 2; ++x) {
                    // S_x(jj+1) is issued as soon as the softmax threads hold S_x(jj) in registers, well ahead of P_x(jj) V
                    if (jj + 1 < n_total) issue_qk(x, jj + 1);
                    issue_pv(x, jj);
                }
            }
        }
    } else {
        // ================================================================= softmax warps (two threads per row)
        setmaxnreg_inc<104>();
        const int sw = warp - 4;
        const int x = sw >> 3;                   // query tile
        const int hf = (sw >> 2) & 1;            // column half of the kv block
        const int qd = warp & 3;                 // TMEM lane quadrant
        const int r = qd * 32 + lane;            // row inside the tile
        const int qrow = q0 + x * HR_BQ + r;     // token index in this batch element
        const uint32_t lane_sel = static_cast<uint32_t>(qd * 32) << 16;
        const uint32_t ts = tmem_base + lane_sel + Dm::TM_S + x * BKV + HC * hf;            // my HC columns of S_x
        const uint32_t tp = tmem_base + lane_sel + Dm::TM_P + x * Dm::PCOLS + PC * hf;      // my PC columns of P_x
        const uint32_t to = tmem_base + lane_sel + Dm::TM_O + x * HDP;
        float* const mx_sh = reinterpret_cast<float*>(sgen + Dm::OFF_MX);
        const uint32_t my_xchg = xchg(x, qd);    // mbarrier of the two warps that share these 32 rows

        float m_ref = -INFINITY;

        // One kv block: S (my HC columns) -> registers, row maximum (exchanged with the partner thread), lazy rescale of my
        // half of O_x, p = exp2(s*sl2 - m_ref*sl2) -> packed bf16 into my PC columns of P_x.
        auto block = [&](int jj, bool cross, bool first) {
            const float sl2 = cross ? sl2_cross : sl2_self;
            uint32_t vw[2];
            if (!cross) {
                const int nvalid = Nv - jj * BKV - HC * hf;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int rem = nvalid - c * 32;
                    vw[c] = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
                }
            } else {
                const int t0 = (jj - n_self) * BKV + HC * hf;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int t = t0 + c * 32 + lane;
                    const bool ok = (c * 32 + lane < HC) && (t < T) && (ymask[b * T + t] != 0);
                    vw[c] = __ballot_sync(0xffffffffu, ok);
                }
            }
            constexpr uint32_t full1 = HC >= 64 ? 0xffffffffu : ((1u << (HC - 32)) - 1u);
            const bool all_valid = vw[0] == 0xffffffffu && (vw[1] & full1) == full1;

            HR_STAMP(0);
            mbar_wait(s_full(x), jj & 1);
            tc_fence_after();
            HR_STAMP(1);
            uint32_t sreg[HC];
            hr_ld_cols<HC>(ts, sreg);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free(x));
            HR_STAMP(2);

            // ---- row maximum of my half, then of the whole row
            float mb;
            if (all_valid) {
                float m0 = __uint_as_float(sreg[0]), m1 = __uint_as_float(sreg[1]), m2 = __uint_as_float(sreg[2]), m3 = __uint_as_float(sreg[3]);
#pragma unroll
                for (int i = 4; i < HC; i += 4) {
                    m0 = fmaxf(m0, __uint_as_float(sreg[i]));
                    m1 = fmaxf(m1, __uint_as_float(sreg[i + 1]));
                    m2 = fmaxf(m2, __uint_as_float(sreg[i + 2]));
                    m3 = fmaxf(m3, __uint_as_float(sreg[i + 3]));
                }
                mb = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            } else {
                mb = -INFINITY;
#pragma unroll
                for (int i = 0; i < HC; ++i) {
                    if (!((vw[i >> 5] >> (i & 31)) & 1u)) sreg[i] = 0xff800000u;   // -inf
                    mb = fmaxf(mb, __uint_as_float(sreg[i]));
                }
            }
            // ---- post my half's maximum for the partner (no waiting here)
            const int par = jj & 1;
            mx_sh[((par * 2 + x) * 2 + hf) * 128 + r] = mb;
            __syncwarp();
            if (lane == 0) mbar_arrive(my_xchg);
            HR_STAMP(3);
            // The exponent reference m_ref may trail the row maximum by up to 2^8 (lazy rescale), so in the common case this
            // block needs no new reference and the exponentials can start before the partner's maximum is known (speculative
            // warps).  A warp whose own half already exceeds the window waits for the partner first.  Either way both partners
            // end the block with the same m_ref: m_new is the maximum over both halves and the rule is the same.
            const bool spec = !first && !__any_sync(0xffffffffu, (mb - m_ref) * sl2 > HR_RESCALE_LOG2);
            bool pv_waited = (jj == 0) || first;    // first block of a segment: nothing in flight that touches P_x / O_x
            auto wait_pv = [&]() {
                if (!pv_waited) {
                    mbar_wait(pv_done(x), (jj - 1) & 1);        // P_x reusable, O_x complete up to block jj-1
                    tc_fence_after();
                    pv_waited = true;
                }
            };
            auto rescale_o = [&](float alpha) {                // my half of O_x (and of the row sum, column HD)
                wait_pv();
#pragma unroll 1
                for (int c = 0; c < Dm::RC / 8; ++c) {
                    uint32_t v[8];
                    tmem_ld_32x32b_x8(to + hf * Dm::RC + c * 8, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                    tmem_st_32x32b_x8(to + hf * Dm::RC + c * 8, v);
                }
            };
            auto partner_max = [&]() {
                mbar_wait(my_xchg, par);
                return mx_sh[((par * 2 + x) * 2 + (hf ^ 1)) * 128 + r];
            };
            uint32_t pk[PC];
            if (!spec) {
                const float m_new = fmaxf(m_ref, fmaxf(mb, partner_max()));
                if (first) {
                    m_ref = (m_new == -INFINITY) ? 0.f : m_new;
                } else if (__any_sync(0xffffffffu, (m_new - m_ref) * sl2 > HR_RESCALE_LOG2)) {
                    rescale_o(hr_ex2((m_ref - m_new) * sl2));
                    m_ref = m_new;
                }
            }
#ifdef HR_XU_TURNS
            if (x == 0) { if (jj > 0) mbar_wait(exps_done(1, qd), (jj - 1) & 1); }
            else mbar_wait(exps_done(0, qd), jj & 1);
#endif
            HR_STAMP(4);
            // ---- p = exp2(s*sl2 - m_ref*sl2), bf16 by truncation, column c of P = kv pair (2c, 2c+1).  Masked entries: -inf -> 0.
            {
                float moff = m_ref * sl2;
                asm volatile("" : "+f"(moff));       // pins the exponentials behind the wait above (they all depend on moff)
#pragma unroll
                for (int i = 0; i < PC; ++i) {
                    const float e0 = hr_ex2(fmaf(__uint_as_float(sreg[2 * i]), sl2, -moff));
                    const float e1 = hr_ex2(fmaf(__uint_as_float(sreg[2 * i + 1]), sl2, -moff));
                    pk[i] = hr_pack_trunc(e0, e1);
                }
            }
#ifdef HR_XU_TURNS
            asm volatile("" :: "r"(pk[PC - 1]) : "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(exps_done(x, qd));
#endif
            HR_STAMP(5);
            if (spec) {
                // the partner's half may have moved the reference: rescale what was computed against the old one (rare)
                const float m_new = fmaxf(m_ref, fmaxf(mb, partner_max()));
                if (__any_sync(0xffffffffu, (m_new - m_ref) * sl2 > HR_RESCALE_LOG2)) {
                    const float alpha = hr_ex2((m_ref - m_new) * sl2);
#pragma unroll
                    for (int i = 0; i < PC; ++i) {
                        const float2 pv = unpack_bf16(pk[i]);
                        pk[i] = hr_pack_trunc(pv.x * alpha, pv.y * alpha);
                    }
                    rescale_o(alpha);
                    m_ref = m_new;
                }
            }
            wait_pv();
            hr_st_cols<PC>(tp, pk);
            HR_STAMP(6);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(p_full(x));
                if (jj == 0 && x == 0) mbar_arrive(stagger_bar);
            }
            HR_STAMP(7);
        };

        // my half of O_x[:, 0:HD] / rowsum (column HD) -> packed bf16.  combine: dst = bf16(dst + bf16(gt * bf16(value)))
        uint32_t o_acc[Dm::OC / 2];
        auto read_o = [&](bool combine, float gt) {
            uint32_t l4[4];
            tmem_ld_32x32b_x4(to + HD, l4);       // HD is a multiple of 4
            tmem_ld_wait();
            const float inv = 1.0f / __uint_as_float(l4[0]);
#pragma unroll
            for (int c = 0; c < Dm::OC / 4; ++c) {
                uint32_t v[4];
                tmem_ld_32x32b_x4(to + hf * Dm::OC + c * 4, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float a = __uint_as_float(v[2 * i]) * inv, cc = __uint_as_float(v[2 * i + 1]) * inv;
                    if (!combine) o_acc[c * 2 + i] = pack_bf16(a, cc);
                    else {
                        const float2 sv = unpack_bf16(o_acc[c * 2 + i]);
                        o_acc[c * 2 + i] = pack_bf16(sv.x + bf16_round(gt * bf16_round(a)), sv.y + bf16_round(gt * bf16_round(cc)));
                    }
                }
            }
        };

#if !defined(HR_XU_TURNS) && !defined(HR_NO_STAGGER)
        if (x == 1) mbar_wait(stagger_bar, 0);   // anti-phase the two tiles: one is in its exp phase while the other is not
#endif
        for (int jj = 0; jj < n_self; ++jj) block(jj, false, jj == 0);
        if (n_cross > 0) {
            // ---- segment switch: O_x holds the complete self-attention numerator -> bf16(O / l) into registers
            mbar_wait(pv_done(x), (n_self - 1) & 1);
            tc_fence_after();
            read_o(false, 0.f);
            tc_fence_before();
            m_ref = -INFINITY;
            for (int jj = n_self; jj < n_total; ++jj) block(jj, true, jj == n_self);
        }
        // ---- epilogue: out = bf16(self + bf16(tanh(gate) * bf16(cross)))   (no caption segment: out = bf16(self))
        mbar_wait(pv_done(x), (n_total - 1) & 1);
        tc_fence_after();
        if (n_cross > 0) read_o(true, gate_tanh[h]);
        else read_o(false, 0.f);
        if (qrow < N) {
            bf16* dst = out + (static_cast<size_t>(b) * N + qrow) * (static_cast<size_t>(H) * HD) + h * HD + hf * Dm::OC;
#pragma unroll
            for (int i = 0; i < Dm::OC / 4; ++i)
                *reinterpret_cast<uint2*>(dst + i * 4) = make_uint2(o_acc[2 * i], o_acc[2 * i + 1]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

template <int HD>
static cudaError_t launch_attention_hr(const AttnPlan& p, cudaStream_t stream) {
    auto kern = attention_hr_kernel<HD>;
    static PerDeviceFlag flags;
    bool& configured = flags.here();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, HrDims<HD>::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const float log2e = 1.4426950408889634f;
    const dim3 grid((p.N + 2 * HR_BQ - 1) / (2 * HR_BQ), p.H, p.B);
    return launch_k(kern, grid, dim3(HR_THREADS), HrDims<HD>::SMEM_BYTES, stream, p.tmQ64, p.tmQ16, p.tmK64, p.tmK16, p.tmVt, p.tmKy64,
                    p.tmKy16, p.tmVyt, p.ymask, p.gate_tanh, p.out, p.N, p.T, p.H, p.Hkv, p.scale_self * log2e, p.scale_cross * log2e, p.kv_len);
}

cudaError_t attention_fused_hr(const AttnPlan& p, cudaStream_t stream) {
    if (p.T < 0 || p.N <= 0 || p.bkv != attention_hr_bkv(p.hd)) return cudaErrorInvalidValue;
    if (p.hd == 72) return launch_attention_hr<72>(p, stream);
    if (p.hd == 48) return launch_attention_hr<48>(p, stream);
    if (p.hd == 96) return launch_attention_hr<96>(p, stream);
    return cudaErrorInvalidValue;
}

}  // namespace ndit
