// bf16 GEMM  C[M,N] = A[M,K] * W[N,K]^T  (both operands K-major, fp32 accumulate in TMEM).
//
// Replaces the cuBLAS calls the reference reaches through F.linear at
//   lumina_next_t2i/models/model.py:358 (wq|wk|wv), :438 (wo), :502 (w1, w3, w2), :421-422 (wk_y, wv_y).
//
// Design (sm_100a): persistent CTAs, one per SM; 192 threads:
//   warp 0     TMA producer   (cp.async.bulk.tensor 128B-swizzled tiles -> smem ring)
//   warp 1     MMA issuer     (tcgen05.mma cta_group::1, M=128, N=BN, K=16 per instruction)
//   warps 2-5  epilogue       (tcgen05.ld TMEM -> regs -> bf16 -> global)
// Two TMEM accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1.
// EPI_SWIGLU: W rows are interleaved per 256-row block as [128 rows of w1 | 128 rows of w3]
// and the epilogue writes silu(a)*b for the pair (FeedForward._forward_silu_gating, model.py:498-499),
// rounding to bf16 where the reference materialises bf16 tensors.
#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace ndit {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 192;

// Eight consecutive output columns (one 16-byte group) of this thread's row that belong to the value heads: V^T[b, g, d + e, n].
// hd % 8 == 0, so a group never straddles two heads; the 32 lanes of a warp hold 32 consecutive tokens -> 64-byte segments.
__device__ __forceinline__ void store_vt8(const GemmVtOut& vt, int row, int col, const uint32_t* v) {
    const int cv = col - vt.col0;
    const int g = cv / vt.hd, d = cv - g * vt.hd;
    const int b = row / vt.ntok, n = row - b * vt.ntok;
    bf16* dst = vt.ptr + (static_cast<size_t>(b * vt.hkv + g) * vt.vrows + d) * vt.npad + n;
#pragma unroll
    for (int e = 0; e < 8; ++e) dst[static_cast<size_t>(e) * vt.npad] = __float2bfloat16_rn(__uint_as_float(v[e]));
}

template <int BN>
struct GemmCfg {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 192 ? 5 : 6);
    static constexpr int TMEM_COLS = BN > 128 ? 512 : 256;  // two accumulator stages of BN columns, power-of-two allocation
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    bf16* __restrict__ C, const bf16* __restrict__ bias, const int* __restrict__ w_row_off, int w_row_mul, int M, int N,
                    int K, int ldc, const GemmVtOut vt, const GemmRowWin rows) {
    using Cfg = GemmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
    // barrier layout (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem ptr
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + s); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * Cfg::STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
    const int n_tiles = (N + BN - 1) / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = (K + GEMM_BK - 1) / GEMM_BK;
    // Tile order: all full-width tiles first (m-major), then the ragged last-N tiles.  With N = 3456 (fused
    // q|k|v) the 64 half-empty tiles land at the end of the persistent round-robin instead of costing a whole wave.
    const int n_full = N / BN;
    const int full_tiles = m_tiles * n_full;
    auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
        if (tile < full_tiles) { m_blk = tile / n_full; n_blk = tile - m_blk * n_full; }
        else { m_blk = tile - full_tiles; n_blk = n_full; }
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_addr, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_trigger();      // the next kernel may start its prologue on SMs this grid has left
    pdl_wait();         // A (and the buffer behind C) belong to the previous kernel until it has completed
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
    // device-side row window (written by an earlier kernel: read after pdl_wait).  The static tile schedule covers the plan's M;
    // every role skips the m-blocks past the window, identically, and counts only the tiles it really processes.
    const int Meff = rows.count != nullptr ? min(M, *rows.count) : M;
    const int row0 = rows.count != nullptr ? *rows.offset : 0;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        int stage = 0;
        uint32_t phase = 0;
        // W may be a stack of matrices (the experts of a time-gated MoE layer): the one to use is chosen on the device
        const int w_off = w_row_off != nullptr ? (*w_row_off) * w_row_mul : 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int m_blk, n_blk;
            tile_coords(tile, m_blk, n_blk);
            if (m_blk * GEMM_BM >= Meff) continue;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1);
                if (lane == 0) {
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
                    tma_load_2d(sa, &tmA, full_bar(stage), kb * GEMM_BK, row0 + m_blk * GEMM_BM);
                    tma_load_2d(sa + Cfg::A_BYTES, &tmB, full_bar(stage), kb * GEMM_BK, w_off + n_blk * BN);
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            // a ragged last-N tile only multiplies the columns that exist (UMMA N is a multiple of 16)
            int m_blk, n_blk;
            tile_coords(tile, m_blk, n_blk);
            if (m_blk * GEMM_BM >= Meff) continue;
            const int n_rem = N - n_blk * BN;
            const uint32_t idesc = make_idesc_bf16(GEMM_BM, n_rem >= BN ? BN : ((n_rem + 15) & ~15));
            const int as = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            mbar_wait(tempty_bar(as), aph ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + as * BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint64_t da = make_smem_desc_kmajor(sa, 1024, UMMA_SW128);
                    const uint64_t db = make_smem_desc_kmajor(sa + Cfg::A_BYTES, 1024, UMMA_SW128);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        // advance 16 elements (32 B) along K inside the 128B swizzle span: +2 in 16B units
                        umma_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(empty_bar(stage));
                    if (kb == num_kb - 1) umma_commit(tfull_bar(as));
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
            ++it;
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int m_blk, n_blk;
            tile_coords(tile, m_blk, n_blk);
            if (m_blk * GEMM_BM >= Meff) continue;
            const int as = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            mbar_wait(tfull_bar(as), aph);
            tc_fence_after();
            const int lrow = m_blk * GEMM_BM + q * 32 + lane;      // row inside the window
            const int row = row0 + lrow;                            // row of A / C
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
            if (EPI == EPI_STORE) {
                bf16* crow = C + static_cast<size_t>(row) * ldc + static_cast<size_t>(n_blk) * BN;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c * 32, v);
                    tmem_ld_wait();
                    const int col0 = n_blk * BN + c * 32;
                    if (lrow < Meff) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (col0 + j * 8 < N) {  // N % 8 == 0
                                if (vt.ptr != nullptr && col0 + j * 8 >= vt.col0) {
                                    store_vt8(vt, row, col0 + j * 8, v + j * 8);
                                    continue;
                                }
                                if (bias != nullptr) {   // Linear bias: added to the fp32 accumulator, one rounding (final_layer.linear)
                                    const uint4 bq = *reinterpret_cast<const uint4*>(bias + col0 + j * 8);
                                    const uint32_t b4[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const float2 bf = unpack_bf16(b4[e]);
                                        v[j * 8 + 2 * e] = __float_as_uint(__uint_as_float(v[j * 8 + 2 * e]) + bf.x);
                                        v[j * 8 + 2 * e + 1] = __float_as_uint(__uint_as_float(v[j * 8 + 2 * e + 1]) + bf.y);
                                    }
                                }
                                uint4 o;
                                o.x = pack_bf16(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1]));
                                o.y = pack_bf16(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
                                o.z = pack_bf16(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
                                o.w = pack_bf16(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
                                *reinterpret_cast<uint4*>(crow + c * 32 + j * 8) = o;
                            }
                        }
                    }
                }
            } else {
                // SwiGLU: tile columns [0,BN/2) = w1 part, [BN/2,BN) = w3 part; output width N/2
                constexpr int HN = BN / 2;
                bf16* crow = C + static_cast<size_t>(row) * ldc + static_cast<size_t>(n_blk) * HN;
#pragma unroll 1
                for (int c = 0; c < HN / 16; ++c) {
                    uint32_t a[16], b[16];
                    tmem_ld_32x32b_x16(taddr + c * 16, a);
                    tmem_ld_32x32b_x16(taddr + HN + c * 16, b);
                    tmem_ld_wait();
                    const int col0 = n_blk * HN + c * 16;
                    if (lrow < Meff) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            if (col0 + j * 8 < N / 2) {
                                float h[8];
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    const float x1 = bf16_round(__uint_as_float(a[j * 8 + e]));
                                    const float x3 = bf16_round(__uint_as_float(b[j * 8 + e]));
                                    h[e] = bf16_round(EPI == EPI_GEGLU ? gelu_tanh_f(x1) : silu_f(x1)) * x3;
                                }
                                uint4 o;
                                o.x = pack_bf16(h[0], h[1]);
                                o.y = pack_bf16(h[2], h[3]);
                                o.z = pack_bf16(h[4], h[5]);
                                o.w = pack_bf16(h[6], h[7]);
                                *reinterpret_cast<uint4*>(crow + c * 16 + j * 8) = o;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(as));
            ++it;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}


// ================================================================================================
// CTA-pair variant (cta_group::2): one 256 x 256 output tile per pair of SMs.  Each CTA loads its own
// 128 rows of A and HALF of the W tile (128 of the 256 rows), so per-SM shared-memory fill and
// operand-read traffic per flop is halved against the single-CTA kernel; the leader CTA issues one
// tcgen05.mma (M=256) that drives both tensor cores, each accumulating its 128 rows in its own TMEM.
// Used when M % 256 == 0 (the four block GEMMs at full size).
// SHARE (3x3 convolution only): a stage holds the A rows of ONE (dy, k-block) with a one-row halo on either side (136 rows: 17 swizzle
// atoms) and the W tiles of the three taps (dy, dx = -1, 0, +1); the three dx taps read the same A bytes through descriptors whose
// start address is advanced by dx rows (128 B; base offset 0, see the MMA warp), so A crosses L2 -> shared memory three times per tile, not nine.
constexpr int GEMM_SHARE_ROWS = 136;
template <int BN_, bool SHARE = false>
struct Gemm2Cfg {
    static constexpr int BN = BN_;                           // 256, or 192 when N is a multiple of 192 only (fused q|k|v)
    static constexpr int A_BYTES = (SHARE ? GEMM_SHARE_ROWS : 128) * GEMM_BK * 2;        // this CTA's A rows
    static constexpr int B_ONE = (BN / 2) * GEMM_BK * 2;     // this CTA's half of one W tile
    static constexpr int B_BYTES = (SHARE ? 3 : 1) * B_ONE;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;    // 32 KB / 28 KB
    static constexpr int STAGES = SHARE ? (BN_ == 128 ? 5 : 3) : (BN_ == 128 ? 8 : 6);
    static constexpr int TMEM_COLS = 512;                    // accumulator stages at columns 0 and 256
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

// EXT (VAE decoder): a second set of four epilogue warps (the 3x3 convolutions with 128 output channels have a main loop of only
// 18 k-blocks per tile, and their epilogue also reads the residual): warps 2-5 take the first half of the tile's columns,
// warps 6-9 the second half, each on the TMEM lane quadrant (warp & 3) it may access.
template <bool EXT> struct Gemm2Threads { static constexpr int N = EXT ? 320 : GEMM_THREADS; static constexpr int EPI_WARPS = EXT ? 8 : 4; };

template <int BN, int EPI, bool EXT, bool SHARE = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Gemm2Threads<EXT>::N, 1)
gemm2_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     bf16* __restrict__ C, const int* __restrict__ w_row_off, int w_row_mul, int M, int N, int K, int ldc,
                     const GemmVtOut vt, const GemmRowWin rows, const GemmExt ext) {
    using Cfg = Gemm2Cfg<BN, SHARE>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };                         // leader's copy is the live one
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };          // per CTA (commit multicast)
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };      // per CTA (commit multicast)
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + s); }; // leader's copy, 8 arrivals
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * Cfg::STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    const int m_tiles = (M + 255) / 256;      // ragged last tile: TMA zero-fills the rows >= M, stores are masked
    const int n_tiles = (N + BN - 1) / BN;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = SHARE ? 3 * ext.kpt : (K + GEMM_BK - 1) / GEMM_BK;     // pipeline stages per tile
    const int n_full = N / BN;
    const int full_tiles = m_tiles * n_full;
    const int n_rem = N - n_full * BN;        // width of the ragged last-N tile (0: none); a multiple of 32
    auto tile_coords = [&](int tile, int& m_blk, int& n_blk) {
        if (tile < full_tiles) { m_blk = tile / n_full; n_blk = tile - m_blk * n_full; }
        else { m_blk = tile - full_tiles; n_blk = n_full; }
    };
    // Static schedule.  Full tiles go round-robin over the pairs.  The narrow last-N tiles (fused q|k|v: N = 3456 = 13 x 256 + 128,
    // 32 half-width tiles) are dealt to the pairs that received one full tile fewer, so no pair carries more than
    // ceil(work / pairs): 6.0 tile-times instead of 6.5 on 74 pairs for the 8192 x 3456 projection.
    const int rag_tiles = num_tiles - full_tiles;
    const int rem_pairs = full_tiles % num_pairs;
    int rag_first = pair, rag_stride = num_pairs;
    if (rem_pairs != 0) { rag_first = pair >= rem_pairs ? pair - rem_pairs : -1; rag_stride = num_pairs - rem_pairs; }
    const int my_full = full_tiles > pair ? (full_tiles - pair + num_pairs - 1) / num_pairs : 0;
    const int my_rag = (rag_first >= 0 && rag_tiles > rag_first) ? (rag_tiles - rag_first + rag_stride - 1) / rag_stride : 0;
    const int my_tiles = my_full + my_rag;
    auto my_tile = [&](int it) { return it < my_full ? pair + it * num_pairs : full_tiles + rag_first + (it - my_full) * rag_stride; };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 2 * Gemm2Threads<EXT>::EPI_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc_pair(tmem_ptr_addr, Cfg::TMEM_COLS);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();              // barriers of both CTAs initialised before any remote arrive / multicast commit
    tc_fence_after();
    pdl_trigger();
    pdl_wait();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
    // device-side row window, as in the single-CTA kernel: m-blocks past it are skipped by every role of both CTAs
    const int Meff = rows.count != nullptr ? min(M, *rows.count) : M;
    const int row0 = rows.count != nullptr ? *rows.offset : 0;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        int stage = 0;
        uint32_t phase = 0;
        const int w_off = w_row_off != nullptr ? (*w_row_off) * w_row_mul : 0;     // device-selected matrix of a stacked W (MoE experts)
        for (int ti = 0; ti < my_tiles; ++ti) {
            int m_blk, n_blk;
            tile_coords(my_tile(ti), m_blk, n_blk);
            if (m_blk * 256 >= Meff) continue;
            // this CTA's half of the W tile: rows [w0, w0 + width/2) land at the start of its B buffer (a narrow tile's box
            // runs past its half - and possibly past N, zero-filled - which the narrower MMA never reads)
            const int w0 = w_off + n_blk * BN + static_cast<int>(rank) * ((n_blk < n_full ? BN : n_rem) / 2);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1);
                if (EXT ? elect_one_sync() : lane == 0) {
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    if (leader) mbar_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);   // both CTAs' bytes land on this barrier
                    int a_col = kb * GEMM_BK, a_row = row0 + m_blk * 256 + static_cast<int>(rank) * 128;
                    if (SHARE) {                    // stage = (dy, k-block): A rows with a one-row halo, W tiles of the three dx taps
                        const int dy = kb / ext.kpt, kc = kb - dy * ext.kpt;
                        tma_load_2d_pair(sa, &tmA, full_bar(stage), kc * GEMM_BK, a_row + (dy - 1) * ext.wp - 1);
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx)
                            tma_load_2d_pair(sa + Cfg::A_BYTES + dx * Cfg::B_ONE, &tmB, full_bar(stage), ((dy * 3 + dx) * ext.kpt + kc) * GEMM_BK, w0);
                    } else {
                        if (EXT && ext.kpt > 0) {   // implicit 3x3 convolution: tap (dy, dx) reads the rows shifted by dy * wp + dx
                            const int tap = kb / ext.kpt;
                            a_col = (kb - tap * ext.kpt) * GEMM_BK;
                            a_row += (tap / 3 - 1) * ext.wp + (tap % 3 - 1);
                        }
                        tma_load_2d_pair(sa, &tmA, full_bar(stage), a_col, a_row);
                        tma_load_2d_pair(sa + Cfg::A_BYTES, &tmB, full_bar(stage), kb * GEMM_BK, w0);
                    }
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        if (leader) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                int m_blk, n_blk;
                tile_coords(my_tile(ti), m_blk, n_blk);
                if (m_blk * 256 >= Meff) continue;
                const uint32_t idesc = make_idesc_bf16(256, n_blk < n_full ? BN : n_rem);
                const int as = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                mbar_wait(tempty_bar(as), aph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * 256;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    // EXT: one elected lane (elect.sync) - the MMAs then issue back to back; under `lane == 0` ptxas wraps every tcgen05
                    // instruction in an ELECT loop of ~45 issue cycles, which a 64-cycle N = 128 MMA cannot hide
                    if (EXT ? elect_one_sync() : lane == 0) {
                        const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                        if (SHARE) {
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx) {
                                // rows dx .. dx + 127 of the haloed tile: start address + dx * 128 B.  MEASURED on B200: the 128B swizzle
                                // of the operand fetch follows the absolute shared-memory address (bits 7-9 select the XOR phase), exactly
                                // like the TMA write did, so the descriptor's base-offset field stays 0 (setting it to dx gives garbage).
                                const uint64_t da = make_smem_desc_kmajor(sa + dx * 128, 1024, UMMA_SW128);
                                const uint64_t db = make_smem_desc_kmajor(sa + Cfg::A_BYTES + dx * Cfg::B_ONE, 1024, UMMA_SW128);
#pragma unroll
                                for (int k = 0; k < GEMM_BK / 16; ++k) umma_ss_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | dx | k) != 0);
                            }
                        } else {
                            const uint64_t da = make_smem_desc_kmajor(sa, 1024, UMMA_SW128);
                            const uint64_t db = make_smem_desc_kmajor(sa + Cfg::A_BYTES, 1024, UMMA_SW128);
#pragma unroll
                            for (int k = 0; k < GEMM_BK / 16; ++k) umma_ss_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        }
                        umma_commit_pair(empty_bar(stage));
                        if (kb == num_kb - 1) umma_commit_pair(tfull_bar(as));
                    }
                    __syncwarp();
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                ++it;
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..5, both CTAs)
        const int q = warp & 3;
        int it = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            int m_blk, n_blk;
            tile_coords(my_tile(ti), m_blk, n_blk);
            if (m_blk * 256 >= Meff) continue;
            const int as = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            mbar_wait(tfull_bar(as), aph);
            tc_fence_after();
            const int lrow = m_blk * 256 + static_cast<int>(rank) * 128 + q * 32 + lane;     // row inside the window
            const int row = row0 + lrow;                                                         // row of A / C
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 256;
            if (EPI == EPI_STORE) {
                bf16* crow = C + static_cast<size_t>(row) * ldc + static_cast<size_t>(n_blk) * BN;
                bool on_border = false;
                if (EXT && ext.hp > 0) {
                    const int rr = row % (ext.hp * ext.wp);
                    const int y = rr / ext.wp, x = rr - y * ext.wp;
                    on_border = y == 0 || y == ext.hp - 1 || x == 0 || x == ext.wp - 1;
                }
                constexpr int CHUNKS = (BN / 32) / (EXT ? 2 : 1);
                const int c_first = EXT ? ((warp - 2) >> 2) * CHUNKS : 0;
#pragma unroll 1
                for (int c = c_first; c < c_first + CHUNKS; ++c) {
                    const int col0 = n_blk * BN + c * 32;
                    if (col0 >= N) break;
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + c * 32, v);
                    uint4 rq4[4];
                    if (EXT && ext.resid != nullptr && lrow < Meff) {      // residual tile in flight while the accumulators arrive
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (col0 + j * 8 < N) rq4[j] = *reinterpret_cast<const uint4*>(ext.resid + static_cast<size_t>(row) * ext.ldr + col0 + j * 8);
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (col0 + j * 8 < N) {
                            if (vt.ptr != nullptr && col0 + j * 8 >= vt.col0) {
                                if (lrow < Meff) store_vt8(vt, row, col0 + j * 8, v + j * 8);
                                continue;
                            }
                            if (EXT) {
                                if (ext.out_f32 != nullptr) {
                                    if (lrow < Meff) {
                                        float* dst = ext.out_f32 + static_cast<size_t>(row) * ldc + col0 + j * 8;
                                        *reinterpret_cast<uint4*>(dst) = make_uint4(v[j * 8], v[j * 8 + 1], v[j * 8 + 2], v[j * 8 + 3]);
                                        *reinterpret_cast<uint4*>(dst + 4) = make_uint4(v[j * 8 + 4], v[j * 8 + 5], v[j * 8 + 6], v[j * 8 + 7]);
                                    }
                                    continue;
                                }
                                if (ext.bias != nullptr) {
                                    const uint4 bq = *reinterpret_cast<const uint4*>(ext.bias + col0 + j * 8);
                                    const uint32_t b4[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const float2 bf = unpack_bf16(b4[e]);
                                        v[j * 8 + 2 * e] = __float_as_uint(__uint_as_float(v[j * 8 + 2 * e]) + bf.x);
                                        v[j * 8 + 2 * e + 1] = __float_as_uint(__uint_as_float(v[j * 8 + 2 * e + 1]) + bf.y);
                                    }
                                }
                                if (ext.resid != nullptr && lrow < Meff) {
                                    const uint32_t r4[4] = {rq4[j].x, rq4[j].y, rq4[j].z, rq4[j].w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const float2 rf = unpack_bf16(r4[e]);
                                        v[j * 8 + 2 * e] = __float_as_uint(bf16_round(__uint_as_float(v[j * 8 + 2 * e])) + rf.x);
                                        v[j * 8 + 2 * e + 1] = __float_as_uint(bf16_round(__uint_as_float(v[j * 8 + 2 * e + 1])) + rf.y);
                                    }
                                }
                                if (on_border) {
#pragma unroll
                                    for (int e = 0; e < 8; ++e) v[j * 8 + e] = 0u;
                                }
                            }
                            uint4 o;
                            o.x = pack_bf16(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1]));
                            o.y = pack_bf16(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
                            o.z = pack_bf16(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
                            o.w = pack_bf16(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
                            if (lrow < Meff) *reinterpret_cast<uint4*>(crow + c * 32 + j * 8) = o;
                        }
                    }
                }
            } else {
                constexpr int HN = BN / 2;
                bf16* crow = C + static_cast<size_t>(row) * ldc + static_cast<size_t>(n_blk) * HN;
#pragma unroll 1
                for (int c = 0; c < HN / 16; ++c) {
                    uint32_t a[16], b[16];
                    tmem_ld_32x32b_x16(taddr + c * 16, a);
                    tmem_ld_32x32b_x16(taddr + HN + c * 16, b);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        float h[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float x1 = bf16_round(__uint_as_float(a[j * 8 + e]));
                            const float x3 = bf16_round(__uint_as_float(b[j * 8 + e]));
                            h[e] = bf16_round(EPI == EPI_GEGLU ? gelu_tanh_f(x1) : silu_f(x1)) * x3;
                        }
                        uint4 o;
                        o.x = pack_bf16(h[0], h[1]); o.y = pack_bf16(h[2], h[3]);
                        o.z = pack_bf16(h[4], h[5]); o.w = pack_bf16(h[6], h[7]);
                        if (lrow < Meff) *reinterpret_cast<uint4*>(crow + c * 16 + j * 8) = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_bar(as), 0);   // the leader's MMA warp owns the accumulator hand-off
            ++it;
        }
    }

    tc_fence_before();
    cluster_sync_all();              // no CTA of the pair may exit (or free TMEM) while its peer can still touch it
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int BN, int EPI, bool EXT = false, bool SHARE = false>
static cudaError_t launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, bf16* C, const int* w_row_off, int w_row_mul, int M, int N,
                                int K, int ldc, int num_sms, const GemmVtOut& vt, const GemmRowWin& rows, cudaStream_t stream,
                                const GemmExt& ext = GemmExt{}) {
    auto kern = gemm2_bf16_tn_kernel<BN, EPI, EXT, SHARE>;
    using Cfg2 = Gemm2Cfg<BN, SHARE>;
    static PerDeviceFlag flags;
    bool& configured = flags.here();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int tiles = ((M + 255) / 256) * ((N + BN - 1) / BN);
    int pairs = num_sms / 2;
    if (pairs > tiles) pairs = tiles;
    return launch_k(kern, dim3(2 * pairs), dim3(Gemm2Threads<EXT>::N), Cfg2::SMEM_BYTES, stream, tmA, tmB, C, w_row_off, w_row_mul, M, N, K, ldc, vt, rows, ext);
}

// ---------------------------------------------------------------------------- host side

template <int BN, int EPI>
static cudaError_t launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, bf16* C, const bf16* bias, const int* w_row_off, int w_row_mul,
                               int M, int N, int K, int ldc, int num_sms, const GemmVtOut& vt, const GemmRowWin& rows, cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    auto kern = gemm_bf16_tn_kernel<BN, EPI>;
    static PerDeviceFlag flags;
    bool& configured = flags.here();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM, n_tiles = (N + BN - 1) / BN;
    int grid = m_tiles * n_tiles;
    if (grid > num_sms) grid = num_sms;
    return launch_k(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, tmA, tmB, C, bias, w_row_off, w_row_mul, M, N, K, ldc, vt, rows);
}

cudaError_t gemm_bf16_tn(const GemmPlan& p, cudaStream_t stream) {
    if (p.N % 8 != 0 || p.K % 8 != 0) return cudaErrorInvalidValue;
    if (p.epi != EPI_STORE && !epi_gated(p.epi)) return cudaErrorInvalidValue;
    if (p.bias != nullptr && (p.pair || p.epi != EPI_STORE)) return cudaErrorInvalidValue;
    if (p.vt.ptr != nullptr && (p.epi != EPI_STORE || p.bias != nullptr || p.vt.col0 % 8 != 0 || p.vt.hd % 8 != 0)) return cudaErrorInvalidValue;
    if (p.use_ext) {
        // VAE decoder GEMMs: CTA-pair kernel, plain store with the GemmExt epilogue / producer; N a multiple of the tile, or (bn 256)
        // one ragged last-N tile a multiple of 32 wide
        if (!p.pair || p.epi != EPI_STORE || p.bias != nullptr || p.vt.ptr != nullptr) return cudaErrorInvalidValue;
        if (p.ext.kpt > 0 && p.K != 9 * p.ext.kpt * GEMM_BK) return cudaErrorInvalidValue;
        if (p.ext.share) {
            if (p.ext.kpt <= 0 || p.N % p.bn != 0) return cudaErrorInvalidValue;
            if (p.bn == 128) return launch_gemm2<128, EPI_STORE, true, true>(p.tmA, p.tmB, p.C, nullptr, 0, p.M, p.N, p.K, p.ldc, p.num_sms, p.vt, p.rows, stream, p.ext);
            if (p.bn == 256) return launch_gemm2<256, EPI_STORE, true, true>(p.tmA, p.tmB, p.C, nullptr, 0, p.M, p.N, p.K, p.ldc, p.num_sms, p.vt, p.rows, stream, p.ext);
            return cudaErrorInvalidValue;
        }
        if (p.bn == 128 && p.N % 128 == 0)
            return launch_gemm2<128, EPI_STORE, true>(p.tmA, p.tmB, p.C, nullptr, 0, p.M, p.N, p.K, p.ldc, p.num_sms, p.vt, p.rows, stream, p.ext);
        if (p.bn == 256 && (p.N % 256) % 32 == 0)
            return launch_gemm2<256, EPI_STORE, true>(p.tmA, p.tmB, p.C, nullptr, 0, p.M, p.N, p.K, p.ldc, p.num_sms, p.vt, p.rows, stream, p.ext);
        return cudaErrorInvalidValue;
    }
    if (p.pair) {
        // ragged M is fine (zero-filled loads, masked stores); a ragged last-N tile (multiple of 32 wide) only for plain stores
        if (p.N % p.bn != 0 && (p.epi != EPI_STORE || p.bn != 256 || (p.N % 256) % 32 != 0)) return cudaErrorInvalidValue;
        if (epi_gated(p.epi)) {
            if (p.bn != 256) return cudaErrorInvalidValue;
            if (p.epi == EPI_GEGLU) return launch_gemm2<256, EPI_GEGLU>(p.tmA, p.tmB, p.C, p.w_row_off, p.w_row_mul, p.M, p.N, p.K, p.ldc, p.num_sms, p.vt, p.rows, stream);
            return launch_gemm2<256, EPI_SWIGLU>(p.tmA, p.tmB, p.C, p.w_row_off, p.w_row_mul, p.M, p.N, p.K, p.ldc, p.num_sms, p.vt, p.rows, stream);
        }
        if (p.bn == 192) return launch_gemm2<192, EPI_STORE>(p.tmA, p.tmB, p.C, p.w_row_off, p.w_row_mul, p.M, p.N, p.K, p.ldc, p.num_sms, p.vt, p.rows, stream);
        return launch_gemm2<256, EPI_STORE>(p.tmA, p.tmB, p.C, p.w_row_off, p.w_row_mul, p.M, p.N, p.K, p.ldc, p.num_sms, p.vt, p.rows, stream);
    }
    if (epi_gated(p.epi)) {
        if (p.bn != 256 || p.N % 256 != 0) return cudaErrorInvalidValue;
        if (p.epi == EPI_GEGLU) return launch_gemm<256, EPI_GEGLU>(p.tmA, p.tmB, p.C, nullptr, p.w_row_off, p.w_row_mul, p.M, p.N, p.K, p.ldc, p.num_sms, p.vt, p.rows, stream);
        return launch_gemm<256, EPI_SWIGLU>(p.tmA, p.tmB, p.C, nullptr, p.w_row_off, p.w_row_mul, p.M, p.N, p.K, p.ldc, p.num_sms, p.vt, p.rows, stream);
    }
    if (p.bn == 256) return launch_gemm<256, EPI_STORE>(p.tmA, p.tmB, p.C, p.bias, p.w_row_off, p.w_row_mul, p.M, p.N, p.K, p.ldc, p.num_sms, p.vt, p.rows, stream);
    if (p.bn == 192) return launch_gemm<192, EPI_STORE>(p.tmA, p.tmB, p.C, p.bias, p.w_row_off, p.w_row_mul, p.M, p.N, p.K, p.ldc, p.num_sms, p.vt, p.rows, stream);
    return launch_gemm<128, EPI_STORE>(p.tmA, p.tmB, p.C, p.bias, p.w_row_off, p.w_row_mul, p.M, p.N, p.K, p.ldc, p.num_sms, p.vt, p.rows, stream);
}

}  // namespace ndit
