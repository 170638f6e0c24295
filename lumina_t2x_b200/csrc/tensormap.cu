// Host-side CUtensorMap construction.  cuTensorMapEncodeTiled is resolved at run time through
// cudaGetDriverEntryPoint so the shared library links only against the (static) CUDA runtime and
// loads on a machine without libcuda (the CPU-side "does the C-ABI load" test).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kernels.h"

namespace ndit {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static thread_local char g_tmap_err[256] = "";
const char* tmap_last_error() { return g_tmap_err; }

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
        snprintf(g_tmap_err, sizeof(g_tmap_err), "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

static CUtensorMapSwizzle swz(int bytes) {
    switch (bytes) {
        case 128: return CU_TENSOR_MAP_SWIZZLE_128B;
        case 64: return CU_TENSOR_MAP_SWIZZLE_64B;
        case 32: return CU_TENSOR_MAP_SWIZZLE_32B;
        default: return CU_TENSOR_MAP_SWIZZLE_NONE;
    }
}

static int encode(CUtensorMap* out, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, int swizzle_bytes) {
    EncodeTiledFn fn = get_encode();
    if (!fn) return -1;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_tmap_err, sizeof(g_tmap_err),
                 "cuTensorMapEncodeTiled failed (%d): rank %d base %p dims [%llu,%llu,%llu] box [%u,%u,%u] swizzle %d",
                 static_cast<int>(r), rank, base, (unsigned long long)dims[0], (unsigned long long)dims[1],
                 (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0, swizzle_bytes);
        return -1;
    }
    return 0;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows,
                 uint32_t box_cols, int swizzle_bytes) {
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {ld_elems * 2};
    const cuuint32_t box[2] = {box_cols, box_rows};
    return encode(out, base, 2, dims, strides, box, swizzle_bytes);
}

int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes,
                 uint64_t s2_bytes, uint32_t b0, uint32_t b1, uint32_t b2, int swizzle_bytes) {
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {s1_bytes, s2_bytes};
    const cuuint32_t box[3] = {b0, b1, b2};
    return encode(out, base, 3, dims, strides, box, swizzle_bytes);
}

int make_gemm_plan(GemmPlan* p, const bf16* A, int lda, const bf16* W, bf16* C, int ldc, int M, int N, int K, int epi,
                   int num_sms, int allow_pair, int w_total_rows) {
    memset(p, 0, sizeof(*p));
    p->C = C; p->M = M; p->N = N; p->K = K; p->ldc = ldc; p->epi = epi; p->num_sms = num_sms;
    // BN = 256 keeps the per-flop shared-memory operand traffic under the 128 B/clk/SM port limit
    // (12 KB per 128-cycle MMA); a ragged last N tile is zero-filled by TMA and masked in the epilogue.
    // BN = 128 only when there would be too few 256-wide tiles to occupy the SMs.
    const int m_tiles = (M + 127) / 128;
    int bn = 256;
    if (!epi_gated(epi) && m_tiles * ((N + 255) / 256) < num_sms) bn = 128;
    // wave quantisation: the fused q|k|v projection (N = 3456 = 13.5 x 256 = 18 x 192) needs 7 rounds of 256-wide tiles on 148
    // SMs (6.05 waves) but 8 rounds of 192-wide ones: 8 x 0.75 = 6.0 tile-times instead of 6.5
    // MEASURED (B200, 8192 x 3456 x 2304): 108.7 us with 192-wide tiles vs 104.1 us with 256-wide ones - the extra shared-memory
    // operand traffic per flop costs more than the saved half round.  Off unless NDIT_GEMM_BN192=1.
    static const int bn192_env = getenv("NDIT_GEMM_BN192") ? atoi(getenv("NDIT_GEMM_BN192")) : 0;
    if (bn192_env && bn == 256 && !epi_gated(epi) && N % 192 == 0 && N % 256 != 0) {
        const long t256 = (long)m_tiles * ((N + 255) / 256), t192 = (long)m_tiles * (N / 192);
        const long r256 = (t256 + num_sms - 1) / num_sms, r192 = (t192 + num_sms - 1) / num_sms;
        if (r192 * 3 < r256 * 4) bn = 192;
    }
    p->bn = bn;
    // CTA-pair kernel for the large GEMMs: 256x256 tiles, each CTA loads 128 rows of A and 128 rows of W
    static const int pair_env = getenv("NDIT_GEMM_PAIR") ? atoi(getenv("NDIT_GEMM_PAIR")) : 1;
    // (N must be a multiple of the tile width: 256, or 192 for the fused q|k|v projection, N = 3456 = 18 x 192)
    p->pair = 0;
    if (allow_pair && pair_env && M >= 256) {
        // BN = 192 (N = 3456) measured slower than the single-CTA kernel (108 vs 103 us): only behind NDIT_GEMM_PAIR=2
        // N need not be a multiple of 256: a narrower last-N tile (a multiple of 32 columns) runs as a narrower cta_group::2 MMA
        // (fused q|k|v, N = 3456: 13 full tiles + one 128-wide per 256 rows).  NDIT_GEMM_PAIR=2: 192-wide tiles instead (slower).
        int pbn = (N % 256 == 0 || (!epi_gated(epi) && (N % 256) % 32 == 0 && N > 256)) ? 256 : 0;
        if (pair_env >= 2 && N % 192 == 0 && N % 256 != 0 && !epi_gated(epi)) pbn = 192;
        if (pbn && ((M + 255) / 256) * ((N + pbn - 1) / pbn) >= num_sms / 2) { p->pair = 1; p->bn = pbn; }
    }
    if (make_tmap_2d(&p->tmA, A, M, K, lda, 128, 64, 128)) return -1;
    if (make_tmap_2d(&p->tmB, W, w_total_rows > N ? w_total_rows : N, K, K, p->pair ? p->bn / 2 : bn, 64, 128)) return -1;
    return 0;
}

}  // namespace ndit
