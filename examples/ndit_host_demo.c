/* A host program in plain C that drives the B200 engine through the C ABI only (include/ndit.h): no Python, no PyTorch.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/ndit_host_demo.c -o ndit_host_demo \
 *       -L lumina_t2x_b200 -lndit_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/lumina_t2x_b200 -lm
 *
 * It builds a small NextDiT (dim 576, 2 blocks, GQA 8/2, head_dim 72, caption width 256) with counter-based pseudo-random
 * weights, hands every tensor to the engine under its reference state-dict key (ndit_set_weight: the loader of a real host
 * would pass the tensors of its checkpoint here), and runs the 5-point Euler solve of one 32x32 latent from HOST buffers
 * (ndit_sample_host).  It prints a checksum; tests/test_host_demo_gpu.py rebuilds the same weights in Python, runs the
 * reference-API mirror and checks that both paths give bit-identical latents. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime.h>

#include "ndit.h"

static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
/* uniform in [-1, 1): element i of stream `seed` */
static float urand(uint64_t seed, uint64_t i) {
    const uint64_t h = splitmix64(seed * 0x100000001B3ull + i);
    return (float)((double)(h >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0);
}
static uint16_t to_bf16(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static uint64_t key_seed(const char* key) {       /* FNV-1a of the state-dict key */
    uint64_t h = 0xcbf29ce484222325ull;
    for (const char* p = key; *p; ++p) { h ^= (uint8_t)*p; h *= 0x100000001b3ull; }
    return h;
}

#define CHECK(call)                                                                        \
    do {                                                                                   \
        int rc_ = (call);                                                                  \
        if (rc_ != 0) {                                                                    \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, ndit_last_error(h));      \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

static ndit_handle h = NULL;

/* value = center + amp * urand(key, i) rounded to bf16; uploaded and registered under `key` */
static int put(const char* key, int64_t rows, int64_t cols, float center, float amp) {
    const int64_t n = rows * (cols ? cols : 1);
    uint16_t* host = (uint16_t*)malloc((size_t)n * 2);
    const uint64_t seed = key_seed(key);
    for (int64_t i = 0; i < n; ++i) host[i] = to_bf16(center + amp * urand(seed, (uint64_t)i));
    void* dev = NULL;
    if (cudaMalloc(&dev, (size_t)n * 2) != cudaSuccess) return 1;
    cudaMemcpy(dev, host, (size_t)n * 2, cudaMemcpyHostToDevice);
    int64_t shape[2] = {rows, cols};
    int rc = ndit_set_weight(h, key, dev, shape, cols ? 2 : 1, NDIT_BF16, NULL);
    cudaDeviceSynchronize();
    cudaFree(dev);
    free(host);
    if (rc != 0) fprintf(stderr, "ndit_set_weight(%s): %s\n", key, ndit_last_error(h));
    return rc;
}

int main(void) {
    enum { D = 576, L = 2, H = 8, HKV = 2, HD = 72, C = 256, CD = 576, F = 1536, T = 16, LAT = 32, STEPS = 5 };
    ndit_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.dim = D; cfg.n_layers = L; cfg.n_heads = H; cfg.n_kv_heads = HKV; cfg.cap_feat_dim = C; cfg.in_channels = 4;
    cfg.patch_size = 2; cfg.multiple_of = 256; cfg.learn_sigma = 1; cfg.norm_eps = 1e-5f;
    cfg.max_tokens = 256; cfg.max_cap_len = 32; cfg.max_batch = 2;
    if (ndit_abi_version() != NDIT_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }
    if (ndit_create(&cfg, &h) != 0) { fprintf(stderr, "ndit_create: %s\n", ndit_last_error(NULL)); return 1; }

    const int KV = HKV * HD;
    int bad = 0;
    bad |= put("pad_token", D, 0, 0.f, 0.02f);
    bad |= put("x_embedder.weight", D, 16, 0.f, 0.25f);        bad |= put("x_embedder.bias", D, 0, 0.f, 0.02f);
    bad |= put("t_embedder.mlp.0.weight", CD, 256, 0.f, 0.06f); bad |= put("t_embedder.mlp.0.bias", CD, 0, 0.f, 0.02f);
    bad |= put("t_embedder.mlp.2.weight", CD, CD, 0.f, 0.04f);  bad |= put("t_embedder.mlp.2.bias", CD, 0, 0.f, 0.02f);
    bad |= put("cap_embedder.0.weight", C, 0, 1.f, 0.1f);       bad |= put("cap_embedder.0.bias", C, 0, 0.f, 0.02f);
    bad |= put("cap_embedder.1.weight", CD, C, 0.f, 0.06f);     bad |= put("cap_embedder.1.bias", CD, 0, 0.f, 0.02f);
    bad |= put("final_layer.linear.weight", 32, D, 0.f, 0.04f); bad |= put("final_layer.linear.bias", 32, 0, 0.f, 0.02f);
    bad |= put("final_layer.adaLN_modulation.1.weight", D, CD, 0.f, 0.02f);
    bad |= put("final_layer.adaLN_modulation.1.bias", D, 0, 0.f, 0.02f);
    for (int l = 0; l < L && !bad; ++l) {
        char k[128];
#define K_(suffix) (snprintf(k, sizeof(k), "layers.%d." suffix, l), k)
        bad |= put(K_("attention.gate"), H, 0, 0.f, 0.5f);
        bad |= put(K_("attention.wq.weight"), D, D, 0.f, 0.04f);     bad |= put(K_("attention.wk.weight"), KV, D, 0.f, 0.04f);
        bad |= put(K_("attention.wv.weight"), KV, D, 0.f, 0.04f);    bad |= put(K_("attention.wo.weight"), D, D, 0.f, 0.04f);
        bad |= put(K_("attention.wk_y.weight"), KV, C, 0.f, 0.06f);  bad |= put(K_("attention.wv_y.weight"), KV, C, 0.f, 0.06f);
        bad |= put(K_("attention.q_norm.weight"), D, 0, 1.f, 0.1f);  bad |= put(K_("attention.q_norm.bias"), D, 0, 0.f, 0.02f);
        bad |= put(K_("attention.k_norm.weight"), KV, 0, 1.f, 0.1f); bad |= put(K_("attention.k_norm.bias"), KV, 0, 0.f, 0.02f);
        bad |= put(K_("attention.ky_norm.weight"), KV, 0, 1.f, 0.1f); bad |= put(K_("attention.ky_norm.bias"), KV, 0, 0.f, 0.02f);
        bad |= put(K_("feed_forward.w1.weight"), F, D, 0.f, 0.04f);  bad |= put(K_("feed_forward.w3.weight"), F, D, 0.f, 0.04f);
        bad |= put(K_("feed_forward.w2.weight"), D, F, 0.f, 0.025f);
        bad |= put(K_("attention_norm1.weight"), D, 0, 1.f, 0.1f);   bad |= put(K_("attention_norm2.weight"), D, 0, 1.f, 0.1f);
        bad |= put(K_("ffn_norm1.weight"), D, 0, 1.f, 0.1f);         bad |= put(K_("ffn_norm2.weight"), D, 0, 1.f, 0.1f);
        bad |= put(K_("attention_y_norm.weight"), C, 0, 1.f, 0.1f);
        bad |= put(K_("adaLN_modulation.1.weight"), 4 * D, CD, 0.f, 0.02f);
        bad |= put(K_("adaLN_modulation.1.bias"), 4 * D, 0, 0.f, 0.02f);
#undef K_
    }
    if (bad) return 1;
    CHECK(ndit_finalize_weights(h, NULL));      /* strict: fails if a reference key is missing */

    /* host inputs: one latent duplicated for cond / uncond, caption features, caption mask (row 1 = short "empty" prompt) */
    const size_t zn = 2u * 4 * LAT * LAT;
    uint16_t* z = (uint16_t*)malloc(zn * 2);
    uint16_t* zout = (uint16_t*)malloc(zn * 2);
    for (size_t i = 0; i < zn / 2; ++i) z[i] = z[i + zn / 2] = to_bf16(1.7f * urand(key_seed("z"), i));
    uint16_t* cap = (uint16_t*)malloc((size_t)2 * T * C * 2);
    for (size_t i = 0; i < (size_t)2 * T * C; ++i) cap[i] = to_bf16(1.5f * urand(key_seed("cap"), i));
    uint8_t mask[2 * T];
    for (int i = 0; i < T; ++i) { mask[i] = 1; mask[T + i] = i < 4; }
    float grid[STEPS];
    for (int i = 0; i < STEPS; ++i) grid[i] = (float)i / (float)(STEPS - 1);     /* linspace(0, 1, 5), no time shift */
    ndit_step_params sp;
    memset(&sp, 0, sizeof(sp));
    sp.cfg_scale = 4.0f; sp.scale_factor = 1.0f; sp.scale_watershed = 1.0f; sp.proportional_attn = 1; sp.base_seqlen = 64;
    CHECK(ndit_sample_host(h, z, cap, mask, 2, LAT, LAT, T, grid, STEPS, NDIT_EULER, &sp, zout, NULL));

    double sum_abs = 0.0;
    uint64_t fnv = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < zn; ++i) {
        uint32_t u = (uint32_t)zout[i] << 16;
        float f;
        memcpy(&f, &u, 4);
        if (!isfinite(f)) { fprintf(stderr, "non-finite output\n"); return 1; }
        sum_abs += fabs(f);
        fnv = (fnv ^ zout[i]) * 0x100000001b3ull;
    }
    printf("ndit_host_demo params=%lld launches=%lld sum_abs=%.6f fnv64=%016llx\n", (long long)ndit_parameter_count(h),
           (long long)ndit_launch_count(h), sum_abs, (unsigned long long)fnv);
    ndit_destroy(h);
    free(z); free(zout); free(cap);
    return 0;
}
