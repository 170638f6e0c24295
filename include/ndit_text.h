/* ndit_text.h - C ABI of the caption-encoder end of the Lumina-Next-T2I sampling path: the Gemma decoder
 * stack that lumina_next_t2i/sample.py runs once per prompt batch,
 *
 *     text_encoder = AutoModel.from_pretrained("google/gemma-2b", torch_dtype=dtype).eval()          (sample.py:111)
 *     prompt_embeds = text_encoder(input_ids=..., attention_mask=..., output_hidden_states=True).hidden_states[-2]   (sample.py:46-50)
 *
 * i.e. transformers' GemmaModel (third-party dependency of the reference, unpinned in its requirements; restated from
 * transformers 5.5.0 models/gemma/modeling_gemma.py, which is importable in the authoring container and pins the oracle):
 * scaled token embedding, then num_hidden_layers - 1 GemmaDecoderLayer (hidden_states[-2] is the output of the second-to-last
 * layer: the last layer and the final norm do not contribute).  Same conventions as ndit.h: extern "C", plain pointers and sizes,
 * int status codes (NDIT_OK / NDIT_ERR_*), device pointers unless a name ends in _host, bf16 compute with fp32 accumulation.
 * Built into the same shared library (lumina_t2x_b200/libndit_b200.so). */
#ifndef NDIT_TEXT_H_
#define NDIT_TEXT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ntxt_engine* ntxt_handle;

/* GemmaConfig fields the stack depends on (configuration_gemma.py); google/gemma-2b: 256000, 2048, 18, 8, 1, 256, 16384, 1e-6, 10000 */
typedef struct ntxt_config {
    int32_t vocab_size;
    int32_t hidden_size;
    int32_t num_hidden_layers;      /* layers of the checkpoint; num_hidden_layers - 1 of them are evaluated */
    int32_t num_attention_heads;
    int32_t num_key_value_heads;
    int32_t head_dim;               /* 256 */
    int32_t intermediate_size;      /* multiple of 128 */
    float rms_norm_eps;
    float rope_theta;
    int32_t max_tokens;             /* batch * sequence length the workspace is sized for */
} ntxt_config;

/* AutoModel.from_pretrained(...).eval() : allocates the packed weights and the workspace on the current device */
int ntxt_create(const ntxt_config* cfg, ntxt_handle* out);
int ntxt_destroy(ntxt_handle h);
const char* ntxt_last_error(ntxt_handle h);

/* load_state_dict: one tensor under its GemmaModel state-dict key ("embed_tokens.weight", "layers.<i>.self_attn.q_proj.weight",
 * "...k_proj...", "...v_proj...", "...o_proj...", "layers.<i>.mlp.{gate,up,down}_proj.weight", "layers.<i>.input_layernorm.weight",
 * "layers.<i>.post_attention_layernorm.weight", "norm.weight"); src_dev bf16 (dtype 0) or fp32 (dtype 1) on the device.  Tensors
 * of the last layer and the final norm are accepted and ignored (they do not reach hidden_states[-2]). */
int ntxt_set_weight(ntxt_handle h, const char* key, const void* src_dev, const int64_t* shape, int32_t ndim, int32_t dtype, void* stream);
/* strict: fails (NDIT_ERR_STATE) if a tensor the evaluated layers need is missing */
int ntxt_finalize_weights(ntxt_handle h, void* stream);

/* text_encoder(input_ids, attention_mask, output_hidden_states=True).hidden_states[-2]  (modeling_gemma.py GemmaModel.forward):
 * input_ids_dev int64 [batch, T], attention_mask_dev int64 or nullptr [batch, T] (1 = token, 0 = padding; causal attention over
 * the non-padded keys), out_dev bf16 [batch, T, hidden_size].  batch * T <= max_tokens, T <= 1024. */
int ntxt_encode(ntxt_handle h, const int64_t* input_ids_dev, const int64_t* attention_mask_dev, int32_t batch, int32_t T, void* out_dev,
                void* stream);

#ifdef __cplusplus
}
#endif
#endif
