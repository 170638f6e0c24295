/* ndit_vae.h - C ABI of the VAE-decode end of the Lumina-Next-T2I sampling path: what
 * lumina_next_t2i/sample.py runs on the finished latent,
 *
 *     vae = AutoencoderKL.from_pretrained("stabilityai/sdxl-vae" | "stabilityai/sd-vae-ft-{mse,ema}", torch_dtype=torch.float32).cuda()   (sample.py:117-120)
 *     samples = vae.decode(samples / factor).sample              (sample.py:238, inside torch.autocast("cuda", bf16), sample.py:173)
 *
 * i.e. diffusers' AutoencoderKL.decode: post_quant_conv (1x1) then Decoder (conv_in, UNetMidBlock2D = ResnetBlock2D + single-head
 * Attention + ResnetBlock2D, four UpDecoderBlock2D of layers_per_block + 1 ResnetBlock2D each with a nearest-2x Upsample2D + 3x3
 * conv between them, GroupNorm + SiLU + conv_out).  diffusers is a third-party dependency of the reference (requirements.txt,
 * unpinned) that is absent from /root/reference and from this image: the architecture is restated from its published
 * implementation (oracle/vae_oracle.py carries the restatement; parity for this end is UNPINNED).
 * Under the reference's autocast the convolutions and linears run in bf16 with fp32 accumulation, GroupNorm and SiLU in fp32;
 * this library follows those rounding points.  Same conventions as ndit.h: extern "C", plain pointers and sizes, int status
 * codes (NDIT_OK / NDIT_ERR_*), device pointers, bf16 tensors.  Built into the same shared library (libndit_b200.so). */
#ifndef NDIT_VAE_H_
#define NDIT_VAE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nvae_engine* nvae_handle;

/* AutoencoderKL config fields the decoder depends on; sdxl-vae and sd-vae-ft-*: 4, 3, {128, 256, 512, 512}, 2, 32 */
typedef struct nvae_config {
    int32_t latent_channels;         /* 4 (<= 16) */
    int32_t out_channels;            /* 3 (<= 4) */
    int32_t block_out_channels[4];   /* encoder order, as in the config; each 128, 256 or 512 */
    int32_t layers_per_block;        /* 2: every up block has layers_per_block + 1 resnets */
    int32_t norm_num_groups;         /* 32 */
} nvae_config;

/* AutoencoderKL.from_pretrained(...).cuda(): allocates the packed decoder weights on the current device */
int nvae_create(const nvae_config* cfg, nvae_handle* out);
int nvae_destroy(nvae_handle h);
const char* nvae_last_error(nvae_handle h);

/* load_state_dict: one tensor under its AutoencoderKL state-dict key ("post_quant_conv.weight", "decoder.conv_in.weight",
 * "decoder.mid_block.resnets.0.norm1.weight", "decoder.mid_block.attentions.0.to_q.weight" (or the older "query" / "key" / "value" /
 * "proj_attn" names), "decoder.up_blocks.2.resnets.0.conv_shortcut.weight", "decoder.up_blocks.0.upsamplers.0.conv.bias", ...);
 * src_dev bf16 (dtype 0) or fp32 (dtype 1) on the device, convolution weights in PyTorch's [Cout, Cin, kh, kw] layout.  Keys under
 * "encoder." and "quant_conv." are accepted and ignored (not on the decode path). */
int nvae_set_weight(nvae_handle h, const char* key, const void* src_dev, const int64_t* shape, int32_t ndim, int32_t dtype, void* stream);
/* strict: fails (NDIT_ERR_STATE) if a tensor the decoder needs is missing */
int nvae_finalize_weights(nvae_handle h, void* stream);

/* vae.decode(z).sample: z_dev bf16 [batch, latent_channels, lat_h, lat_w] (already divided by the scaling factor, as sample.py:238
 * does), out_dev bf16 [batch, out_channels, 8 * lat_h, 8 * lat_w].  The workspace for (batch, lat_h, lat_w) is allocated on first
 * use and kept (about 2.7 GB for one 1024 x 1024 image). */
int nvae_decode(nvae_handle h, const void* z_dev, int32_t batch, int32_t lat_h, int32_t lat_w, void* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif
