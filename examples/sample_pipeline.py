"""The sampling pipeline of ``lumina_next_t2i/sample.py`` (:23-53 encode_prompt, :177-240 main loop) with all three engines of this
repository in the reference's own call sequence: caption encoder (``text_encoder(...).hidden_states[-2]``), flow-matching ODE solve
around ``NextDiT.forward_with_cfg``, VAE decode.  Everything below ``generate`` is the reference's code with the three objects swapped
for their B200 mirrors; tokenisation stays with the caller (the tokenizer is a host-side HF object on either side).

    python examples/sample_pipeline.py            # tiny random-weight models, prints the shapes of every stage (needs a B200)

``tests/test_pipeline_gpu.py`` runs the same function and checks every stage against its oracle.
"""
from __future__ import annotations

import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from lumina_t2x_b200.transport import Sampler, create_transport  # noqa: E402


def encode_prompt(text_encoder, input_ids: torch.Tensor, attention_mask: torch.Tensor):
    """sample.py:23-53 after the tokenizer call: prompt_embeds = hidden_states[-2], prompt_masks = attention_mask."""
    with torch.no_grad():
        prompt_embeds = text_encoder(input_ids=input_ids.cuda(), attention_mask=attention_mask.cuda(), output_hidden_states=True).hidden_states[-2]
    return prompt_embeds, attention_mask


def generate(text_encoder, model, vae, input_ids, attention_mask, z, *, cfg_scale=4.0, num_sampling_steps=30, sampling_method="euler",
             time_shifting_factor=1.0, proportional_attn=True, train_image_size=1024, scaling_method="Time-aware", scaling_watershed=0.3,
             vae_factor=0.13025, stages=None):
    """sample.py:177-240 for one caption: input_ids / attention_mask [2, T] = (caption, "") as the tokenizer returns them (padded to a
    multiple of 8), z [1, 4, h/8, w/8] the initial noise.  Returns the image batch [1, 3, h, w] in [0, 1].  ``stages``: optional dict that
    receives the intermediate tensors (caption features, final latent, decoded image before the clamp)."""
    dtype = torch.bfloat16
    with torch.autocast("cuda", dtype):
        sample_fn = Sampler(create_transport("Linear", "velocity", None, None, None)).sample_ode(
            sampling_method=sampling_method, num_steps=num_sampling_steps, atol=1e-6, rtol=1e-3, reverse=False,
            time_shifting_factor=time_shifting_factor)
        latent_w, latent_h = z.shape[2], z.shape[3]
        w, h = latent_w * 8, latent_h * 8
        do_extrapolation = max(w, h) > train_image_size
        z = z.to("cuda", dtype).repeat(2, 1, 1, 1)
        cap_feats, cap_mask = encode_prompt(text_encoder, input_ids, attention_mask)
        cap_mask = cap_mask.to(cap_feats.device)
        model_kwargs = dict(cap_feats=cap_feats, cap_mask=cap_mask, cfg_scale=cfg_scale)
        if proportional_attn:
            model_kwargs["proportional_attn"] = True
            model_kwargs["base_seqlen"] = (train_image_size // 16) ** 2
        else:
            model_kwargs["proportional_attn"] = False
            model_kwargs["base_seqlen"] = None
        if do_extrapolation and scaling_method == "Time-aware":
            model_kwargs["scale_factor"] = math.sqrt(w * h / train_image_size ** 2)
            model_kwargs["scale_watershed"] = scaling_watershed
        else:
            model_kwargs["scale_factor"] = 1.0
            model_kwargs["scale_watershed"] = 1.0
        samples = sample_fn(z, model.forward_with_cfg, **model_kwargs)[-1]
        samples = samples[:1]
        decoded = vae.decode(samples / vae_factor).sample
        images = (decoded + 1.0) / 2.0
        images.clamp_(0.0, 1.0)
    if stages is not None:
        stages.update(cap_feats=cap_feats, cap_mask=cap_mask, model_kwargs=model_kwargs, latent=samples, decoded=decoded)
    return images


def build_tiny(seed: int = 0):
    """Three small random-weight models with the reference architectures (oracle-side synthetic weights: test infrastructure, used here
    only to have something to run)."""
    from lumina_t2x_b200 import models
    from lumina_t2x_b200.text_encoder import GemmaTextEncoder
    from lumina_t2x_b200.vae import AutoencoderKL
    from oracle import gemma_oracle as G
    from oracle import nextdit_oracle as O
    from oracle import vae_oracle as VO
    gcfg = G.config_tiny()
    GW = G.synthetic_weights(gcfg, seed=seed)
    enc = GemmaTextEncoder(None, vocab_size=gcfg.vocab_size, hidden_size=gcfg.hidden_size, num_hidden_layers=gcfg.num_hidden_layers,
                           num_attention_heads=gcfg.num_attention_heads, num_key_value_heads=gcfg.num_key_value_heads, head_dim=gcfg.head_dim,
                           intermediate_size=gcfg.intermediate_size, rms_norm_eps=gcfg.rms_norm_eps, rope_theta=gcfg.rope_theta, max_tokens=256)
    enc.load_state_dict(GW, strict=True)
    enc = enc.eval().to("cuda", dtype=torch.bfloat16)
    import dataclasses
    dcfg = dataclasses.replace(O.config_tiny(n_layers=2), cap_feat_dim=gcfg.hidden_size)
    DW = O.synthetic_weights(dcfg, seed=seed + 1)
    dit = models.NextDiT(dim=dcfg.dim, n_layers=dcfg.n_layers, n_heads=dcfg.n_heads, n_kv_heads=dcfg.n_kv_heads, qk_norm=True,
                         cap_feat_dim=dcfg.cap_feat_dim, max_tokens=256, max_cap_len=32)
    dit.load_state_dict(DW, strict=True)
    dit = dit.eval().to("cuda", dtype=torch.bfloat16)
    vcfg = VO.config_tiny()
    VW = VO.synthetic_weights(vcfg, seed=seed + 2)
    vae = AutoencoderKL(latent_channels=vcfg.latent_channels, out_channels=vcfg.out_channels, block_out_channels=vcfg.block_out_channels,
                        layers_per_block=vcfg.layers_per_block, norm_num_groups=vcfg.norm_num_groups)
    vae.load_state_dict(VW, strict=True)
    vae = vae.cuda()
    return (enc, gcfg, GW), (dit, dcfg, DW), (vae, vcfg, VW)


def tiny_prompt(gcfg, T: int = 16, valid: int = 11, seed: int = 3):
    """Token ids as a tokenizer would return them for (caption, ""): right-padded to a multiple of 8, the empty prompt = one BOS token."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, gcfg.vocab_size, (2, T), generator=g)
    mask = torch.zeros(2, T, dtype=torch.int64)
    mask[0, :valid] = 1
    mask[1, :1] = 1
    ids = ids * mask                      # pad id 0
    ids[:, 0] = 2                         # <bos>
    return ids, mask


if __name__ == "__main__":
    (enc, gcfg, _), (dit, _, _), (vae, _, _) = build_tiny()
    ids, mask = tiny_prompt(gcfg)
    z = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(1))
    st = {}
    img = generate(enc, dit, vae, ids, mask, z, num_sampling_steps=6, train_image_size=128, stages=st)
    print("caption features", tuple(st["cap_feats"].shape), "latent", tuple(st["latent"].shape), "image", tuple(img.shape),
          "range", float(img.min()), float(img.max()))
