"""CPU oracle of the VAE-decode end.  TEST INFRASTRUCTURE ONLY (same rules as nextdit_oracle.py).

What the reference runs on the finished latent (lumina_next_t2i/sample.py:117-120, :237-240, inside autocast(bf16) :173):

    vae = AutoencoderKL.from_pretrained("stabilityai/sdxl-vae" | "stabilityai/sd-vae-ft-{mse,ema}", torch_dtype=torch.float32).cuda()
    samples = vae.decode(samples / factor).sample

PARITY UNPINNED AGAINST DIFFUSERS.  The algorithm lives in a third-party dependency, Hugging Face ``diffusers`` (``requirements.txt``
lists it unpinned), which is absent from /root/reference AND from this image, so no fixture could be recorded from it.  What IS pinned:
the arithmetic of every building block (ResnetBlock, 1-head attention block, nearest-2x upsample + conv, GroupNorm eps, swish, mid block)
against the independent implementation of the same LDM / taming-transformers decoder blocks that ships with transformers
(models/janus/modeling_janus.py), assembled in diffusers' order with this oracle's weights
(tests/test_oracle_vs_golden.py::test_vae_oracle_blocks_against_an_independent_ldm_decoder_implementation, 1e-5, tiny and sdxl
architectures); the assembly order, the state-dict key names and the published parameter count remain a restatement of diffusers'
published code.  This file restates that implementation (diffusers 0.2x; file names relative to ``src/diffusers/models``):

    AutoencoderKL.decode / _decode     autoencoders/autoencoder_kl.py   z = post_quant_conv(z); dec = decoder(z); DecoderOutput(sample=dec)
    Decoder.forward                    autoencoders/vae.py              conv_in -> mid_block -> up_blocks -> conv_norm_out -> SiLU -> conv_out
    UNetMidBlock2D.forward             unets/unet_2d_blocks.py          resnets[0] -> attentions[0] -> resnets[1]
    UpDecoderBlock2D.forward           unets/unet_2d_blocks.py          layers_per_block + 1 resnets, then Upsample2D (all but the last block)
    ResnetBlock2D.forward (temb None)  resnet.py                        x' = conv2(dropout(silu(norm2(conv1(silu(norm1(x))))))); (shortcut(x) + x') / 1.0
    Upsample2D.forward                 upsampling.py                    F.interpolate(scale_factor=2.0, mode="nearest") -> conv 3x3
    Attention.forward (AttnProcessor)  attention_processor.py           heads = 1: group_norm -> to_q / to_k / to_v -> softmax(q k^T / sqrt(C)) v
                                                                        -> to_out[0] -> + residual, / rescale_output_factor (1.0)

Anchors in the reference's own tree: the call sites above, the VAE scaling factors (0.18215 / 0.13025, sample.py:237) and the 8x
spatial ratio the sampler assumes (latent = resolution // 8, sample.py:204-206).

``precision``: "fp32", or "bf16" = the reference's autocast: convolutions / linears take bf16 inputs and weights, accumulate in
fp32 and return bf16; GroupNorm and SiLU compute in fp32 on the bf16 tensor (autocast's fp32 list) and their result is rounded
once, at the next convolution's input; softmax in fp32 (SDPA), probabilities and output bf16; residual sums bf16.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from .nextdit_oracle import _Prec

Tensor = torch.Tensor


@dataclass
class VaeCfg:
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32


def config_tiny() -> VaeCfg:
    return VaeCfg(block_out_channels=(128, 128, 256, 256), layers_per_block=1)


def _conv(p: _Prec, x: Tensor, W: Dict[str, Tensor], pre: str, padding: int) -> Tensor:
    return p.r(F.conv2d(p.r(x), p.r(W[pre + ".weight"].float()), p.r(W[pre + ".bias"].float()), padding=padding))


def _gn(x: Tensor, W: Dict[str, Tensor], pre: str, groups: int) -> Tensor:
    return F.group_norm(x.float(), groups, W[pre + ".weight"].float(), W[pre + ".bias"].float(), eps=1e-6)      # fp32 under autocast


def _resnet(p: _Prec, cfg: VaeCfg, W: Dict[str, Tensor], pre: str, x: Tensor) -> Tensor:
    h = _conv(p, F.silu(_gn(x, W, pre + ".norm1", cfg.norm_num_groups)), W, pre + ".conv1", 1)
    h = _conv(p, F.silu(_gn(h, W, pre + ".norm2", cfg.norm_num_groups)), W, pre + ".conv2", 1)
    if pre + ".conv_shortcut.weight" in W:
        x = _conv(p, x, W, pre + ".conv_shortcut", 0)
    return p.r(x + h)


def _attention(p: _Prec, cfg: VaeCfg, W: Dict[str, Tensor], pre: str, x: Tensor) -> Tensor:
    B, C, H, Wd = x.shape
    u = _gn(x, W, pre + ".group_norm", cfg.norm_num_groups).view(B, C, H * Wd).transpose(1, 2)      # [B, HW, C]
    lin = lambda name: p.linear(u, W[f"{pre}.{name}.weight"].float().view(C, C), W[f"{pre}.{name}.bias"])   # noqa: E731
    q, k, v = lin("to_q"), lin("to_k"), lin("to_v")
    a = torch.softmax((q @ k.transpose(1, 2)) * (1.0 / math.sqrt(C)), dim=-1)
    o = p.r(p.r(a) @ v)
    o = p.linear(o, W[pre + ".to_out.0.weight"].float().view(C, C), W[pre + ".to_out.0.bias"])
    return p.r(o.transpose(1, 2).reshape(B, C, H, Wd) + x)


def decode(cfg: VaeCfg, W: Dict[str, Tensor], z: Tensor, precision: str = "fp32") -> Tensor:
    """vae.decode(z).sample: z [B, latent_channels, h, w] -> [B, out_channels, 8h, 8w]."""
    p = _Prec(precision)
    x = _conv(p, z.float(), W, "post_quant_conv", 0)
    x = _conv(p, x, W, "decoder.conv_in", 1)
    x = _resnet(p, cfg, W, "decoder.mid_block.resnets.0", x)
    x = _attention(p, cfg, W, "decoder.mid_block.attentions.0", x)
    x = _resnet(p, cfg, W, "decoder.mid_block.resnets.1", x)
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            x = _resnet(p, cfg, W, f"decoder.up_blocks.{i}.resnets.{j}", x)
        if i < n - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = _conv(p, x, W, f"decoder.up_blocks.{i}.upsamplers.0.conv", 1)
    x = F.silu(_gn(x, W, "decoder.conv_norm_out", cfg.norm_num_groups))
    return _conv(p, x, W, "decoder.conv_out", 1)


def state_dict_shapes(cfg: VaeCfg) -> Dict[str, Tuple[int, ...]]:
    """Keys and shapes of the decode half of AutoencoderKL.state_dict() (``post_quant_conv.*`` and ``decoder.*``)."""
    ch = tuple(reversed(cfg.block_out_channels))
    S: Dict[str, Tuple[int, ...]] = {}

    def conv(pre, cin, cout, k):
        S[pre + ".weight"], S[pre + ".bias"] = (cout, cin, k, k), (cout,)

    def norm(pre, c):
        S[pre + ".weight"], S[pre + ".bias"] = (c,), (c,)

    def res(pre, cin, cout):
        norm(pre + ".norm1", cin); conv(pre + ".conv1", cin, cout, 3); norm(pre + ".norm2", cout); conv(pre + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(pre + ".conv_shortcut", cin, cout, 1)

    conv("post_quant_conv", cfg.latent_channels, cfg.latent_channels, 1)
    conv("decoder.conv_in", cfg.latent_channels, ch[0], 3)
    res("decoder.mid_block.resnets.0", ch[0], ch[0]); res("decoder.mid_block.resnets.1", ch[0], ch[0])
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", ch[0])
    for name in ("to_q", "to_k", "to_v", "to_out.0"):
        S[f"{a}.{name}.weight"], S[f"{a}.{name}.bias"] = (ch[0], ch[0]), (ch[0],)
    for i in range(len(ch)):
        cin = ch[0] if i == 0 else ch[i - 1]
        for j in range(cfg.layers_per_block + 1):
            res(f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else ch[i], ch[i])
        if i < len(ch) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", ch[i], ch[i], 3)
    norm("decoder.conv_norm_out", ch[-1])
    conv("decoder.conv_out", ch[-1], cfg.out_channels, 3)
    return S


def synthetic_weights(cfg: VaeCfg, seed: int = 0) -> Dict[str, Tensor]:
    """Seeded fp32 weights (the reference loads the VAE in fp32), scaled so that activations stay O(1) through the decoder."""
    g = torch.Generator().manual_seed(seed)
    W: Dict[str, Tensor] = {}
    for k, shp in state_dict_shapes(cfg).items():
        if "norm" in k:
            W[k] = (1.0 + 0.1 * torch.randn(shp, generator=g)) if k.endswith("weight") else 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            W[k] = 0.05 * torch.randn(shp, generator=g)
        else:
            fan_in = shp[1] * (shp[2] * shp[3] if len(shp) == 4 else 1)
            W[k] = torch.randn(shp, generator=g) * (1.0 / math.sqrt(fan_in))
    return W
