"""CPU oracle of the caption-encoder end.  TEST INFRASTRUCTURE ONLY (same rules as nextdit_oracle.py).

What the reference runs once per prompt batch (lumina_next_t2i/sample.py:46-50, :109-111):

    text_encoder = AutoModel.from_pretrained("google/gemma-2b", torch_dtype=dtype, device_map="cuda").eval()
    prompt_embeds = text_encoder(input_ids=ids, attention_mask=mask, output_hidden_states=True).hidden_states[-2]

The algorithm lives in a third-party dependency that is absent from /root/reference: Hugging Face ``transformers``
(``requirements.txt`` lists it unpinned); this image has transformers 5.5.0, and this file restates its
``models/gemma/modeling_gemma.py`` (line numbers below are that file's):

    scaled embedding       GemmaTextScaledWordEmbedding.forward  :60-61   (embed * embed_scale.to(weight.dtype))
    RMSNorm                GemmaRMSNorm                          :70-78   ((x * rsqrt(mean x^2 + eps)) * (1 + w), computed in fp32)
    rotary table           GemmaRotaryEmbedding                  :140-163 (fp32 angles, cos / sin cast to the activation dtype)
    rotary application     rotate_half / apply_rotary_pos_emb    :165-195
    attention              eager_attention_forward, GemmaAttention :210-300 (GQA repeat_kv, causal + padding mask, fp32 softmax)
    MLP                    GemmaMLP.forward                      :95-97   (down(act(gate(x)) * up(x)), act = gelu_pytorch_tanh)
    decoder layer          GemmaDecoderLayer.forward             :316-340
    model                  GemmaModel.forward                    :397-447; output_hidden_states collects the embedding output and every
                           layer output, the last entry being norm(last layer): hidden_states[-2] = output of layer n-2 (0-based)

Pinned: ``oracle/make_golden.py gemma`` builds transformers' own ``GemmaModel`` with seeded random weights (a small config with the
real head_dim 256 and the real activation), runs it in fp32 and in bf16 on the CPU and stores ids / mask / hidden_states[-2] under
``tests/golden/gemma_tiny.pt``; ``tests/test_oracle_vs_golden.py`` checks this restatement against it (fp32 <= 2e-5).

``precision``: "fp32", or "bf16" = fp32 arithmetic with a rounding wherever the bf16 module materialises a bf16 tensor.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .nextdit_oracle import _Prec

Tensor = torch.Tensor


@dataclass
class GemmaCfg:
    vocab_size: int = 256000
    hidden_size: int = 2048
    num_hidden_layers: int = 18
    num_attention_heads: int = 8
    num_key_value_heads: int = 1
    head_dim: int = 256
    intermediate_size: int = 16384
    rms_norm_eps: float = 1e-6
    rope_theta: float = 10000.0


def config_tiny() -> GemmaCfg:
    return GemmaCfg(vocab_size=512, hidden_size=512, num_hidden_layers=4, num_attention_heads=2, num_key_value_heads=1, head_dim=256,
                    intermediate_size=1024)


def rms_norm(p: _Prec, x: Tensor, w: Tensor, eps: float) -> Tensor:
    x32 = x.float()
    out = x32 * torch.rsqrt(x32.pow(2).mean(-1, keepdim=True) + eps)
    return p.r(out * (1.0 + w.float()))


def rotary(p: _Prec, cfg: GemmaCfg, T: int):
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, cfg.head_dim, 2, dtype=torch.int64).float() / cfg.head_dim))
    freqs = torch.arange(T).float()[:, None] * inv[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return p.r(emb.cos()), p.r(emb.sin())                      # [T, head_dim], cast to the activation dtype


def rotate_half(x: Tensor) -> Tensor:
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def hidden_states_m2(cfg: GemmaCfg, W: Dict[str, Tensor], ids: Tensor, mask: Optional[Tensor], precision: str = "fp32") -> Tensor:
    """text_encoder(ids, mask, output_hidden_states=True).hidden_states[-2]: [B, T, hidden_size]."""
    p = _Prec(precision)
    B, T = ids.shape
    H, Hkv, hd = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    scale = torch.tensor(cfg.hidden_size ** 0.5)
    if p.bf16:
        scale = scale.to(torch.bfloat16).float()
    x = p.r(p.r(W["embed_tokens.weight"].float())[ids] * scale)
    cos, sin = rotary(p, cfg, T)
    keep = torch.tril(torch.ones(T, T, dtype=torch.bool))[None, None]             # causal
    if mask is not None:
        keep = keep & (mask != 0)[:, None, None, :]                               # padded keys
    for l in range(cfg.num_hidden_layers - 1):
        pre = f"layers.{l}."
        u = rms_norm(p, x, W[pre + "input_layernorm.weight"], cfg.rms_norm_eps)
        q = p.linear(u, W[pre + "self_attn.q_proj.weight"]).view(B, T, H, hd).transpose(1, 2)
        k = p.linear(u, W[pre + "self_attn.k_proj.weight"]).view(B, T, Hkv, hd).transpose(1, 2)
        v = p.linear(u, W[pre + "self_attn.v_proj.weight"]).view(B, T, Hkv, hd).transpose(1, 2)
        q = p.r(p.r(q * cos) + p.r(rotate_half(q) * sin))
        k = p.r(p.r(k * cos) + p.r(rotate_half(k) * sin))
        k = k.repeat_interleave(H // Hkv, dim=1)
        v = v.repeat_interleave(H // Hkv, dim=1)
        s = p.r(p.r(q @ k.transpose(2, 3)) * (hd ** -0.5))
        s = s.masked_fill(~keep, float("-inf"))
        a = p.r(torch.softmax(s.float(), dim=-1))
        a = torch.nan_to_num(a)                                                    # a fully masked row cannot occur (the diagonal is kept)
        o = p.r(a @ v).transpose(1, 2).reshape(B, T, H * hd)
        x = p.r(x + p.linear(o, W[pre + "self_attn.o_proj.weight"]))
        u = rms_norm(p, x, W[pre + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
        g = p.linear(u, W[pre + "mlp.gate_proj.weight"])
        h = p.r(p.r(F.gelu(g, approximate="tanh")) * p.linear(u, W[pre + "mlp.up_proj.weight"]))
        x = p.r(x + p.linear(h, W[pre + "mlp.down_proj.weight"]))
    return x


def synthetic_weights(cfg: GemmaCfg, seed: int = 0) -> Dict[str, Tensor]:
    """Seeded weights under GemmaModel's state-dict keys (bf16), scaled so that activations stay O(1) through the stack."""
    g = torch.Generator().manual_seed(seed)
    D, F_, H, Hkv, hd = cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim
    W: Dict[str, Tensor] = {"embed_tokens.weight": torch.randn(cfg.vocab_size, D, generator=g) * (1.0 / math.sqrt(D))}
    for l in range(cfg.num_hidden_layers):
        pre = f"layers.{l}."
        W[pre + "self_attn.q_proj.weight"] = torch.randn(H * hd, D, generator=g) / math.sqrt(D)
        W[pre + "self_attn.k_proj.weight"] = torch.randn(Hkv * hd, D, generator=g) / math.sqrt(D)
        W[pre + "self_attn.v_proj.weight"] = torch.randn(Hkv * hd, D, generator=g) / math.sqrt(D)
        W[pre + "self_attn.o_proj.weight"] = torch.randn(D, H * hd, generator=g) / math.sqrt(H * hd)
        W[pre + "mlp.gate_proj.weight"] = torch.randn(F_, D, generator=g) / math.sqrt(D)
        W[pre + "mlp.up_proj.weight"] = torch.randn(F_, D, generator=g) / math.sqrt(D)
        W[pre + "mlp.down_proj.weight"] = torch.randn(D, F_, generator=g) / math.sqrt(F_)
        W[pre + "input_layernorm.weight"] = 0.1 * torch.randn(D, generator=g)
        W[pre + "post_attention_layernorm.weight"] = 0.1 * torch.randn(D, generator=g)
    W["norm.weight"] = 0.1 * torch.randn(D, generator=g)
    return {k: v.to(torch.bfloat16) for k, v in W.items()}


def synthetic_inputs(cfg: GemmaCfg, B: int = 2, T: int = 24, seed: int = 1):
    """Right-padded token ids + attention mask like the reference's tokenizer call (padding=True, pad_to_multiple_of=8)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1, cfg.vocab_size, (B, T), generator=g)
    mask = torch.ones(B, T, dtype=torch.long)
    for b in range(1, B):
        n = max(1, T - 5 * b)
        mask[b, n:] = 0
        ids[b, n:] = 0
    return ids, mask
