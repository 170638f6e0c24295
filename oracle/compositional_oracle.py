"""CPU restatement of the compositional (region-masked) NextDiT of ``lumina_next_compositional_generation/models/model.py``.

TEST INFRASTRUCTURE: imported by ``tests/``, ``oracle/make_golden.py`` only - never by ``lumina_t2x_b200/``.

The compositional model is the Lumina-Next-T2I NextDiT (``oracle/nextdit_oracle.py``) with three changes, each cited below:

  * ``forward`` (:852-899): the adaLN conditioning pools the GLOBAL caption (one row, broadcast to the cond and the uncond row,
    :866-870); a boolean ``region_mask`` [num_y, H/2 * W/2] is built from ``h_split_num`` x ``w_split_num`` rectangles, rectangle
    (i, j) -> caption row ``(i + 1) * (j + 1) - 1`` (:872-884, the reference's own formula, collisions included), last row (the
    unconditional caption) = all ones (:885);
  * ``Attention.forward`` (:421-446): the RoPE'd queries of the cond row are repeated for every region caption, every caption runs its
    own masked SDPA (``y_mask & region_mask``), rows without a valid key give NaN -> ``nan_to_num`` -> 0, the gated outputs of the cond
    captions are summed (:444) and added to the self-attention output;
  * ``forward_with_cfg`` (:902-953): the extra kwargs are passed through.

Pinned against the unmodified reference by ``oracle/make_golden.py compositional`` (``tests/golden/compositional_*.pt``), fp32 on CPU
(the reference's fp32 SDPA branch does not repeat kv heads, model.py:407-417, so the fixtures use n_kv_heads = n_heads).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import Tensor

from . import nextdit_oracle as O
from .nextdit_oracle import NextDiTConfig, _Prec, apply_rope, feed_forward, modulate, patchify, rms_norm, rope_angles, timestep_embedding, unpatchify


def region_mask(num_y: int, hp: int, wp: int, h_split: int, w_split: int) -> Tensor:
    """model.py:872-887 on the token grid: hp = H // patch_size, wp = W // patch_size.  Bool [num_y, hp * wp]."""
    m = torch.zeros(num_y, hp, wp)
    hps, wps = hp // h_split, wp // w_split          # == H // h_split // patch_size (nested floors)
    for i in range(h_split):
        for j in range(w_split):
            rid = (i + 1) * (j + 1) - 1
            m[rid, hps * i: hps * (i + 1), wps * j: wps * (j + 1)] = 1          # IndexError beyond num_y rows, like the reference
    m[-1] = 1
    return m.flatten(1) > 0.5


def _sdpa_full_mask(p: _Prec, q: Tensor, k: Tensor, v: Tensor, scale: float, mask: Tensor) -> Tensor:
    """softmax(q k^T * scale + mask) v, mask bool [B, 1 or H, N, T]; fully masked rows -> NaN (as torch SDPA's math path)."""
    if not p.bf16:
        return F.scaled_dot_product_attention(q, k, v, attn_mask=mask.expand(-1, q.shape[1], -1, -1), scale=scale)
    s = torch.matmul(q, k.transpose(-1, -2)) * scale
    s = s.masked_fill(~mask, float("-inf"))
    a = torch.softmax(s, dim=-1)                      # all -inf -> NaN
    return p.r(torch.matmul(p.r(a), v))


def attention(p: _Prec, cfg: NextDiTConfig, W: Dict[str, Tensor], pre: str, x: Tensor, ang: Tensor, y: Tensor, y_mask: Tensor,
              rmask: Tensor, softmax_scale: float) -> Tensor:
    """model.py:337-450.  x [2, N, D] (cond, uncond); y [num_y, T, C]; y_mask bool [num_y, T]; rmask bool [num_y, N]."""
    B, N, _ = x.shape
    H, Hkv, hd = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim
    xq = p.linear(x, W[pre + "wq.weight"])
    xk = p.linear(x, W[pre + "wk.weight"])
    xv = p.linear(x, W[pre + "wv.weight"])
    xq = F.layer_norm(xq, (H * hd,), W[pre + "q_norm.weight"].float(), W[pre + "q_norm.bias"].float(), 1e-5)
    xk = F.layer_norm(xk, (Hkv * hd,), W[pre + "k_norm.weight"].float(), W[pre + "k_norm.bias"].float(), 1e-5)
    xq = p.r(apply_rope(xq.view(B, N, H, hd), ang))
    xk = p.r(apply_rope(xk.view(B, N, Hkv, hd), ang))
    xv = xv.view(B, N, Hkv, hd)
    rep = H // Hkv
    q = xq.permute(0, 2, 1, 3)
    k = xk.repeat_interleave(rep, dim=2).permute(0, 2, 1, 3)
    v = xv.repeat_interleave(rep, dim=2).permute(0, 2, 1, 3)
    out = O._sdpa(p, q, k, v, softmax_scale, None)
    # region-masked caption cross-attention (:421-446)
    num_y = y.shape[0]
    q2 = torch.cat([q[0:1].repeat(num_y - 1, 1, 1, 1), q[-1:]], dim=0)                  # :423
    yk = p.linear(y, W[pre + "wk_y.weight"])
    yk = F.layer_norm(yk, (Hkv * hd,), W[pre + "ky_norm.weight"].float(), W[pre + "ky_norm.bias"].float(), 1e-5)
    yk = p.r(yk).view(num_y, -1, Hkv, hd).repeat_interleave(rep, dim=2).permute(0, 2, 1, 3)
    yv = p.linear(y, W[pre + "wv_y.weight"]).view(num_y, -1, Hkv, hd).repeat_interleave(rep, dim=2).permute(0, 2, 1, 3)
    m = y_mask[:, None, None, :] & rmask[:, None, :, None]                              # :426-431 -> [num_y, 1, N, T]
    out_y = torch.nan_to_num(_sdpa_full_mask(p, q2, yk, yv, 1.0 / math.sqrt(hd), m))    # default SDPA scale; :442
    gate = p.r(torch.tanh(p.r(W[pre + "gate"].float())))
    out_y = p.r(out_y * gate.view(1, -1, 1, 1))                                         # :443
    out_y = torch.cat([p.r(out_y[:-1].sum(dim=0, keepdim=True)), out_y[-1:]], dim=0)    # :444-446 (bf16 sum: fp32 accumulate, one rounding)
    out = p.r(out + out_y)
    out = out.permute(0, 2, 1, 3).reshape(B, N, H * hd)
    return p.linear(out, W[pre + "wo.weight"])


def block(p: _Prec, cfg: NextDiTConfig, W: Dict[str, Tensor], i: int, x: Tensor, ang: Tensor, y: Tensor, y_mask: Tensor, rmask: Tensor,
          c: Tensor, softmax_scale: float) -> Tensor:
    """model.py:587-641 (adaln_input branch); identical to the base block except for the region mask."""
    pre = f"layers.{i}."
    mod = p.linear(p.r(F.silu(c)), W[pre + "adaLN_modulation.1.weight"], W[pre + "adaLN_modulation.1.bias"])
    scale_msa, gate_msa, scale_mlp, gate_mlp = mod.chunk(4, dim=1)
    yn = rms_norm(p, y, W[pre + "attention_y_norm.weight"], cfg.norm_eps)
    a = attention(p, cfg, W, pre + "attention.",
                  modulate(p, rms_norm(p, x, W[pre + "attention_norm1.weight"], cfg.norm_eps), scale_msa), ang, yn, y_mask, rmask, softmax_scale)
    x = p.r(x + p.r(p.r(torch.tanh(gate_msa)).unsqueeze(1) * rms_norm(p, a, W[pre + "attention_norm2.weight"], cfg.norm_eps)))
    f = feed_forward(p, W, pre + "feed_forward.", modulate(p, rms_norm(p, x, W[pre + "ffn_norm1.weight"], cfg.norm_eps), scale_mlp))
    x = p.r(x + p.r(p.r(torch.tanh(gate_mlp)).unsqueeze(1) * rms_norm(p, f, W[pre + "ffn_norm2.weight"], cfg.norm_eps)))
    return x


def forward(cfg: NextDiTConfig, W: Dict[str, Tensor], x: Tensor, t: Tensor, cap_feats: Tensor, cap_mask: Tensor, global_cap_feats: Tensor,
            global_cap_mask: Tensor, h_split_num: int = 1, w_split_num: int = 1, *, scale_factor: float = 1.0, scale_watershed: float = 1.0,
            rope_timestep: float = 1.0, base_seqlen: Optional[int] = None, proportional_attn: bool = False, precision: str = "fp32") -> Tensor:
    """model.py:852-899.  x [2, C, H, W]; cap_feats [num_y, T, C] (region captions + the unconditional one); global caption [1, Tg, C]."""
    p = _Prec(precision)
    ps = cfg.patch_size
    B, C, H, Wd = x.shape
    hp, wp = H // ps, Wd // ps
    N = hp * wp
    x = p.r(x.float())
    X = p.linear(patchify(x, ps), W["x_embedder.weight"], W["x_embedder.bias"])
    ang = rope_angles(cfg.head_dim, hp, wp, scale_factor, scale_watershed, rope_timestep)
    temb = p.r(timestep_embedding(t))
    temb = p.linear(temb, W["t_embedder.mlp.0.weight"], W["t_embedder.mlp.0.bias"])
    temb = p.linear(p.r(F.silu(temb)), W["t_embedder.mlp.2.weight"], W["t_embedder.mlp.2.bias"])
    g = p.r(global_cap_feats.float())
    m = global_cap_mask.float().unsqueeze(-1)
    pool = p.r((g * m).sum(dim=1) / m.sum(dim=1))                                       # :866-868, [1, C]
    pool = F.layer_norm(pool, (cfg.cap_feat_dim,), W["cap_embedder.0.weight"].float(), W["cap_embedder.0.bias"].float(), 1e-5)
    cap_emb = p.linear(pool, W["cap_embedder.1.weight"], W["cap_embedder.1.bias"])
    c = p.r(temb + cap_emb)                                                             # [2, cd] + [1, cd]
    rmask = region_mask(cap_feats.shape[0], hp, wp, h_split_num, w_split_num)
    if proportional_attn:
        assert base_seqlen is not None
        softmax_scale = math.sqrt(math.log(N, base_seqlen) / cfg.head_dim)
    else:
        softmax_scale = math.sqrt(1.0 / cfg.head_dim)
    cap = p.r(cap_feats.float())
    ymask = cap_mask.bool()
    for i in range(cfg.n_layers):
        X = block(p, cfg, W, i, X, ang, cap, ymask, rmask, c, softmax_scale)
    scale = p.linear(p.r(F.silu(c)), W["final_layer.adaLN_modulation.1.weight"], W["final_layer.adaLN_modulation.1.bias"])
    Xn = F.layer_norm(X, (cfg.dim,), None, None, 1e-6)
    Xn = Xn * p.r(1.0 + scale).unsqueeze(1)
    Oo = p.linear(Xn, W["final_layer.linear.weight"], W["final_layer.linear.bias"])
    out = unpatchify(Oo, H, Wd, ps, cfg.out_channels)
    if cfg.learn_sigma:
        out = out[:, : cfg.in_channels]
    return out


def forward_with_cfg(cfg: NextDiTConfig, W: Dict[str, Tensor], x: Tensor, t: Tensor, cap_feats: Tensor, cap_mask: Tensor, cfg_scale: float,
                     scale_factor: float = 1.0, scale_watershed: float = 1.0, base_seqlen: Optional[int] = None, proportional_attn: bool = False,
                     global_cap_feats: Optional[Tensor] = None, global_cap_mask: Optional[Tensor] = None, h_split_num: int = 1,
                     w_split_num: int = 1, precision: str = "fp32") -> Tensor:
    """model.py:902-953, incl. the 3-channel CFG quirk."""
    p = _Prec(precision)
    half = x[: len(x) // 2]
    combined = torch.cat([half, half], dim=0)
    out = forward(cfg, W, combined, t, cap_feats, cap_mask, global_cap_feats, global_cap_mask, h_split_num, w_split_num,
                  scale_factor=scale_factor, scale_watershed=scale_watershed, rope_timestep=float(t[0]), base_seqlen=base_seqlen,
                  proportional_attn=proportional_attn, precision=precision)
    eps, rest = out[:, :3], out[:, 3:]
    cond, unc = torch.split(eps, len(eps) // 2, dim=0)
    half_eps = p.r(unc + p.r(cfg_scale * p.r(cond - unc)))
    return torch.cat([torch.cat([half_eps, half_eps], dim=0), rest], dim=1)


def synthetic_inputs(cfg: NextDiTConfig, hw, n_regions: int, T: int, seed: int):
    """One latent (repeated for cond / uncond), n_regions region captions of different valid lengths + the unconditional caption
    (short, like the padded empty prompt) + a global caption."""
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(1, cfg.in_channels, hw[0], hw[1], generator=g).to(torch.bfloat16).repeat(2, 1, 1, 1)
    cap = torch.randn(n_regions + 1, T, cfg.cap_feat_dim, generator=g).to(torch.bfloat16)
    mask = torch.zeros(n_regions + 1, T, dtype=torch.int32)
    for r in range(n_regions):
        mask[r, : max(1, T - 3 * r)] = 1
    mask[-1, : min(T, 8)] = 1
    gcap = torch.randn(1, T, cfg.cap_feat_dim, generator=g).to(torch.bfloat16)
    gmask = torch.ones(1, T, dtype=torch.int32)
    gmask[0, T - 2:] = 0
    return z, cap, mask, gcap, gmask
