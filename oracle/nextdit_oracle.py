"""CPU oracle for the Next-DiT denoising hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch restatement of the reference algorithm for the one
path this repo accelerates: ``transport`` fixed-grid ODE sampling driving
``NextDiT.forward_with_cfg``.  It is *not* the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``lumina_t2x_b200``) never
does, and fails loudly when its CUDA library is missing.

Parity pinning: the reference ships no tests, golden vectors or fixtures for this
path, so the oracle is pinned against *outputs of the
reference itself*: ``oracle/make_golden.py`` imports the unmodified
``/root/reference/lumina_next_t2i_mini/models/nextdit.py`` (CPU, fp32), runs it on
seeded inputs and stores the tensors under ``tests/golden/``;
``tests/test_oracle_vs_golden.py`` checks this restatement against them.

Reference citations (relative to /root/reference/lumina_next_t2i_mini/ unless
noted; the fairscale flavour lumina_next_t2i/models/model.py has the same math):

  timestep embedding        models/nextdit.py:61-87
  attention                 models/nextdit.py:318-397   (LN over all heads :330-331,
                            rope :232-262, proportional scale :341-344)
  feed forward              models/nextdit.py:467-472
  block                     models/nextdit.py:566-604
  final layer               models/nextdit.py:641-646
  patchify / unpatchify     models/nextdit.py:742-757 / :713-740
  forward                   models/nextdit.py:808-836
  forward_with_cfg          models/nextdit.py:838-885
  rope table                models/nextdit.py:887-928
  RMSNorm                   models/components.py:29-54
  ODE grid / time shift     transport.py:57-111, lumina_next_t2i/transport/integrators.py:79-116
  euler / midpoint step     torchdiffeq (third party, unpinned, absent from /root/reference):
                            fixed-grid solvers; in-tree restatement of midpoint at
                            visual_anagrams/generate.py:212-219

``precision``:
  "fp32"  - everything in float32 (the reference run without autocast).
  "bf16"  - float32 arithmetic with a round-to-bf16 at every point where the
            reference under ``torch.autocast("cuda", bf16)`` with bf16 parameters
            materialises a bf16 tensor.  This is what the
            CUDA engine is compared against: same rounding points, different
            (fp32) summation order inside each contraction.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class NextDiTConfig:
    """Architecture sizes (models/nextdit.py:613-629 ctor args)."""

    dim: int = 2304
    n_layers: int = 24
    n_heads: int = 32
    n_kv_heads: int = 8
    cap_feat_dim: int = 2048
    patch_size: int = 2
    in_channels: int = 4
    multiple_of: int = 256
    norm_eps: float = 1e-5
    learn_sigma: bool = True
    qk_norm: bool = True                            # False: q_norm / k_norm / ky_norm are nn.Identity (models/nextdit.py:166-175)
    ffn_dim_multiplier: Optional[float] = None      # models/nextdit.py:422-424

    @property
    def head_dim(self) -> int:
        return self.dim // self.n_heads

    @property
    def ffn_dim(self) -> int:
        # models/nextdit.py:421-425
        h = int(2 * (4 * self.dim) / 3)
        if self.ffn_dim_multiplier is not None:
            h = int(self.ffn_dim_multiplier * h)
        return self.multiple_of * ((h + self.multiple_of - 1) // self.multiple_of)

    @property
    def cond_dim(self) -> int:
        return min(self.dim, 1024)

    @property
    def out_channels(self) -> int:
        return self.in_channels * 2 if self.learn_sigma else self.in_channels


def config_2b_gqa() -> NextDiTConfig:
    """NextDiT_2B_GQA_patch2 (models/nextdit.py:943-944)."""
    return NextDiTConfig()


def config_tiny(n_layers: int = 2) -> NextDiTConfig:
    """Small config with the flagship head_dim (72) for fast parity tests."""
    return NextDiTConfig(dim=576, n_layers=n_layers, n_heads=8, n_kv_heads=2, cap_feat_dim=256)


# --------------------------------------------------------------------------- helpers


class _Prec:
    def __init__(self, precision: str):
        assert precision in ("fp32", "bf16")
        self.bf16 = precision == "bf16"

    def r(self, x: Tensor) -> Tensor:
        """Round to bf16 (kept as float32 storage) when emulating autocast."""
        return x.to(torch.bfloat16).float() if self.bf16 else x

    def linear(self, x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
        """autocast Linear: bf16 inputs, fp32 accumulate, one rounding of the output."""
        y = F.linear(self.r(x), self.r(w.float()), None if b is None else self.r(b.float()))
        return self.r(y)


def rms_norm(p: _Prec, x: Tensor, w: Tensor, eps: float) -> Tensor:
    """components.py:40,53-54: normalise in fp32, round, then multiply by weight."""
    x32 = x.float()
    n = x32 * torch.rsqrt(x32.pow(2).mean(-1, keepdim=True) + eps)
    return p.r(p.r(n) * p.r(w.float()))


def modulate(p: _Prec, x: Tensor, scale: Tensor) -> Tensor:
    """models/nextdit.py:25-26; (1+scale) is itself a bf16 tensor under autocast."""
    return p.r(x * p.r(1.0 + scale).unsqueeze(1))


def timestep_embedding(t: Tensor, dim: int = 256, max_period: float = 10000.0) -> Tensor:
    """models/nextdit.py:61-81 (t is used as-is, in [0,1]; no x1000)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def rope_angles(head_dim: int, hp: int, wp: int, scale_factor: float, scale_watershed: float,
                timestep: float, theta: float = 10000.0) -> Tensor:
    """Angles of the complex table of models/nextdit.py:887-928, restricted to the
    [hp, wp] grid actually used (:753).  Returns [hp*wp, head_dim//2] float32:
    complex index 2i -> row position * w_i, 2i+1 -> column position * w_i."""
    if timestep < scale_watershed:
        linear_factor, ntk_factor = scale_factor, 1.0
    else:
        linear_factor, ntk_factor = 1.0, scale_factor
    theta = theta * ntk_factor
    idx = torch.arange(0, head_dim, 4)[: head_dim // 4].float()
    freqs = 1.0 / (theta ** (idx / head_dim)) / linear_factor                 # [hd/4]
    ah = torch.outer(torch.arange(hp, dtype=torch.float32), freqs).float()    # [hp, hd/4]
    aw = torch.outer(torch.arange(wp, dtype=torch.float32), freqs).float()    # [wp, hd/4]
    ang = torch.stack([ah[:, None, :].expand(hp, wp, -1), aw[None, :, :].expand(hp, wp, -1)], dim=-1)
    return ang.flatten(2).flatten(0, 1)                                        # [hp*wp, hd/2]


def apply_rope(x: Tensor, ang: Tensor) -> Tensor:
    """models/nextdit.py:232-262: x [B,N,H,hd] fp32, pairs (2m,2m+1) rotated by ang[:,m]."""
    B, N, H, hd = x.shape
    xr = x.float().reshape(B, N, H, hd // 2, 2)
    c, s = torch.cos(ang)[None, :, None, :], torch.sin(ang)[None, :, None, :]
    re = xr[..., 0] * c - xr[..., 1] * s
    im = xr[..., 0] * s + xr[..., 1] * c
    return torch.stack([re, im], dim=-1).flatten(3)


def _sdpa(p: _Prec, q: Tensor, k: Tensor, v: Tensor, scale: float, mask: Optional[Tensor]) -> Tensor:
    """softmax(q k^T * scale + mask) v with bf16 q/k/v, fp32 softmax, bf16 output.
    q [B,H,N,hd], k/v [B,H,T,hd]; mask [B,T] bool or None."""
    if not p.bf16:
        # fp32 mode: the same fused CPU kernel the reference's non-flash branch calls (nextdit.py:358-373)
        am = None if mask is None else mask[:, None, None, :].expand(-1, q.shape[1], q.shape[2], -1)
        return F.scaled_dot_product_attention(q, k, v, attn_mask=am, scale=scale)
    s = torch.matmul(q, k.transpose(-1, -2)) * scale
    if mask is not None:
        s = s.masked_fill(~mask[:, None, None, :], float("-inf"))
    a = torch.softmax(s, dim=-1)
    return p.r(torch.matmul(p.r(a), v))       # flash kernels round P to bf16 before P.V


# --------------------------------------------------------------------------- model


def attention(p: _Prec, cfg: NextDiTConfig, W: Dict[str, Tensor], pre: str, x: Tensor, ang: Tensor,
              y: Tensor, y_mask: Tensor, softmax_scale: float) -> Tensor:
    """models/nextdit.py:318-397."""
    B, N, _ = x.shape
    H, Hkv, hd = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim
    xq = p.linear(x, W[pre + "wq.weight"])
    xk = p.linear(x, W[pre + "wk.weight"])
    xv = p.linear(x, W[pre + "wv.weight"])
    # LayerNorm over ALL heads jointly, fp32 output under autocast (:330-331)
    if cfg.qk_norm:
        xq = F.layer_norm(xq, (H * hd,), W[pre + "q_norm.weight"].float(), W[pre + "q_norm.bias"].float(), 1e-5)
        xk = F.layer_norm(xk, (Hkv * hd,), W[pre + "k_norm.weight"].float(), W[pre + "k_norm.bias"].float(), 1e-5)
    xq = p.r(apply_rope(xq.view(B, N, H, hd), ang))           # rope fp32 then .to(dtype) (:337-340)
    xk = p.r(apply_rope(xk.view(B, N, Hkv, hd), ang))
    xv = xv.view(B, N, Hkv, hd)
    rep = H // Hkv
    q = xq.permute(0, 2, 1, 3)
    k = xk.repeat_interleave(rep, dim=2).permute(0, 2, 1, 3)
    v = xv.repeat_interleave(rep, dim=2).permute(0, 2, 1, 3)
    out = _sdpa(p, q, k, v, softmax_scale, None)               # x_mask is all ones for tensor input
    # gated cross-attention to the caption tokens, reusing the RoPE'd q (:381-394)
    yk = p.linear(y, W[pre + "wk_y.weight"])
    if cfg.qk_norm:
        yk = F.layer_norm(yk, (Hkv * hd,), W[pre + "ky_norm.weight"].float(), W[pre + "ky_norm.bias"].float(), 1e-5)
    yk = p.r(yk).view(B, -1, Hkv, hd).repeat_interleave(rep, dim=2).permute(0, 2, 1, 3)
    yv = p.linear(y, W[pre + "wv_y.weight"]).view(B, -1, Hkv, hd).repeat_interleave(rep, dim=2).permute(0, 2, 1, 3)
    out_y = _sdpa(p, q, yk, yv, 1.0 / math.sqrt(hd), y_mask)
    gate = p.r(torch.tanh(p.r(W[pre + "gate"].float())))
    out = p.r(out + p.r(out_y * gate.view(1, -1, 1, 1)))
    out = out.permute(0, 2, 1, 3).reshape(B, N, H * hd)
    return p.linear(out, W[pre + "wo.weight"])


def feed_forward(p: _Prec, W: Dict[str, Tensor], pre: str, x: Tensor) -> Tensor:
    """models/nextdit.py:467-472."""
    x1 = p.linear(x, W[pre + "w1.weight"])
    x3 = p.linear(x, W[pre + "w3.weight"])
    h = p.r(p.r(F.silu(x1)) * x3)
    return p.linear(h, W[pre + "w2.weight"])


def block(p: _Prec, cfg: NextDiTConfig, W: Dict[str, Tensor], i: int, x: Tensor, ang: Tensor, y: Tensor,
          y_mask: Tensor, c: Tensor, softmax_scale: float) -> Tensor:
    """models/nextdit.py:566-604 (adaln_input branch)."""
    pre = f"layers.{i}."
    mod = p.linear(p.r(F.silu(c)), W[pre + "adaLN_modulation.1.weight"], W[pre + "adaLN_modulation.1.bias"])
    scale_msa, gate_msa, scale_mlp, gate_mlp = mod.chunk(4, dim=1)
    yn = rms_norm(p, y, W[pre + "attention_y_norm.weight"], cfg.norm_eps)
    a = attention(p, cfg, W, pre + "attention.",
                  modulate(p, rms_norm(p, x, W[pre + "attention_norm1.weight"], cfg.norm_eps), scale_msa),
                  ang, yn, y_mask, softmax_scale)
    x = p.r(x + p.r(p.r(torch.tanh(gate_msa)).unsqueeze(1) * rms_norm(p, a, W[pre + "attention_norm2.weight"], cfg.norm_eps)))
    f = feed_forward(p, W, pre + "feed_forward.",
                     modulate(p, rms_norm(p, x, W[pre + "ffn_norm1.weight"], cfg.norm_eps), scale_mlp))
    x = p.r(x + p.r(p.r(torch.tanh(gate_mlp)).unsqueeze(1) * rms_norm(p, f, W[pre + "ffn_norm2.weight"], cfg.norm_eps)))
    return x


def patchify(x: Tensor, ps: int) -> Tensor:
    """models/nextdit.py:746-750: token feature order (c, ph, pw)."""
    B, C, H, Wd = x.shape
    return x.view(B, C, H // ps, ps, Wd // ps, ps).permute(0, 2, 4, 1, 3, 5).flatten(3).flatten(1, 2)


def unpatchify(x: Tensor, H: int, Wd: int, ps: int, out_ch: int) -> Tensor:
    """models/nextdit.py:719-725: token feature order (ph, pw, c)."""
    B = x.shape[0]
    x = x.view(B, H // ps, Wd // ps, ps, ps, out_ch)
    return x.permute(0, 5, 1, 3, 2, 4).flatten(4, 5).flatten(2, 3)


def forward(cfg: NextDiTConfig, W: Dict[str, Tensor], x: Tensor, t: Tensor, cap_feats: Tensor, cap_mask: Tensor,
            *, scale_factor: float = 1.0, scale_watershed: float = 1.0, rope_timestep: float = 1.0,
            base_seqlen: Optional[int] = None, proportional_attn: bool = False,
            precision: str = "fp32", taps: Optional[dict] = None, seqlen_for_scale: Optional[int] = None) -> Tensor:
    """NextDiT.forward (models/nextdit.py:808-836) with the rope table of the enclosing
    forward_with_cfg call (:846-852).  ``taps``: optional dict filled with intermediates.
    ``seqlen_for_scale``: sequence length the proportional-attention scale sees (the PADDED length of a list input)."""
    p = _Prec(precision)
    ps = cfg.patch_size
    B, C, H, Wd = x.shape
    N = (H // ps) * (Wd // ps)
    x = p.r(x.float())
    X = p.linear(patchify(x, ps), W["x_embedder.weight"], W["x_embedder.bias"])
    ang = rope_angles(cfg.head_dim, H // ps, Wd // ps, scale_factor, scale_watershed, rope_timestep)
    # conditioning vector (:816-821)
    temb = p.r(timestep_embedding(t))
    temb = p.linear(temb, W["t_embedder.mlp.0.weight"], W["t_embedder.mlp.0.bias"])
    temb = p.linear(p.r(F.silu(temb)), W["t_embedder.mlp.2.weight"], W["t_embedder.mlp.2.bias"])
    cap = p.r(cap_feats.float())
    m = cap_mask.float().unsqueeze(-1)
    pool = p.r((cap * m).sum(dim=1) / m.sum(dim=1))
    pool = F.layer_norm(pool, (cfg.cap_feat_dim,), W["cap_embedder.0.weight"].float(), W["cap_embedder.0.bias"].float(), 1e-5)
    cap_emb = p.linear(pool, W["cap_embedder.1.weight"], W["cap_embedder.1.bias"])
    c = p.r(temb + cap_emb)
    if proportional_attn:
        assert base_seqlen is not None
        softmax_scale = math.sqrt(math.log(seqlen_for_scale or N, base_seqlen) / cfg.head_dim)
    else:
        softmax_scale = math.sqrt(1.0 / cfg.head_dim)
    ymask = cap_mask.bool()
    if taps is not None:
        taps["x_embed"] = X.clone()
        taps["c"] = c.clone()
    for i in range(cfg.n_layers):
        X = block(p, cfg, W, i, X, ang, cap, ymask, c, softmax_scale)
        if taps is not None:
            taps[f"block{i}"] = X.clone()
    # final layer (:641-646): LN without affine in fp32, modulate in fp32, Linear -> bf16
    scale = p.linear(p.r(F.silu(c)), W["final_layer.adaLN_modulation.1.weight"], W["final_layer.adaLN_modulation.1.bias"])
    Xn = F.layer_norm(X, (cfg.dim,), None, None, 1e-6)
    Xn = Xn * p.r(1.0 + scale).unsqueeze(1)                   # (1+scale) is bf16; fp32 * bf16 -> fp32
    O = p.linear(Xn, W["final_layer.linear.weight"], W["final_layer.linear.bias"])
    out = unpatchify(O, H, Wd, ps, cfg.out_channels)
    if cfg.learn_sigma:
        out = out[:, : cfg.in_channels]
    return out


def forward_list(cfg: NextDiTConfig, W: Dict[str, Tensor], xs, t: Tensor, cap_feats: Tensor, cap_mask: Tensor, **kw):
    """NextDiT.forward with a list of latents of different sizes (models/nextdit.py:761-806 patchify_and_embed, :362-377 varlen
    attention): pad tokens are masked out of every valid token's attention and dropped at the end, so each image is an independent
    batch-1 forward - except that the proportional-attention scale sees the padded (longest) sequence length."""
    ps = cfg.patch_size
    lmax = max((v.shape[1] // ps) * (v.shape[2] // ps) for v in xs)
    return [forward(cfg, W, v.unsqueeze(0), t[i:i + 1], cap_feats[i:i + 1], cap_mask[i:i + 1], seqlen_for_scale=lmax, **kw)[0]
            for i, v in enumerate(xs)]


def forward_with_cfg(cfg: NextDiTConfig, W: Dict[str, Tensor], x: Tensor, t: Tensor, cap_feats: Tensor,
                     cap_mask: Tensor, cfg_scale: float, scale_factor: float = 1.0, scale_watershed: float = 1.0,
                     base_seqlen: Optional[int] = None, proportional_attn: bool = False,
                     precision: str = "fp32", taps: Optional[dict] = None) -> Tensor:
    """NextDiT.forward_with_cfg (models/nextdit.py:838-885), incl. the 3-channel CFG quirk."""
    p = _Prec(precision)
    half = x[: len(x) // 2]
    combined = torch.cat([half, half], dim=0)
    out = forward(cfg, W, combined, t, cap_feats, cap_mask, scale_factor=scale_factor,
                  scale_watershed=scale_watershed, rope_timestep=float(t[0]), base_seqlen=base_seqlen,
                  proportional_attn=proportional_attn, precision=precision, taps=taps)
    eps, rest = out[:, :3], out[:, 3:]
    cond, unc = torch.split(eps, len(eps) // 2, dim=0)
    half_eps = p.r(unc + p.r(cfg_scale * p.r(cond - unc)))
    eps = torch.cat([half_eps, half_eps], dim=0)
    return torch.cat([eps, rest], dim=1)


# --------------------------------------------------------------------------- eager bf16 (stock-PyTorch stand-in)


def forward_with_cfg_eager(cfg: NextDiTConfig, W: Dict[str, Tensor], x: Tensor, t: Tensor, cap_feats: Tensor, cap_mask: Tensor,
                           cfg_scale: float, scale_factor: float = 1.0, scale_watershed: float = 1.0,
                           base_seqlen: Optional[int] = None, proportional_attn: bool = False) -> Tensor:
    """The same forward_with_cfg written the way the reference RUNS it on a GPU: native bf16 tensors, cuBLAS Linears,
    fused SDPA (flash) attention with GQA, fp32 only where ``torch.autocast`` keeps fp32 (norm statistics, LayerNorm,
    RoPE) - i.e. eager PyTorch, one ATen kernel per op.  Used by ``bench.py`` as the same-GPU "stock CUDA path" baseline
    (the reference itself cannot travel to the GPU box).  W: bf16 tensors on x's device."""
    bf = torch.bfloat16
    dev = x.device
    ps, D, H, Hkv, hd = cfg.patch_size, cfg.dim, cfg.n_heads, cfg.n_kv_heads, cfg.head_dim
    half = x[: len(x) // 2]
    x = torch.cat([half, half], dim=0).to(bf)
    B, C, Hh, Wd = x.shape
    N = (Hh // ps) * (Wd // ps)

    def rms(v, w):
        vf = v.float()
        return (vf * torch.rsqrt(vf.pow(2).mean(-1, keepdim=True) + cfg.norm_eps)).to(bf) * w

    X = F.linear(patchify(x, ps), W["x_embedder.weight"], W["x_embedder.bias"])
    ang = rope_angles(hd, Hh // ps, Wd // ps, scale_factor, scale_watershed, float(t[0])).to(dev)
    cos, sin = torch.cos(ang)[None, :, None, :], torch.sin(ang)[None, :, None, :]

    def rope(v):                                     # [B,N,h,hd] fp32 in, bf16 out
        vr = v.reshape(B, N, -1, hd // 2, 2)
        re = vr[..., 0] * cos - vr[..., 1] * sin
        im = vr[..., 0] * sin + vr[..., 1] * cos
        return torch.stack([re, im], dim=-1).flatten(3).to(bf)

    temb = timestep_embedding(t.cpu()).to(dev, bf)
    temb = F.linear(F.silu(F.linear(temb, W["t_embedder.mlp.0.weight"], W["t_embedder.mlp.0.bias"])),
                    W["t_embedder.mlp.2.weight"], W["t_embedder.mlp.2.bias"])
    cap = cap_feats.to(bf)
    m = cap_mask.float().unsqueeze(-1)
    pool = ((cap * m).sum(dim=1) / m.sum(dim=1)).to(bf)
    pool = F.layer_norm(pool.float(), (cfg.cap_feat_dim,), W["cap_embedder.0.weight"].float(), W["cap_embedder.0.bias"].float(), 1e-5)
    c = temb + F.linear(pool.to(bf), W["cap_embedder.1.weight"], W["cap_embedder.1.bias"])
    scale_self = math.sqrt(math.log(N, base_seqlen) / hd) if proportional_attn else math.sqrt(1.0 / hd)
    ymask = cap_mask.bool()[:, None, None, :]
    sc = F.silu(c)
    for i in range(cfg.n_layers):
        pre = f"layers.{i}."
        a = pre + "attention."
        s_a, g_a, s_m, g_m = F.linear(sc, W[pre + "adaLN_modulation.1.weight"], W[pre + "adaLN_modulation.1.bias"]).chunk(4, dim=1)
        u = rms(X, W[pre + "attention_norm1.weight"]) * (1 + s_a.unsqueeze(1))
        q = F.layer_norm(F.linear(u, W[a + "wq.weight"]).float(), (H * hd,), W[a + "q_norm.weight"].float(), W[a + "q_norm.bias"].float(), 1e-5)
        k = F.layer_norm(F.linear(u, W[a + "wk.weight"]).float(), (Hkv * hd,), W[a + "k_norm.weight"].float(), W[a + "k_norm.bias"].float(), 1e-5)
        v = F.linear(u, W[a + "wv.weight"]).view(B, N, Hkv, hd)
        q = rope(q.view(B, N, H, hd)).permute(0, 2, 1, 3)
        k = rope(k.view(B, N, Hkv, hd)).permute(0, 2, 1, 3)
        o = F.scaled_dot_product_attention(q, k, v.permute(0, 2, 1, 3), scale=scale_self, enable_gqa=True)
        yn = rms(cap, W[pre + "attention_y_norm.weight"])
        yk = F.layer_norm(F.linear(yn, W[a + "wk_y.weight"]).float(), (Hkv * hd,), W[a + "ky_norm.weight"].float(),
                          W[a + "ky_norm.bias"].float(), 1e-5).to(bf).view(B, -1, Hkv, hd).permute(0, 2, 1, 3)
        yv = F.linear(yn, W[a + "wv_y.weight"]).view(B, -1, Hkv, hd).permute(0, 2, 1, 3)
        oy = F.scaled_dot_product_attention(q, yk, yv, attn_mask=ymask, enable_gqa=True)
        o = o + oy * torch.tanh(W[a + "gate"]).view(1, -1, 1, 1)
        o = F.linear(o.permute(0, 2, 1, 3).reshape(B, N, H * hd), W[a + "wo.weight"])
        X = X + torch.tanh(g_a).unsqueeze(1) * rms(o, W[pre + "attention_norm2.weight"])
        mm = rms(X, W[pre + "ffn_norm1.weight"]) * (1 + s_m.unsqueeze(1))
        f = pre + "feed_forward."
        ff = F.linear(F.silu(F.linear(mm, W[f + "w1.weight"])) * F.linear(mm, W[f + "w3.weight"]), W[f + "w2.weight"])
        X = X + torch.tanh(g_m).unsqueeze(1) * rms(ff, W[pre + "ffn_norm2.weight"])
    scale = F.linear(sc, W["final_layer.adaLN_modulation.1.weight"], W["final_layer.adaLN_modulation.1.bias"])
    Xn = F.layer_norm(X.float(), (D,), None, None, 1e-6) * (1 + scale.unsqueeze(1))
    O = F.linear(Xn.to(bf), W["final_layer.linear.weight"], W["final_layer.linear.bias"])
    out = unpatchify(O, Hh, Wd, ps, cfg.out_channels)
    if cfg.learn_sigma:
        out = out[:, : cfg.in_channels]
    eps, rest = out[:, :3], out[:, 3:]
    cond, unc = torch.split(eps, len(eps) // 2, dim=0)
    half_eps = unc + cfg_scale * (cond - unc)
    return torch.cat([torch.cat([half_eps, half_eps], dim=0), rest], dim=1)


# --------------------------------------------------------------------------- sampler


def time_grid(num_steps: int, time_shifting_factor: Optional[float] = None, t0: float = 0.0, t1: float = 1.0) -> Tensor:
    """transport.py:71-74 / integrators.py:97-99: S grid points (S-1 integration steps)."""
    t = torch.linspace(t0, t1, num_steps)
    if time_shifting_factor:
        t = t / (t + time_shifting_factor - time_shifting_factor * t)
    return t


def sample_ode(cfg: NextDiTConfig, W: Dict[str, Tensor], z: Tensor, cap_feats: Tensor, cap_mask: Tensor, *,
               num_steps: int, method: str = "euler", time_shifting_factor: Optional[float] = None,
               cfg_scale: float = 4.0, scale_factor: float = 1.0, scale_watershed: float = 1.0,
               base_seqlen: Optional[int] = None, proportional_attn: bool = False,
               precision: str = "fp32", max_calls: Optional[int] = None, velocity_fn=None,
               cpu_mul_scalar_quirk: bool = False) -> Tensor:
    """ODE.sample (transport.py:87-111) with torchdiffeq's fixed-grid euler / midpoint.

    torchdiffeq (0.2.x ``_PerturbFunc.forward``) casts the time handed to the model to the
    state dtype, so in "bf16" mode the model sees bf16-rounded t; the step ``dt`` (a 0-dim fp32 tensor) is
    cast to the state dtype by type promotion when multiplied with the bf16 velocity, and the state update
    ``y + dt*f`` rounds to the state dtype after each op.
    Returns [num_steps, *z.shape] (all grid states), like odeint."""
    p = _Prec(precision)
    grid = time_grid(num_steps, time_shifting_factor)
    y = p.r(z.float())
    sols = [y]
    calls = 0

    def f(tt: Tensor, yy: Tensor) -> Tensor:
        nonlocal calls
        calls += 1
        tm = p.r(tt.float())                                   # t.to(y.dtype)
        tv = torch.ones(yy.shape[0]) * tm
        if velocity_fn is not None:              # test hook: pin the stepping semantics with a toy model
            return p.r(velocity_fn(yy, tv))
        return forward_with_cfg(cfg, W, yy, tv, cap_feats, cap_mask, cfg_scale, scale_factor, scale_watershed,
                                base_seqlen, proportional_attn, precision)

    for ta, tb in zip(grid[:-1], grid[1:]):
        if max_calls is not None and calls >= max_calls:
            break
        dt = tb - ta
        # torchdiffeq multiplies by the 0-dim *tensor* dt; with a bf16 state PyTorch's type promotion casts that
        # tensor to bf16 before the multiply, so the step size itself is bf16-rounded (identity in fp32 mode).
        if method == "euler":
            y = p.r(y + p.r(p.r(dt) * f(ta, y)))
        elif method == "midpoint":
            half_dt = 0.5 * dt
            # `f0 * half_dt` (tensor * 0-dim tensor): on CUDA both operands are cast to bf16; ATen's CPU mul kernel
            # instead keeps the fp32 scalar when it is the SECOND operand (BinaryOpsKernel.cpp mul_kernel), which
            # is what the CPU-generated fixture tests/golden/toy_midpoint_bf16.pt contains.
            hd_ = half_dt if cpu_mul_scalar_quirk else p.r(half_dt)
            ymid = p.r(y + p.r(f(ta, y) * hd_))
            y = p.r(y + p.r(p.r(dt) * f(ta + half_dt, ymid)))
        else:
            raise ValueError(f"oracle supports euler/midpoint, got {method}")
        sols.append(y)
    return torch.stack(sols, dim=0)


# --------------------------------------------------------------------------- synthetic weights


def state_dict_shapes(cfg: NextDiTConfig) -> Dict[str, tuple]:
    """Reference state-dict keys and shapes."""
    D, Hkv, hd, C, Fh, cd = cfg.dim, cfg.n_kv_heads, cfg.head_dim, cfg.cap_feat_dim, cfg.ffn_dim, cfg.cond_dim
    s: Dict[str, tuple] = {
        "pad_token": (D,),
        "x_embedder.weight": (D, cfg.patch_size ** 2 * cfg.in_channels), "x_embedder.bias": (D,),
        "t_embedder.mlp.0.weight": (cd, 256), "t_embedder.mlp.0.bias": (cd,),
        "t_embedder.mlp.2.weight": (cd, cd), "t_embedder.mlp.2.bias": (cd,),
        "cap_embedder.0.weight": (C,), "cap_embedder.0.bias": (C,),
        "cap_embedder.1.weight": (cd, C), "cap_embedder.1.bias": (cd,),
        "final_layer.linear.weight": (cfg.patch_size ** 2 * cfg.out_channels, D),
        "final_layer.linear.bias": (cfg.patch_size ** 2 * cfg.out_channels,),
        "final_layer.adaLN_modulation.1.weight": (D, cd), "final_layer.adaLN_modulation.1.bias": (D,),
    }
    for i in range(cfg.n_layers):
        a = f"layers.{i}.attention."
        s[a + "gate"] = (cfg.n_heads,)
        s[a + "wq.weight"] = (D, D)
        s[a + "wk.weight"] = (Hkv * hd, D)
        s[a + "wv.weight"] = (Hkv * hd, D)
        s[a + "wk_y.weight"] = (Hkv * hd, C)
        s[a + "wv_y.weight"] = (Hkv * hd, C)
        s[a + "wo.weight"] = (D, D)
        for n, w in (("q_norm", D), ("k_norm", Hkv * hd), ("ky_norm", Hkv * hd)) if cfg.qk_norm else ():
            s[a + n + ".weight"] = (w,)
            s[a + n + ".bias"] = (w,)
        f = f"layers.{i}.feed_forward."
        s[f + "w1.weight"] = (Fh, D)
        s[f + "w2.weight"] = (D, Fh)
        s[f + "w3.weight"] = (Fh, D)
        for n in ("attention_norm1", "attention_norm2", "ffn_norm1", "ffn_norm2"):
            s[f"layers.{i}.{n}.weight"] = (D,)
        s[f"layers.{i}.attention_y_norm.weight"] = (C,)
        s[f"layers.{i}.adaLN_modulation.1.weight"] = (4 * D, cd)
        s[f"layers.{i}.adaLN_modulation.1.bias"] = (4 * D,)
    return s


def synthetic_weights(cfg: NextDiTConfig, seed: int = 0, dtype: torch.dtype = torch.bfloat16) -> Dict[str, Tensor]:
    """Deterministic random weights with the reference's key names.  Every tensor the
    reference zero-initialises (adaLN, final layer, cap_embedder, gates: nextdit.py:152,
    :553-554, :628-639, :667-668) is drawn non-zero instead, otherwise the output is 0.
    Scales are chosen so activations stay O(1) through the stack (xavier-like)."""
    g = torch.Generator().manual_seed(seed)
    W: Dict[str, Tensor] = {}
    for k, shp in state_dict_shapes(cfg).items():
        if len(shp) == 2:
            std = 1.0 / math.sqrt(shp[1])
            if "adaLN" in k:
                std *= 0.5
            w = torch.randn(shp, generator=g) * std
        elif k.endswith("norm.weight") or k.endswith("norm1.weight") or k.endswith("norm2.weight") \
                or k == "cap_embedder.0.weight":
            w = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith(".gate"):
            w = 0.5 * torch.randn(shp, generator=g)
        else:  # biases, pad_token
            w = 0.02 * torch.randn(shp, generator=g)
        W[k] = w.to(dtype)
    return W


def synthetic_inputs(cfg: NextDiTConfig, latent_hw=(128, 128), T: int = 128, uncond_len: int = 8, seed: int = 1,
                     dtype: torch.dtype = torch.bfloat16):
    """Noise / caption inputs of SURVEY.md section 8(d): one latent repeated for the cond/uncond
    pair, caption features random, uncond mask = first ``uncond_len`` tokens."""
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(1, cfg.in_channels, latent_hw[0], latent_hw[1], generator=g).to(dtype).repeat(2, 1, 1, 1)
    g2 = torch.Generator().manual_seed(seed + 1)
    cap = torch.randn(2, T, cfg.cap_feat_dim, generator=g2).to(dtype)
    mask = torch.zeros(2, T, dtype=torch.int32)
    mask[0, :] = 1
    mask[1, :uncond_len] = 1
    return z, cap, mask
