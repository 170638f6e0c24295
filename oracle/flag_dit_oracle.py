"""CPU oracle for the Flag-DiT text-to-image model (Lumina-T2I, BASELINE config 4).  TEST INFRASTRUCTURE ONLY.

Plain-PyTorch restatement of ``lumina_t2i/models/model.py`` ``DiT_Llama.forward_with_cfg`` (the 5B model of
SURVEY.md section 8 row a15).  Same rules as ``nextdit_oracle.py``: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline legs may import it; the product never does.

Pinned against the reference itself: ``oracle/make_golden.py::make_flag_dit`` imports the unmodified
``/root/reference/lumina_t2i/models/model.py`` (fairscale replaced by the world-size-1 stub of
``oracle/harness/ref_import.py``), strict-loads this file's synthetic weights and stores ``forward_with_cfg``
outputs under ``tests/golden/flagdit_*.pt``; ``tests/test_oracle_vs_golden.py`` compares.

What differs from Next-DiT (citations relative to /root/reference/lumina_t2i/models/model.py):
  modulate with shift                x*(1+scale)+shift                         :28-29
  block                              6-chunk adaLN (shift, scale, gate) x2, plain gate (no tanh), no
                                     post-norms, one RMSNorm per sub-block        :505-622
  attention                          same as Next-DiT (LN over all heads, gated caption cross-attention) but a
                                     1-D RoPE over the token index               :264-287, :347-443
  final layer                        LN(no affine) * (1+scale) + shift            :625-662
  patchify                           a learned [eol] token closes every row of patches; unpatchify drops it
                                                                                   :779-785, :745-755
  rope table                         freqs 1/(theta*ntk)^(2i/hd), angle = (pos / rope_scaling) * freq  :925-960
  proportional attention             scale sqrt(log_base(seqlen)/hd), seqlen counts the eol tokens   :365-368
  forward / forward_with_cfg         :833-866 / :868-923
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .nextdit_oracle import _Prec, _sdpa, apply_rope, feed_forward, patchify, rms_norm, timestep_embedding

Tensor = torch.Tensor


@dataclass
class FlagDiTConfig:
    """DiT_Llama ctor args (:665-680)."""

    dim: int = 3072
    n_layers: int = 32
    n_heads: int = 32
    n_kv_heads: int = 32
    cap_feat_dim: int = 4096
    patch_size: int = 2
    in_channels: int = 4
    multiple_of: int = 256
    norm_eps: float = 1e-5
    learn_sigma: bool = True

    @property
    def head_dim(self) -> int:
        return self.dim // self.n_heads

    @property
    def ffn_dim(self) -> int:
        h = int(2 * (4 * self.dim) / 3)                        # :477-481
        return self.multiple_of * ((h + self.multiple_of - 1) // self.multiple_of)

    @property
    def cond_dim(self) -> int:
        return min(self.dim, 1024)

    @property
    def out_channels(self) -> int:
        return self.in_channels * 2 if self.learn_sigma else self.in_channels


def config_5b() -> FlagDiTConfig:
    """DiT_Llama_5B_patch2 (:989-990) with the LLaMA-7B caption width of the released checkpoint."""
    return FlagDiTConfig()


def config_tiny96(n_layers: int = 2) -> FlagDiTConfig:
    """Small model with the flagship head_dim (96)."""
    return FlagDiTConfig(dim=384, n_layers=n_layers, n_heads=4, n_kv_heads=4, cap_feat_dim=256)


def rope_angles_1d(head_dim: int, n_tokens: int, rope_scaling_factor: float = 1.0, ntk_factor: float = 1.0,
                   theta: float = 10000.0) -> Tensor:
    """:925-960 -> [n_tokens, head_dim//2] float32 angles (the first n_tokens rows of the 40000-row table, :848)."""
    theta = theta * ntk_factor
    freqs = 1.0 / (theta ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    t = torch.arange(n_tokens, dtype=torch.float32) / rope_scaling_factor
    return torch.outer(t, freqs).float()


def modulate(p: _Prec, x: Tensor, shift: Tensor, scale: Tensor) -> Tensor:
    """:28-29; every intermediate is a bf16 tensor under autocast with bf16 parameters."""
    return p.r(p.r(x * p.r(1.0 + scale).unsqueeze(1)) + shift.unsqueeze(1))


def attention(p: _Prec, cfg: FlagDiTConfig, W: Dict[str, Tensor], pre: str, x: Tensor, ang: Tensor, y: Tensor,
              y_mask: Tensor, softmax_scale: float) -> Tensor:
    """:347-443."""
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
    out = _sdpa(p, q, k, v, softmax_scale, None)
    yk = p.linear(y, W[pre + "wk_y.weight"])
    yk = F.layer_norm(yk, (Hkv * hd,), W[pre + "ky_norm.weight"].float(), W[pre + "ky_norm.bias"].float(), 1e-5)
    yk = p.r(yk).view(B, -1, Hkv, hd).repeat_interleave(rep, dim=2).permute(0, 2, 1, 3)
    yv = p.linear(y, W[pre + "wv_y.weight"]).view(B, -1, Hkv, hd).repeat_interleave(rep, dim=2).permute(0, 2, 1, 3)
    out_y = _sdpa(p, q, yk, yv, 1.0 / math.sqrt(hd), y_mask)
    gate = p.r(torch.tanh(p.r(W[pre + "gate"].float())))
    out = p.r(out + p.r(out_y * gate.view(1, -1, 1, 1)))
    out = out.permute(0, 2, 1, 3).reshape(B, N, H * hd)
    return p.linear(out, W[pre + "wo.weight"])


def block(p: _Prec, cfg: FlagDiTConfig, W: Dict[str, Tensor], i: int, x: Tensor, ang: Tensor, y: Tensor, y_mask: Tensor,
          c: Tensor, softmax_scale: float) -> Tensor:
    """:586-622 (adaln_input branch)."""
    pre = f"layers.{i}."
    mod = p.linear(p.r(F.silu(c)), W[pre + "adaLN_modulation.1.weight"], W[pre + "adaLN_modulation.1.bias"])
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = mod.chunk(6, dim=1)
    yn = rms_norm(p, y, W[pre + "attention_y_norm.weight"], cfg.norm_eps)
    a = attention(p, cfg, W, pre + "attention.",
                  modulate(p, rms_norm(p, x, W[pre + "attention_norm.weight"], cfg.norm_eps), shift_msa, scale_msa),
                  ang, yn, y_mask, softmax_scale)
    x = p.r(x + p.r(gate_msa.unsqueeze(1) * a))
    f = feed_forward(p, W, pre + "feed_forward.",
                     modulate(p, rms_norm(p, x, W[pre + "ffn_norm.weight"], cfg.norm_eps), shift_mlp, scale_mlp))
    return p.r(x + p.r(gate_mlp.unsqueeze(1) * f))


def n_tokens(cfg: FlagDiTConfig, H: int, Wd: int) -> int:
    return (H // cfg.patch_size) * (Wd // cfg.patch_size + 1)


def forward(cfg: FlagDiTConfig, W: Dict[str, Tensor], x: Tensor, t: Tensor, cap_feats: Tensor, cap_mask: Tensor, *,
            rope_scaling_factor: float = 1.0, ntk_factor: float = 1.0, base_seqlen: Optional[int] = None,
            proportional_attn: bool = False, precision: str = "fp32", taps: Optional[dict] = None) -> Tensor:
    """:833-866."""
    p = _Prec(precision)
    ps = cfg.patch_size
    B, C, H, Wd = x.shape
    hp, wp = H // ps, Wd // ps
    x = p.r(x.float())
    X = p.linear(patchify(x, ps), W["x_embedder.weight"], W["x_embedder.bias"]).view(B, hp, wp, cfg.dim)
    eol = p.r(W["eol_token"].float()).view(1, 1, 1, -1).expand(B, hp, 1, -1)
    X = torch.cat([X, eol], dim=2).flatten(1, 2)                               # [B, hp*(wp+1), D]  (:779-785)
    N = X.shape[1]
    ang = rope_angles_1d(cfg.head_dim, N, rope_scaling_factor, ntk_factor)
    temb = p.r(timestep_embedding(t))
    temb = p.linear(temb, W["t_embedder.mlp.0.weight"], W["t_embedder.mlp.0.bias"])
    temb = p.linear(p.r(F.silu(temb)), W["t_embedder.mlp.2.weight"], W["t_embedder.mlp.2.bias"])
    cap = p.r(cap_feats.float())
    m = cap_mask.float().unsqueeze(-1)
    pool = p.r((cap * m).sum(dim=1) / m.sum(dim=1))
    pool = F.layer_norm(pool, (cfg.cap_feat_dim,), W["cap_embedder.0.weight"].float(), W["cap_embedder.0.bias"].float(), 1e-5)
    c = p.r(temb + p.linear(pool, W["cap_embedder.1.weight"], W["cap_embedder.1.bias"]))
    if proportional_attn:
        assert base_seqlen is not None
        softmax_scale = math.sqrt(math.log(N, base_seqlen) / cfg.head_dim)
    else:
        softmax_scale = math.sqrt(1.0 / cfg.head_dim)
    ymask = cap_mask.bool()
    if taps is not None:
        taps["x_embed"], taps["c"] = X.clone(), c.clone()
    for i in range(cfg.n_layers):
        X = block(p, cfg, W, i, X, ang, cap, ymask, c, softmax_scale)
        if taps is not None:
            taps[f"block{i}"] = X.clone()
    # final layer (:655-660): LN(no affine, fp32 under autocast) * (1+scale) + shift, then the bf16 Linear
    fm = p.linear(p.r(F.silu(c)), W["final_layer.adaLN_modulation.1.weight"], W["final_layer.adaLN_modulation.1.bias"])
    shift, scale = fm.chunk(2, dim=1)
    Xn = F.layer_norm(X, (cfg.dim,), None, None, 1e-6)
    Xn = Xn * p.r(1.0 + scale).unsqueeze(1) + shift.unsqueeze(1)
    O = p.linear(Xn, W["final_layer.linear.weight"], W["final_layer.linear.bias"])
    O = O.view(B, hp, wp + 1, ps, ps, cfg.out_channels)[:, :, :-1]             # drop the eol column (:745-755)
    out = O.permute(0, 5, 1, 3, 2, 4).flatten(4, 5).flatten(2, 3)
    if cfg.learn_sigma:
        out = out[:, : cfg.in_channels]
    return out


def forward_with_cfg(cfg: FlagDiTConfig, W: Dict[str, Tensor], x: Tensor, t: Tensor, cap_feats: Tensor, cap_mask: Tensor,
                     cfg_scale: float, rope_scaling_factor: Optional[float] = None, ntk_factor: Optional[float] = None,
                     base_seqlen: Optional[int] = None, proportional_attn: bool = False, precision: str = "fp32",
                     taps: Optional[dict] = None) -> Tensor:
    """:868-923 (ctor defaults rope_scaling_factor = ntk_factor = 1.0; 3-channel CFG)."""
    p = _Prec(precision)
    half = x[: len(x) // 2]
    out = forward(cfg, W, torch.cat([half, half], dim=0), t, cap_feats, cap_mask,
                  rope_scaling_factor=1.0 if rope_scaling_factor is None else rope_scaling_factor,
                  ntk_factor=1.0 if ntk_factor is None else ntk_factor, base_seqlen=base_seqlen,
                  proportional_attn=proportional_attn, precision=precision, taps=taps)
    eps, rest = out[:, :3], out[:, 3:]
    cond, unc = torch.split(eps, len(eps) // 2, dim=0)
    half_eps = p.r(unc + p.r(cfg_scale * p.r(cond - unc)))
    return torch.cat([torch.cat([half_eps, half_eps], dim=0), rest], dim=1)


# --------------------------------------------------------------------------- synthetic weights / inputs


def state_dict_shapes(cfg: FlagDiTConfig) -> Dict[str, tuple]:
    """Reference key names and shapes (pinned by make_golden's strict load)."""
    D, F_, C, cd = cfg.dim, cfg.ffn_dim, cfg.cap_feat_dim, cfg.cond_dim
    H, Hkv, hd = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim
    pp = cfg.patch_size * cfg.patch_size
    s = {
        "eol_token": (D,), "pad_token": (D,),
        "x_embedder.weight": (D, pp * cfg.in_channels), "x_embedder.bias": (D,),
        "t_embedder.mlp.0.weight": (cd, 256), "t_embedder.mlp.0.bias": (cd,),
        "t_embedder.mlp.2.weight": (cd, cd), "t_embedder.mlp.2.bias": (cd,),
        "cap_embedder.0.weight": (C,), "cap_embedder.0.bias": (C,),
        "cap_embedder.1.weight": (cd, C), "cap_embedder.1.bias": (cd,),
        "final_layer.linear.weight": (pp * cfg.out_channels, D), "final_layer.linear.bias": (pp * cfg.out_channels,),
        "final_layer.adaLN_modulation.1.weight": (2 * D, cd), "final_layer.adaLN_modulation.1.bias": (2 * D,),
    }
    for i in range(cfg.n_layers):
        a = f"layers.{i}.attention."
        s.update({
            a + "gate": (H,), a + "wq.weight": (H * hd, D), a + "wk.weight": (Hkv * hd, D), a + "wv.weight": (Hkv * hd, D),
            a + "wk_y.weight": (Hkv * hd, C), a + "wv_y.weight": (Hkv * hd, C), a + "wo.weight": (D, H * hd),
            a + "q_norm.weight": (H * hd,), a + "q_norm.bias": (H * hd,),
            a + "k_norm.weight": (Hkv * hd,), a + "k_norm.bias": (Hkv * hd,),
            a + "ky_norm.weight": (Hkv * hd,), a + "ky_norm.bias": (Hkv * hd,),
        })
        b = f"layers.{i}."
        s.update({
            b + "feed_forward.w1.weight": (F_, D), b + "feed_forward.w3.weight": (F_, D), b + "feed_forward.w2.weight": (D, F_),
            b + "attention_norm.weight": (D,), b + "ffn_norm.weight": (D,), b + "attention_y_norm.weight": (C,),
            b + "adaLN_modulation.1.weight": (6 * D, cd), b + "adaLN_modulation.1.bias": (6 * D,),
        })
    return s


def synthetic_weights(cfg: FlagDiTConfig, seed: int = 0, dtype: torch.dtype = torch.bfloat16) -> Dict[str, Tensor]:
    """Seeded weights with every branch live (the reference zero-inits adaLN / gate / final layer)."""
    g = torch.Generator().manual_seed(seed)
    W = {}
    for k, shp in state_dict_shapes(cfg).items():
        if k.endswith("norm.weight") or k.endswith("cap_embedder.0.weight"):
            w = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("attention.gate"):
            w = 0.5 * torch.randn(shp, generator=g)
        elif len(shp) == 2:
            std = 0.5 / math.sqrt(shp[1]) if "adaLN" in k else 1.0 / math.sqrt(shp[1])
            w = std * torch.randn(shp, generator=g)
        else:
            w = 0.05 * torch.randn(shp, generator=g)
        W[k] = w.to(dtype)
    return W


def synthetic_inputs(cfg: FlagDiTConfig, latent_hw=(128, 128), T: int = 128, uncond_len: int = 8, seed: int = 1,
                     dtype: torch.dtype = torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(1, cfg.in_channels, *latent_hw, generator=g).to(dtype).repeat(2, 1, 1, 1)
    cap = torch.randn(2, T, cfg.cap_feat_dim, generator=g).to(dtype)
    mask = torch.zeros(2, T, dtype=torch.int64)
    mask[0, :] = 1
    mask[1, :uncond_len] = 1
    return z, cap, mask
