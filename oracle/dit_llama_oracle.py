"""CPU oracle for the class-conditional Next-DiT (BASELINE config 1 / SURVEY.md 8a14).  TEST INFRASTRUCTURE ONLY.

Restates ``DiT_Llama.forward_with_cfg`` of the reference's Next-DiT-ImageNet sub-project, reusing the shared pieces
of ``oracle/nextdit_oracle.py``.  Pinned by ``oracle/make_golden.py`` against the unmodified
``/root/reference/Next-DiT-ImageNet/models/models.py`` (imported with the world-size-1 fairscale stub of
``oracle/harness/ref_import.py``) -> ``tests/golden/imagenet_*.pt``.

Reference citations (relative to /root/reference/Next-DiT-ImageNet/):
  block (TransformerBlockSandwichNorm2)   models/models.py:759-796   (weight-free pre-norm PFRMSNorm :76-117,
                                          tanh-gated post-normed residual, 4-chunk adaLN)
  attention                               models/models.py:358-404   (LN over all heads, 2-D rope, default 1/sqrt(hd) scale)
  final layer (shift + scale)             models/models.py:829-833
  label embedding                         models/models.py:216-225
  forward / forward_with_cfg              models/models.py:920-974
  rope table (rope_scaling_factor, ntk)   models/models.py:977-1012

``precision="bf16"`` uses the same rounding points as the T2I oracle (autocast-like: LayerNorm outputs stay fp32),
which is how the CUDA engine computes; the reference's sample.py runs this model without autocast (fp32/tf32 by
default), so the fp32 mode is the one compared with the reference, the bf16 mode is the engine's contract.

Mixture-of-experts variants (BASELINE config 5 / SURVEY.md 8a16; /root/reference/Next-DiT-MoE/models/), ``cfg.moe``:
  "time"   models.py   MoeLayer :451-477 gated by the timestep embedding (8 experts, top-2): every token of a sample
                       uses the same two experts; block :663-771, forward passes time_input = t_embedder(t) :895-903
  "space"  models1.py  MoeLayer :451-477 gated per token by its own (modulated) input (8 experts, top-2)
  "both"   models2.py  TimeMoeLayer :451-477 then SpaceMoeLayer :480-506 (4 experts each), 6-chunk adaLN
                       (scale, gate) x {attention, time FFN, space FFN} :760-808
Expert outputs are accumulated in expert-index order into a zero tensor of the activation dtype
(``results[idx] += w * expert(x[idx])``), weights = softmax over the top-k logits in fp32, cast to the activation dtype.
Pinned by ``make_golden.make_moe`` -> ``tests/golden/moe_*.pt``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import nextdit_oracle as T

Tensor = torch.Tensor


@dataclass
class DiTLlamaConfig:
    dim: int = 1536
    n_layers: int = 16
    n_heads: int = 32
    n_kv_heads: Optional[int] = None
    num_classes: int = 1000
    patch_size: int = 2
    in_channels: int = 4
    multiple_of: int = 256
    norm_eps: float = 1e-5
    learn_sigma: bool = True
    moe: str = ""               # "", "time", "space", "both"

    @property
    def ffn_blocks(self):
        """(state-dict prefix, post-norm key, gate kind, experts) of the FFN sub-blocks of one layer, in execution order."""
        return {"": (("feed_forward", "ffn_norm", "", 0),),
                "time": (("feed_forward", "ffn_norm", "time", 8),),
                "space": (("feed_forward", "ffn_norm", "space", 8),),
                "both": (("feed_forward_time", "ffn_norm_time", "time", 4), ("feed_forward_space", "ffn_norm_space", "space", 4))}[self.moe]

    @property
    def kv_heads(self) -> int:
        return self.n_kv_heads or self.n_heads

    @property
    def head_dim(self) -> int:
        return self.dim // self.n_heads

    @property
    def ffn_dim(self) -> int:
        h = int(2 * (4 * self.dim) / 3)
        return self.multiple_of * ((h + self.multiple_of - 1) // self.multiple_of)

    @property
    def cond_dim(self) -> int:
        return min(self.dim, 1024)

    @property
    def out_channels(self) -> int:
        return self.in_channels * 2 if self.learn_sigma else self.in_channels


def config_600m() -> DiTLlamaConfig:
    """DiT_Llama_600M_patch2 (models/models.py:1042-1043): head_dim 48."""
    return DiTLlamaConfig()


def config_tiny48(n_layers: int = 2) -> DiTLlamaConfig:
    return DiTLlamaConfig(dim=384, n_layers=n_layers, n_heads=8, num_classes=10)


def config_tiny72(n_layers: int = 2) -> DiTLlamaConfig:
    return DiTLlamaConfig(dim=576, n_layers=n_layers, n_heads=8, num_classes=10)


def config_600m_moe(moe: str = "both") -> DiTLlamaConfig:
    """DiT_Llama_600M_patch2 / _Spatial / _Both of Next-DiT-MoE (models.py:1015, models1.py:1015, models2.py:1063)."""
    return DiTLlamaConfig(moe=moe)


def config_tiny_moe(moe: str, n_layers: int = 2) -> DiTLlamaConfig:
    return DiTLlamaConfig(dim=384, n_layers=n_layers, n_heads=8, num_classes=10, moe=moe)


def state_dict_shapes(cfg: DiTLlamaConfig) -> Dict[str, tuple]:
    D, KV, Fh, cd = cfg.dim, cfg.kv_heads * cfg.head_dim, cfg.ffn_dim, cfg.cond_dim
    po = cfg.patch_size ** 2
    s: Dict[str, tuple] = {
        "x_embedder.weight": (D, po * cfg.in_channels), "x_embedder.bias": (D,),
        "t_embedder.mlp.0.weight": (cd, 256), "t_embedder.mlp.0.bias": (cd,),
        "t_embedder.mlp.2.weight": (cd, cd), "t_embedder.mlp.2.bias": (cd,),
        "y_embedder.embedding_table.weight": (cfg.num_classes + 1, cd),
        "final_layer.linear.weight": (po * cfg.out_channels, D), "final_layer.linear.bias": (po * cfg.out_channels,),
        "final_layer.adaLN_modulation.1.weight": (2 * D, cd), "final_layer.adaLN_modulation.1.bias": (2 * D,),
    }
    for i in range(cfg.n_layers):
        a = f"layers.{i}.attention."
        s[a + "wq.weight"] = (D, D)
        s[a + "wk.weight"] = (KV, D)
        s[a + "wv.weight"] = (KV, D)
        s[a + "wo.weight"] = (D, D)
        for n, w in (("q_norm", D), ("k_norm", KV)):
            s[a + n + ".weight"] = (w,)
            s[a + n + ".bias"] = (w,)
        for name, norm, gate, E in cfg.ffn_blocks:
            f = f"layers.{i}.{name}."
            for pre in ([f] if E == 0 else [f"{f}experts.{j}." for j in range(E)]):
                s[pre + "w1.weight"] = (Fh, D)
                s[pre + "w2.weight"] = (D, Fh)
                s[pre + "w3.weight"] = (Fh, D)
            if E:
                s[f + "gate.weight"] = (E, cd if gate == "time" else D)
        nch = 2 + 2 * len(cfg.ffn_blocks)
        s[f"layers.{i}.attention_norm.weight"] = (D,)
        for name, norm, gate, E in cfg.ffn_blocks:
            s[f"layers.{i}.{norm}.weight"] = (D,)
        s[f"layers.{i}.adaLN_modulation.1.weight"] = (nch * D, cd)
        s[f"layers.{i}.adaLN_modulation.1.bias"] = (nch * D,)
    return s


def synthetic_weights(cfg: DiTLlamaConfig, seed: int = 0, dtype: torch.dtype = torch.bfloat16) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    W: Dict[str, Tensor] = {}
    for k, shp in state_dict_shapes(cfg).items():
        if k == "y_embedder.embedding_table.weight":
            w = 0.5 * torch.randn(shp, generator=g)
        elif len(shp) == 2:
            w = torch.randn(shp, generator=g) * ((0.5 if "adaLN" in k else 1.0) / math.sqrt(shp[1]))
        elif k.endswith("norm.weight"):
            w = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            w = 0.02 * torch.randn(shp, generator=g)
        W[k] = w.to(dtype)
    return W


def synthetic_inputs(cfg: DiTLlamaConfig, latent_hw=(32, 32), labels=(207,), seed: int = 1, dtype: torch.dtype = torch.bfloat16):
    """sample.py:168-183: z repeated for cond/uncond, y = [labels, num_classes * n]."""
    g = torch.Generator().manual_seed(seed)
    n = len(labels)
    z = torch.randn(n, cfg.in_channels, latent_hw[0], latent_hw[1], generator=g).to(dtype)
    z = torch.cat([z, z], 0)
    y = torch.tensor(list(labels) + [cfg.num_classes] * n, dtype=torch.int64)
    return z, y


def rope_angles(head_dim: int, hp: int, wp: int, rope_scaling_factor: float, ntk_factor: float, theta: float = 10000.0) -> Tensor:
    """models/models.py:977-1012 restricted to the [hp, wp] grid: positions / rope_scaling_factor, theta * ntk_factor."""
    return T.rope_angles(head_dim, hp, wp, rope_scaling_factor, 2.0, 0.0, theta=theta * ntk_factor)


def moe_layer(p, W: Dict[str, Tensor], pre: str, x: Tensor, cond: Optional[Tensor], E: int, top_k: int = 2) -> Tensor:
    """MoeLayer / TimeMoeLayer / SpaceMoeLayer.forward (Next-DiT-MoE models.py:459-477, models2.py:459-506).
    cond [B, cd] -> time gate (same logits for every token of a sample); cond None -> gate on the token itself."""
    B, N, D = x.shape
    xs = x.reshape(B * N, D)
    if cond is not None:
        logits = p.linear(cond, W[pre + "gate.weight"]).repeat(1, N).view(B * N, -1)
    else:
        logits = p.linear(xs, W[pre + "gate.weight"])
    wts, sel = torch.topk(logits, top_k)
    wts = p.r(torch.softmax(wts.float(), dim=1))
    res = torch.zeros_like(xs)
    for e in range(E):
        rows, nth = torch.where(sel == e)
        if rows.numel() == 0:
            continue
        oe = T.feed_forward(p, W, f"{pre}experts.{e}.", xs[rows])
        res[rows] = p.r(res[rows] + p.r(wts[rows, nth, None] * oe))
    return res.view(B, N, D)


def forward(cfg: DiTLlamaConfig, W: Dict[str, Tensor], x: Tensor, t: Tensor, y: Tensor,
            rope_scaling_factor: Optional[float] = None, ntk_factor: Optional[float] = None,
            precision: str = "fp32", taps: Optional[dict] = None) -> Tensor:
    """DiT_Llama.forward (Next-DiT-ImageNet/models/models.py:920-944): every row its own sample, timestep and label; returns the first
    C of the 2C output channels.  The rope table is the one the module currently holds (ctor: both factors 1; forward_with_cfg
    overwrites it, :952-960)."""
    p = T._Prec(precision)
    ps, H, hd, Hkv = cfg.patch_size, cfg.n_heads, cfg.head_dim, cfg.kv_heads
    x = p.r(x.float())
    B, C, Hh, Ww = x.shape
    N = (Hh // ps) * (Ww // ps)
    X = p.linear(T.patchify(x, ps), W["x_embedder.weight"], W["x_embedder.bias"])
    ang = rope_angles(hd, Hh // ps, Ww // ps, rope_scaling_factor or 1.0, ntk_factor or 1.0)
    temb = p.r(T.timestep_embedding(t))
    temb = p.linear(temb, W["t_embedder.mlp.0.weight"], W["t_embedder.mlp.0.bias"])
    temb = p.linear(p.r(F.silu(temb)), W["t_embedder.mlp.2.weight"], W["t_embedder.mlp.2.bias"])
    yemb = p.r(W["y_embedder.embedding_table.weight"].float()[y])
    c = p.r(temb + yemb)
    scale_attn = 1.0 / math.sqrt(hd)
    rep = H // Hkv
    for i in range(cfg.n_layers):
        pre = f"layers.{i}."
        mod = p.linear(p.r(F.silu(c)), W[pre + "adaLN_modulation.1.weight"], W[pre + "adaLN_modulation.1.bias"])
        chunks = mod.chunk(2 + 2 * len(cfg.ffn_blocks), dim=1)
        s_a, g_a = chunks[0], chunks[1]
        ones = torch.ones(cfg.dim)
        u = T.modulate(p, T.rms_norm(p, X, ones, cfg.norm_eps), s_a)                       # PFRMSNorm = unit weight
        a = pre + "attention."
        xq = p.linear(u, W[a + "wq.weight"])
        xk = p.linear(u, W[a + "wk.weight"])
        xv = p.linear(u, W[a + "wv.weight"]).view(B, N, Hkv, hd)
        xq = F.layer_norm(xq, (H * hd,), W[a + "q_norm.weight"].float(), W[a + "q_norm.bias"].float(), 1e-5)
        xk = F.layer_norm(xk, (Hkv * hd,), W[a + "k_norm.weight"].float(), W[a + "k_norm.bias"].float(), 1e-5)
        xq = p.r(T.apply_rope(xq.view(B, N, H, hd), ang)).permute(0, 2, 1, 3)
        xk = p.r(T.apply_rope(xk.view(B, N, Hkv, hd), ang)).repeat_interleave(rep, dim=2).permute(0, 2, 1, 3)
        v = xv.repeat_interleave(rep, dim=2).permute(0, 2, 1, 3)
        o = T._sdpa(p, xq, xk, v, scale_attn, None).permute(0, 2, 1, 3).reshape(B, N, H * hd)
        o = p.linear(o, W[a + "wo.weight"])
        X = p.r(X + p.r(p.r(torch.tanh(g_a)).unsqueeze(1) * T.rms_norm(p, o, W[pre + "attention_norm.weight"], cfg.norm_eps)))
        for bi, (name, norm, gate, E) in enumerate(cfg.ffn_blocks):
            s_m, g_m = chunks[2 + 2 * bi], chunks[3 + 2 * bi]
            m = T.modulate(p, T.rms_norm(p, X, ones, cfg.norm_eps), s_m)
            if E == 0:
                f = T.feed_forward(p, W, f"{pre}{name}.", m)
            else:
                f = moe_layer(p, W, f"{pre}{name}.", m, temb if gate == "time" else None, E)
            X = p.r(X + p.r(p.r(torch.tanh(g_m)).unsqueeze(1) * T.rms_norm(p, f, W[f"{pre}{norm}.weight"], cfg.norm_eps)))
        if taps is not None:
            taps[f"block{i}"] = X.clone()
    fin = p.linear(p.r(F.silu(c)), W["final_layer.adaLN_modulation.1.weight"], W["final_layer.adaLN_modulation.1.bias"])
    shift, scale = fin.chunk(2, dim=1)
    Xn = F.layer_norm(X, (cfg.dim,), None, None, 1e-6)
    Xn = Xn * p.r(1.0 + scale).unsqueeze(1) + shift.unsqueeze(1)
    O = p.linear(Xn, W["final_layer.linear.weight"], W["final_layer.linear.bias"])
    out = T.unpatchify(O, Hh, Ww, ps, cfg.out_channels)
    if cfg.learn_sigma:
        out = out[:, : cfg.in_channels]
    return out


def forward_with_cfg(cfg: DiTLlamaConfig, W: Dict[str, Tensor], x: Tensor, t: Tensor, y: Tensor, cfg_scale: float,
                     rope_scaling_factor: Optional[float] = None, ntk_factor: Optional[float] = None,
                     precision: str = "fp32", taps: Optional[dict] = None) -> Tensor:
    """DiT_Llama.forward_with_cfg (models.py:946-974): cond rows duplicated, guidance on the first three channels."""
    p = T._Prec(precision)
    half = x[: len(x) // 2]
    out = forward(cfg, W, torch.cat([half, half], dim=0), t, y, rope_scaling_factor, ntk_factor, precision, taps)
    eps, rest = out[:, :3], out[:, 3:]
    cond, unc = torch.split(eps, len(eps) // 2, dim=0)
    half_eps = p.r(unc + p.r(cfg_scale * p.r(cond - unc)))
    return torch.cat([torch.cat([half_eps, half_eps], dim=0), rest], dim=1)
