"""Drop-in mirror of ``lumina_t2i.models`` (Lumina-T2I, the Flag-DiT family) for the sampling hot path.

Same constructor arguments, factory name, state-dict keys / shapes and ``forward_with_cfg`` signature as
``lumina_t2i/models/model.py:661-991`` (``DiT_Llama``, ``DiT_Llama_5B_patch2``), so ``lumina_t2i/demo.py`` /
``lumina_t2i/sample.py`` keep working with ``import lumina_t2x_b200.models.lumina_t2i as models``.  The module only
holds parameters; all compute is done by libndit_b200.so (``ndit_config.flag_dit = 1``).  No PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from .. import _lib
from .nextdit import EngineModule, NextDiT, _Attention, _FeedForward, _TimestepEmbedder, _Weight, _linear

__all__ = ["DiT_Llama", "DiT_Llama_5B_patch2"]


class _Block(nn.Module):
    """model.py:505-565: one weighted RMSNorm per sub-block, 6-chunk adaLN."""

    def __init__(self, dim, n_heads, n_kv_heads, hidden, qk_norm, y_dim):
        super().__init__()
        self.attention = _Attention(dim, n_heads, n_kv_heads, qk_norm, y_dim)
        self.feed_forward = _FeedForward(dim, hidden)
        self.attention_norm, self.ffn_norm = _Weight(dim), _Weight(dim)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), _linear(min(dim, 1024), 6 * dim, True, "zeros"))
        self.attention_y_norm = _Weight(y_dim)


class _FinalLayer(nn.Module):
    """model.py:625-662: adaLN gives shift and scale."""

    def __init__(self, dim, patch_size, out_channels):
        super().__init__()
        self.linear = _linear(dim, patch_size * patch_size * out_channels, True, "zeros")
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), _linear(min(dim, 1024), 2 * dim, True, "zeros"))


class DiT_Llama(NextDiT):
    """B200 engine behind the reference Flag-DiT ``DiT_Llama`` API (model.py:661-991)."""

    def __init__(self, patch_size: int = 2, in_channels: int = 4, dim: int = 4096, n_layers: int = 32, n_heads: int = 32,
                 n_kv_heads: Optional[int] = None, multiple_of: int = 256, ffn_dim_multiplier: Optional[float] = None,
                 norm_eps: float = 1e-5, learn_sigma: bool = True, qk_norm: bool = False, cap_feat_dim: int = 5120,
                 rope_scaling_factor: float = 1.0, ntk_factor: float = 1.0,
                 max_tokens: int = 4160, max_cap_len: int = 256, max_batch: int = 2) -> None:
        EngineModule.__init__(self)
        self.learn_sigma, self.in_channels, self.patch_size = learn_sigma, in_channels, patch_size
        self.out_channels = in_channels * 2 if learn_sigma else in_channels
        self.dim, self.n_heads, self.n_layers = dim, n_heads, n_layers
        self.n_kv_heads = n_kv_heads or n_heads
        self.cap_feat_dim, self.norm_eps, self.multiple_of = cap_feat_dim, norm_eps, multiple_of
        self.qk_norm = qk_norm
        self.rope_scaling_factor, self.ntk_factor = rope_scaling_factor, ntk_factor
        hidden = int(2 * (4 * dim) / 3)
        if ffn_dim_multiplier is not None:
            hidden = int(ffn_dim_multiplier * hidden)
        hidden = multiple_of * ((hidden + multiple_of - 1) // multiple_of)
        self.ffn_dim = hidden
        self._ffn_dim_multiplier = ffn_dim_multiplier
        self.x_embedder = _linear(patch_size * patch_size * in_channels, dim, True)
        self.t_embedder = _TimestepEmbedder(min(dim, 1024))
        self.cap_embedder = nn.Sequential(nn.LayerNorm(cap_feat_dim), _linear(cap_feat_dim, min(dim, 1024), True, "zeros"))
        self.layers = nn.ModuleList([_Block(dim, n_heads, n_kv_heads, hidden, qk_norm, cap_feat_dim) for _ in range(n_layers)])
        self.final_layer = _FinalLayer(dim, patch_size, self.out_channels)
        self.eol_token = nn.Parameter(torch.empty(dim))
        self.pad_token = nn.Parameter(torch.empty(dim))
        nn.init.normal_(self.eol_token, std=0.02)
        nn.init.normal_(self.pad_token, std=0.02)
        self._init_engine_state(max_tokens, max_cap_len, max_batch)

    def _ndit_config(self):
        return _lib.NditConfig(self.dim, self.n_layers, self.n_heads, self.n_kv_heads, self.cap_feat_dim, self.in_channels,
                               self.patch_size, self.multiple_of, int(self.learn_sigma), float(self.norm_eps), *self._limits, 0, 1)

    def _tokens_for(self, Hh: int, Ww: int) -> int:          # one learned [eol] token closes every row of patches
        return (Hh // self.patch_size) * (Ww // self.patch_size + 1)

    def _flag_step_params(self, cfg_scale, rope_scaling_factor, ntk_factor, base_seqlen, proportional_attn):
        """model.py:880-899: kwargs override the ctor's rope scaling / NTK factor, and the override is sticky."""
        if rope_scaling_factor is not None or ntk_factor is not None:
            self.rope_scaling_factor = rope_scaling_factor if rope_scaling_factor is not None else self.rope_scaling_factor
            self.ntk_factor = ntk_factor if ntk_factor is not None else self.ntk_factor
        if proportional_attn:
            assert base_seqlen is not None
        return _lib.NditStepParams(float(cfg_scale), float(self.rope_scaling_factor), 1.0, int(bool(proportional_attn)),
                                   int(base_seqlen) if base_seqlen is not None else 0, float(self.ntk_factor))

    @torch.no_grad()
    def forward_with_cfg(self, x, t, cap_feats, cap_mask, cfg_scale, rope_scaling_factor=None, ntk_factor=None,
                         base_seqlen: Optional[int] = None, proportional_attn: bool = False):
        """model.py:868-923.  x [2n,C,H,W]; first half = cond, second half ignored on input."""
        if not isinstance(x, torch.Tensor):
            raise NotImplementedError("list-of-tensors (variable resolution) input is not supported by the B200 engine")
        self._check_inputs(x, cap_feats, cap_mask)
        lib, h = self._engine(x.device)
        with torch.cuda.device(x.device):
            stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            self._ensure_capacity(lib, h, self._tokens_for(x.shape[2], x.shape[3]), cap_feats.shape[1], x.shape[0])
            self._set_caption(lib, h, cap_feats, cap_mask, stream)
            sp = self._flag_step_params(cfg_scale, rope_scaling_factor, ntk_factor, base_seqlen, proportional_attn)
            return self._run_forward(lib, h, x, t, sp)

    @torch.no_grad()
    def sample_fixed_grid(self, z, t_grid, method: str, cap_feats, cap_mask, cfg_scale, rope_scaling_factor=None, ntk_factor=None,
                          base_seqlen: Optional[int] = None, proportional_attn: bool = False, return_trajectory: bool = True):
        self._check_inputs(z, cap_feats, cap_mask)
        lib, h = self._engine(z.device)
        with torch.cuda.device(z.device):
            stream = C.c_void_p(torch.cuda.current_stream(z.device).cuda_stream)
            self._ensure_capacity(lib, h, self._tokens_for(z.shape[2], z.shape[3]), cap_feats.shape[1], z.shape[0])
            self._set_caption(lib, h, cap_feats, cap_mask, stream)
            sp = self._flag_step_params(cfg_scale, rope_scaling_factor, ntk_factor, base_seqlen, proportional_attn)
            return self._run_sample(lib, h, z, t_grid, method, sp, return_trajectory)


def DiT_Llama_5B_patch2(**kwargs):
    """model.py:989-990."""
    return DiT_Llama(patch_size=2, dim=3072, n_layers=32, n_heads=32, **kwargs)
