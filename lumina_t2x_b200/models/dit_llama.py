"""Drop-in mirror of the reference's class-conditional Next-DiT (``Next-DiT-ImageNet/models/models.py:836-1056``,
``DiT_Llama`` with ``TransformerBlockSandwichNorm2`` blocks) on the same B200 engine: same constructor
arguments, factory names, state-dict keys and ``forward_with_cfg(x, t, y, cfg_scale, rope_scaling_factor,
ntk_factor)`` signature as used by ``Next-DiT-ImageNet/sample.py:168-186``.  head_dim must be 48, 72 or 96
(the 600M, 2B and 3B factories); the 7B factory (head_dim 128) is not supported by the attention kernel.

``moe`` selects the mixture-of-experts FFN of ``Next-DiT-MoE/models/`` (see ``lumina_t2x_b200.models.moe`` for the
factory names of that package): "time" (models.py), "space" (models1.py), "both" (models2.py)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from .. import _lib
from .nextdit import EngineModule, _FeedForward, _TimestepEmbedder, _Weight, _linear


class _Attention(nn.Module):
    def __init__(self, dim, n_heads, n_kv_heads, qk_norm):
        super().__init__()
        kv = (n_kv_heads or n_heads) * (dim // n_heads)
        self.wq, self.wk, self.wv, self.wo = _linear(dim, dim, False), _linear(dim, kv, False), _linear(dim, kv, False), _linear(dim, dim, False)
        if qk_norm:
            self.q_norm, self.k_norm = nn.LayerNorm(dim), nn.LayerNorm(kv)
        else:
            self.q_norm = self.k_norm = nn.Identity()


class _MoeLayer(nn.Module):
    """MoeLayer / TimeMoeLayer / SpaceMoeLayer (Next-DiT-MoE models.py:451-477, models2.py:451-506): parameter holder."""

    def __init__(self, dim, hidden, gate_in, num_experts):
        super().__init__()
        self.experts = nn.ModuleList([_FeedForward(dim, hidden) for _ in range(num_experts)])
        self.gate = nn.Linear(gate_in, num_experts, bias=False)


# moe kind -> ((module name, post-norm name, gate input "time"|"space", experts), ...)
_MOE_BLOCKS = {
    "": (("feed_forward", "ffn_norm", "", 0),),
    "time": (("feed_forward", "ffn_norm", "time", 8),),
    "space": (("feed_forward", "ffn_norm", "space", 8),),
    "both": (("feed_forward_time", "ffn_norm_time", "time", 4), ("feed_forward_space", "ffn_norm_space", "space", 4)),
}


class _Block(nn.Module):
    """TransformerBlockSandwichNorm2 (models.py:692-796): the pre-norms (PFRMSNorm) carry no parameters."""

    def __init__(self, dim, n_heads, n_kv_heads, hidden, qk_norm, moe=""):
        super().__init__()
        self.attention = _Attention(dim, n_heads, n_kv_heads, qk_norm)
        for name, norm, gate, E in _MOE_BLOCKS[moe]:
            ffn = _FeedForward(dim, hidden) if E == 0 else _MoeLayer(dim, hidden, min(dim, 1024) if gate == "time" else dim, E)
            setattr(self, name, ffn)
        self.attention_norm = _Weight(dim)
        for name, norm, gate, E in _MOE_BLOCKS[moe]:
            setattr(self, norm, _Weight(dim))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), _linear(min(dim, 1024), (2 + 2 * len(_MOE_BLOCKS[moe])) * dim, True, "zeros"))


class _LabelEmbedder(nn.Module):
    def __init__(self, num_classes, hidden, dropout_prob):
        super().__init__()
        self.embedding_table = nn.Embedding(num_classes + int(dropout_prob > 0), hidden)
        nn.init.normal_(self.embedding_table.weight, std=0.02)
        self.num_classes, self.dropout_prob = num_classes, dropout_prob


class _FinalLayer(nn.Module):
    def __init__(self, dim, patch_size, out_channels):
        super().__init__()
        self.linear = _linear(dim, patch_size * patch_size * out_channels, True, "zeros")
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), _linear(min(dim, 1024), 2 * dim, True, "zeros"))


class DiT_Llama(EngineModule):
    def __init__(self, input_size: int = 32, patch_size: int = 2, in_channels: int = 4, dim: int = 4096, n_layers: int = 32,
                 n_heads: int = 32, n_kv_heads: Optional[int] = None, multiple_of: int = 256,
                 ffn_dim_multiplier: Optional[float] = None, norm_eps: float = 1e-5, class_dropout_prob: float = 0.1,
                 num_classes: int = 1000, learn_sigma: bool = True, qk_norm: bool = False,
                 max_tokens: Optional[int] = None, max_batch: int = 2, moe: str = "") -> None:
        super().__init__()
        self.moe = moe
        self.learn_sigma, self.in_channels, self.input_size, self.patch_size = learn_sigma, in_channels, input_size, patch_size
        self.out_channels = in_channels * 2 if learn_sigma else in_channels
        self.dim, self.n_heads, self.n_layers = dim, n_heads, n_layers
        self.n_kv_heads = n_kv_heads or n_heads
        self.norm_eps, self.multiple_of, self.qk_norm, self.num_classes = norm_eps, multiple_of, qk_norm, num_classes
        self._ffn_dim_multiplier, self._class_dropout_prob = ffn_dim_multiplier, class_dropout_prob
        hidden = int(2 * (4 * dim) / 3)
        if ffn_dim_multiplier is not None:
            hidden = int(ffn_dim_multiplier * hidden)
        hidden = multiple_of * ((hidden + multiple_of - 1) // multiple_of)
        self.x_embedder = _linear(patch_size * patch_size * in_channels, dim, True)
        self.t_embedder = _TimestepEmbedder(min(dim, 1024))
        self.y_embedder = _LabelEmbedder(num_classes, min(dim, 1024), class_dropout_prob)
        self.layers = nn.ModuleList([_Block(dim, n_heads, n_kv_heads, hidden, qk_norm, moe) for _ in range(n_layers)])
        self.final_layer = _FinalLayer(dim, patch_size, self.out_channels)
        assert (dim // n_heads) % 4 == 0, "2d rope needs head dim to be divisible by 4"
        self._init_engine_state(max_tokens or max(256, (input_size // patch_size) ** 2), 0, max_batch)
        self._label_key = None

    def _check_supported(self) -> None:
        if not self.qk_norm:
            raise NotImplementedError("the B200 engine implements the qk_norm=True architecture")
        if self._ffn_dim_multiplier is not None:
            raise NotImplementedError("ffn_dim_multiplier is not supported by the B200 engine")
        if self._class_dropout_prob <= 0:
            raise NotImplementedError("the B200 engine expects the CFG null-class row (class_dropout_prob > 0)")
        if self.dim // self.n_heads not in (48, 72, 96):
            raise NotImplementedError("the B200 attention kernel supports head_dim 48, 72 and 96 (600M / 2B / 3B factories)")

    def _ndit_config(self):
        return _lib.NditConfig(self.dim, self.n_layers, self.n_heads, self.n_kv_heads, 0, self.in_channels, self.patch_size,
                               self.multiple_of, int(self.learn_sigma), float(self.norm_eps), self._limits[0], 0, self._limits[2],
                               self.num_classes, 0, *{"": (0, 0), "time": (8, 0), "space": (0, 8), "both": (4, 4)}[self.moe])

    def _set_labels(self, lib, h, y: torch.Tensor, stream):
        key = self._tensor_key(y)
        if key is not None and key == self._label_key:
            return
        yl = y.detach().to(torch.int64).contiguous()
        _lib.check(lib.ndit_set_labels(h, C.c_void_p(yl.data_ptr()), yl.numel(), stream), h)
        self._label_key, self._cap_keepalive = key, (y, yl)

    def _engine(self, device):
        fresh = self._handle is None or self._dirty
        lib, h = super()._engine(device)
        if fresh:
            self._label_key = None
        return lib, h

    def _step_params(self, cfg_scale, rope_scaling_factor, ntk_factor):
        """models.py:952-960: the kwargs overwrite self.freqs_cis, so an override stays in effect for later calls that pass
        None (the ctor's table is rope_scaling_factor = ntk_factor = 1)."""
        if rope_scaling_factor is not None or ntk_factor is not None:
            assert rope_scaling_factor is not None and ntk_factor is not None       # models.py:952-953
            self._rope_override = (float(rope_scaling_factor), float(ntk_factor))
        lin, ntk = getattr(self, "_rope_override", (1.0, 1.0))
        return _lib.NditStepParams(float(cfg_scale), lin, 1.0, 0, 0, ntk)

    @torch.no_grad()
    def forward(self, x, t, y):
        """models.py:920-944 (inference): x [N,C,H,W], t [N], y [N] labels -> first C output channels [N,C,H,W]; no guidance, one
        timestep and label per row, rows in groups of at most max_batch (ndit_forward).  Uses the rope factors the module currently
        holds (ctor: 1, 1; a forward_with_cfg call with explicit factors overwrites them, :952-960)."""
        self._check_inputs(x, y)
        lib, h = self._engine(x.device)
        lin, ntk = getattr(self, "_rope_override", (1.0, 1.0))
        sp = _lib.NditStepParams(0.0, lin, 1.0, 0, 0, ntk)
        n = x.shape[0]
        tv = (t.detach().float().reshape(-1).tolist() if isinstance(t, torch.Tensor) else [float(t)] * n)
        if len(tv) == 1:
            tv = tv * n
        if len(tv) != n or y.numel() != n:
            raise ValueError(f"t / y have {len(tv)} / {y.numel()} entries for a batch of {n}")
        xb = x.detach().to(torch.bfloat16).contiguous()
        out = torch.empty_like(xb)
        _, _, Hh, Ww = xb.shape
        with torch.cuda.device(x.device):
            stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            self._ensure_capacity(lib, h, self._tokens_for(Hh, Ww), 0, 1)
            step = self._limits[2]
            for i in range(0, n, step):
                j = min(n, i + step)
                self._label_key = None
                self._set_labels(lib, h, y[i:j], stream)
                ta = (C.c_float * (j - i))(*tv[i:j])
                _lib.check(lib.ndit_forward(h, C.c_void_p(xb[i:j].data_ptr()), ta, j - i, Hh, Ww, C.byref(sp),
                                            C.c_void_p(out[i:j].data_ptr()), stream), h)
            self._label_key = None
        return out.to(x.dtype)

    @torch.no_grad()
    def forward_with_cfg(self, x, t, y, cfg_scale, rope_scaling_factor=None, ntk_factor=None):
        """models.py:946-974.  x [2n,C,H,W] (first half cond, second half ignored on input); y [2n] labels."""
        self._check_inputs(x, y)
        lib, h = self._engine(x.device)
        with torch.cuda.device(x.device):
            stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            self._ensure_capacity(lib, h, self._tokens_for(x.shape[2], x.shape[3]), 0, x.shape[0])
            self._set_labels(lib, h, y, stream)
            return self._run_forward(lib, h, x, t, self._step_params(cfg_scale, rope_scaling_factor, ntk_factor))

    @torch.no_grad()
    def sample_fixed_grid(self, z, t_grid, method: str, y, cfg_scale, rope_scaling_factor=None, ntk_factor=None,
                          return_trajectory: bool = True):
        self._check_inputs(z, y)
        lib, h = self._engine(z.device)
        with torch.cuda.device(z.device):
            stream = C.c_void_p(torch.cuda.current_stream(z.device).cuda_stream)
            self._ensure_capacity(lib, h, self._tokens_for(z.shape[2], z.shape[3]), 0, z.shape[0])
            self._set_labels(lib, h, y, stream)
            return self._run_sample(lib, h, z, t_grid, method, self._step_params(cfg_scale, rope_scaling_factor, ntk_factor),
                                    return_trajectory)


def DiT_Llama_600M_patch2(**kwargs):
    """models.py:1042-1043."""
    return DiT_Llama(patch_size=2, dim=1536, n_layers=16, n_heads=32, **kwargs)


def DiT_Llama_2B_patch2(**kwargs):
    """models.py:1046-1047."""
    return DiT_Llama(patch_size=2, dim=2304, n_layers=24, n_heads=32, **kwargs)


def DiT_Llama_3B_patch2(**kwargs):
    """models.py:1050-1051 (head_dim 96)."""
    return DiT_Llama(patch_size=2, dim=3072, n_layers=32, n_heads=32, **kwargs)


def DiT_Llama_7B_patch2(**kwargs):
    """models.py:1054-1055.  head_dim 128: constructs (same state dict as the reference) but the engine refuses to run it -
    the attention kernel covers head_dim 48 / 72 / 96."""
    return DiT_Llama(patch_size=2, dim=4096, n_layers=32, n_heads=32, **kwargs)
