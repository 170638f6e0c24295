"""Drop-in mirror of the reference ``models.NextDiT`` for the sampling hot path.

Same constructor arguments, factory names, state-dict keys/shapes and ``forward_with_cfg`` signature as
``lumina_next_t2i/models/model.py:665-999`` (fairscale flavour) and
``lumina_next_t2i_mini/models/nextdit.py:607-944``, so ``sample.py`` / ``demo.py`` keep working with only
``sys.path`` changed.  The module only *holds* parameters (PyTorch owns them); all compute is done by
libndit_b200.so through the C ABI in include/ndit.h.  There is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import warnings
from typing import Optional

import torch
import torch.nn as nn

from .. import _lib


class _Weight(nn.Module):
    """Parameter holder for RMSNorm (reference: models/components.py:11-54)."""

    def __init__(self, dim: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))


def _linear(i: int, o: int, bias: bool, init: str = "xavier") -> nn.Linear:
    m = nn.Linear(i, o, bias=bias)
    if init == "xavier":
        nn.init.xavier_uniform_(m.weight)
    elif init == "zeros":
        nn.init.zeros_(m.weight)
    else:
        nn.init.normal_(m.weight, std=0.02)
    if bias:
        nn.init.zeros_(m.bias)
    return m


class _Attention(nn.Module):
    def __init__(self, dim, n_heads, n_kv_heads, qk_norm, y_dim):
        super().__init__()
        hd = dim // n_heads
        kv = (n_kv_heads or n_heads) * hd
        self.wq, self.wk, self.wv = _linear(dim, dim, False), _linear(dim, kv, False), _linear(dim, kv, False)
        self.wk_y, self.wv_y = _linear(y_dim, kv, False), _linear(y_dim, kv, False)
        self.gate = nn.Parameter(torch.zeros(n_heads))
        self.wo = _linear(dim, dim, False)
        if qk_norm:
            self.q_norm, self.k_norm, self.ky_norm = nn.LayerNorm(dim), nn.LayerNorm(kv), nn.LayerNorm(kv)
        else:
            self.q_norm = self.k_norm = self.ky_norm = nn.Identity()


class _FeedForward(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.w1, self.w2, self.w3 = _linear(dim, hidden, False), _linear(hidden, dim, False), _linear(dim, hidden, False)


class _Block(nn.Module):
    def __init__(self, dim, n_heads, n_kv_heads, hidden, qk_norm, y_dim):
        super().__init__()
        self.attention = _Attention(dim, n_heads, n_kv_heads, qk_norm, y_dim)
        self.feed_forward = _FeedForward(dim, hidden)
        self.attention_norm1, self.ffn_norm1 = _Weight(dim), _Weight(dim)
        self.attention_norm2, self.ffn_norm2 = _Weight(dim), _Weight(dim)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), _linear(min(dim, 1024), 4 * dim, True, "zeros"))
        self.attention_y_norm = _Weight(y_dim)


class _TimestepEmbedder(nn.Module):
    def __init__(self, hidden, freq=256):
        super().__init__()
        self.mlp = nn.Sequential(_linear(freq, hidden, True, "normal"), nn.SiLU(), _linear(hidden, hidden, True, "normal"))


class _FinalLayer(nn.Module):
    def __init__(self, dim, patch_size, out_channels):
        super().__init__()
        self.linear = _linear(dim, patch_size * patch_size * out_channels, True, "zeros")
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), _linear(min(dim, 1024), dim, True, "zeros"))


class EngineModule(nn.Module):
    """Parameter holder whose compute lives in libndit_b200.so: owns the engine handle, re-packs the weights
    after load_state_dict / .to(), and offers the option / instrumentation hooks shared by the model mirrors."""

    def _init_engine_state(self, max_tokens: int, max_cap_len: int, max_batch: int) -> None:
        self._limits = (max_tokens, max_cap_len, max_batch)
        self._handle: Optional[C.c_void_p] = None
        self._dirty = True
        self._cap_key = None
        self._cap_keepalive = None

    def _ndit_config(self) -> "_lib.NditConfig":          # pragma: no cover - provided by the subclasses
        raise NotImplementedError

    def _check_supported(self) -> None:
        pass

    def _apply(self, fn, *a, **k):           # .to() / .cuda() / .bfloat16() move the parameters
        self._dirty = True
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._dirty = True
        self._packed_path = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def mark_weights_changed(self) -> None:
        """Call after modifying parameters in place; the engine re-packs them on the next forward."""
        self._dirty = True

    def _destroy(self):
        if getattr(self, "_handle", None) is not None:
            _lib.load().ndit_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def reserve(self, max_tokens: Optional[int] = None, max_cap_len: Optional[int] = None, max_batch: Optional[int] = None):
        """Resize the engine workspace limits (recreates the engine on the next call)."""
        t, c, b = self._limits
        self._limits = (max_tokens or t, max_cap_len or c, max_batch or b)
        self._dirty = True

    def _engine(self, device: torch.device):
        lib = _lib.load()
        if self._handle is not None and not self._dirty:
            return lib, self._handle
        if getattr(self, "_packed_path", None) is not None:
            # the engine was created from a packed weight file (checkpoint.load_packed): the nn.Parameters of this module were
            # never populated, so a .to() / .cuda() re-creates the engine from the file, not from them
            from .. import checkpoint
            checkpoint.load_packed(self, self._packed_path, device)
            return lib, self._handle
        self._check_supported()
        if device.type != "cuda":
            raise RuntimeError(f"{type(self).__name__} (B200 engine) needs its parameters on a CUDA device; there is no CPU path")
        self._destroy()
        cfg = self._ndit_config()
        h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.ndit_create(C.byref(cfg), C.byref(h)), None)
            stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
            for key, p in self.state_dict().items():
                t = p.detach()
                if t.device != device:
                    raise RuntimeError(f"parameter {key} is on {t.device}, expected {device}")
                if t.dtype == torch.bfloat16:
                    dt = _lib.NDIT_BF16
                elif t.dtype == torch.float32:
                    dt = _lib.NDIT_F32
                else:
                    t, dt = t.float(), _lib.NDIT_F32
                t = t.contiguous()
                shape = (C.c_int64 * t.dim())(*t.shape)
                _lib.check(lib.ndit_set_weight(h, key.encode(), C.c_void_p(t.data_ptr()), shape, t.dim(), dt, stream), h)
            torch.cuda.current_stream(device).synchronize()   # temporaries above must outlive the copies
            _lib.check(lib.ndit_finalize_weights(h, stream), h)
        self._handle, self._dirty, self._cap_key = h, False, None
        return lib, h

    def parameter_count(self) -> int:
        return sum(p.numel() for p in self.parameters())

    # ------------------------------------------------------------------ shared argument handling
    def _ensure_capacity(self, lib, h, tokens: int, cap_len: int, batch: int) -> None:
        """Grow the engine workspace (weights stay packed) when a call needs more tokens / caption tokens / rows than
        the handle was created for: an unmodified sample.py at 2048x2048 must not fail on a default-constructed model."""
        t, c, b = self._limits
        batch += batch & 1                       # the workspace is sized for 2 or 4 rows
        if batch > 4:
            raise ValueError("the B200 engine runs at most 4 rows per call (a CFG pair of 2 samples, or 4 plain rows)")
        if tokens > t or cap_len > c or batch > b:
            self._limits = (max(t, tokens), max(c, cap_len), max(b, batch))
            _lib.check(lib.ndit_reserve(h, *self._limits), h)
            self._cap_key = None
            if hasattr(self, "_label_key"):
                self._label_key = None

    def _tokens_for(self, Hh: int, Ww: int) -> int:
        return (Hh // self.patch_size) * (Ww // self.patch_size)

    @staticmethod
    def _tensor_key(*tensors):
        """Cache key of conditioning tensors; None (= do not cache) for tensors without a version counter
        (created under torch.inference_mode())."""
        try:
            return tuple((t.data_ptr(), t._version, tuple(t.shape), t.dtype, str(t.device)) for t in tensors)
        except RuntimeError:
            return None

    def _check_inputs(self, x: torch.Tensor, *cond: torch.Tensor) -> None:
        """The engine computes in bf16 whatever the module / input dtype is (weights are packed to bf16, activations are bf16,
        fp32 accumulation): say so once instead of silently downgrading, and refuse tensors on another device."""
        if not x.is_cuda:
            raise RuntimeError(f"{type(self).__name__} (B200 engine) needs CUDA inputs; there is no CPU path")
        for c in cond:
            if isinstance(c, torch.Tensor) and c.device != x.device:
                raise RuntimeError(f"conditioning tensor on {c.device}, input on {x.device}: all tensors must be on the engine's device")
        pd = torch.bfloat16 if getattr(self, "_packed_path", None) is not None else next(self.parameters()).dtype
        if (pd != torch.bfloat16 or x.dtype != torch.bfloat16) and not getattr(self, "_warned_dtype", False):
            self._warned_dtype = True
            warnings.warn(f"{type(self).__name__}: parameters are {pd}, input is {x.dtype}; the B200 engine computes in bfloat16 "
                          "(bf16 weights and activations, fp32 accumulation) and returns the input dtype - the same numerics as the "
                          "reference under torch.autocast('cuda', torch.bfloat16) with bf16 parameters", stacklevel=3)

    @staticmethod
    def _uniform_t(t) -> float:
        """forward_with_cfg is called with t = ones(B) * t (transport.py:106): one timestep for all rows.  The engine embeds
        a single t; per-row timesteps would silently be wrong, so they are rejected (one D2H read, like model.py:888)."""
        if not isinstance(t, torch.Tensor):
            return float(t)
        vals = t.detach().reshape(-1).tolist()
        if any(v != vals[0] for v in vals):
            raise ValueError("the B200 engine needs the same timestep for every row of a forward_with_cfg call (got %r)" % (vals,))
        return float(vals[0])

    def launch_count(self) -> int:
        return int(_lib.load().ndit_launch_count(self._handle)) if self._handle is not None else 0

    def graph_replay_count(self) -> int:
        """Fixed-grid solves that ran as one CUDA-graph launch (ndit_graph_replay_count)."""
        return int(_lib.load().ndit_graph_replay_count(self._handle)) if self._handle is not None else 0

    def set_option(self, name: str, value: int) -> None:
        lib, h = self._engine(next(self.parameters()).device)
        _lib.check(lib.ndit_set_option(h, name.encode(), int(value)), h)

    def read_residual_tap(self, rows: int) -> torch.Tensor:
        """Debug: the residual stream [rows, dim] (bf16, token-major) recorded after the block chosen with
        ``set_option("tap_layer", l)`` during the last forward (ndit_debug_read_residual)."""
        dev = next(self.parameters()).device
        lib, h = self._engine(dev)
        out = torch.empty(rows, self.dim, dtype=torch.bfloat16, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.ndit_debug_read_residual(h, C.c_void_p(out.data_ptr()), rows, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), h)
        return out

    def get_fsdp_wrap_module_list(self):
        return list(self.layers)

    @torch.no_grad()
    def _run_forward(self, lib, h, x, t, sp):
        xb = x.detach().to(torch.bfloat16).contiguous()
        out = torch.empty_like(xb)
        tv = self._uniform_t(t)                                                 # the reference syncs here too (model.py:888)
        B, _, Hh, Ww = xb.shape
        self._ensure_capacity(lib, h, self._tokens_for(Hh, Ww), 0, B)
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(lib.ndit_forward_cfg(h, C.c_void_p(xb.data_ptr()), tv, B, Hh, Ww, C.byref(sp), C.c_void_p(out.data_ptr()), stream), h)
        return out.to(x.dtype)

    @torch.no_grad()
    def _run_sample(self, lib, h, z, t_grid, method, sp, return_trajectory):
        zb = z.detach().to(torch.bfloat16).contiguous()
        grid = [float(v) for v in t_grid]
        n = len(grid)
        garr = (C.c_float * n)(*grid)
        B, _, Hh, Ww = zb.shape
        self._ensure_capacity(lib, h, self._tokens_for(Hh, Ww), 0, B)
        traj = torch.empty((n,) + tuple(zb.shape), dtype=torch.bfloat16, device=z.device) if return_trajectory else None
        final = torch.empty_like(zb)
        m = {"euler": _lib.NDIT_EULER, "midpoint": _lib.NDIT_MIDPOINT, "rk4": _lib.NDIT_RK4}[method]
        stream = C.c_void_p(torch.cuda.current_stream(z.device).cuda_stream)
        _lib.check(lib.ndit_sample(h, C.c_void_p(zb.data_ptr()), B, Hh, Ww, garr, n, m, C.byref(sp),
                                   C.c_void_p(traj.data_ptr()) if traj is not None else None, C.c_void_p(final.data_ptr()), stream), h)
        return (traj if return_trajectory else final).to(z.dtype)


class NextDiT(EngineModule):
    """B200 engine behind the reference ``NextDiT`` API (model.py:665-989)."""

    def __init__(self, patch_size: int = 2, in_channels: int = 4, dim: int = 4096, n_layers: int = 32, n_heads: int = 32,
                 n_kv_heads: Optional[int] = None, multiple_of: int = 256, ffn_dim_multiplier: Optional[float] = None,
                 norm_eps: float = 1e-5, learn_sigma: bool = True, qk_norm: bool = False, cap_feat_dim: int = 5120,
                 scale_factor: float = 1.0, use_flash_attn: bool = True,
                 max_tokens: int = 4096, max_cap_len: int = 256, max_batch: int = 2) -> None:
        super().__init__()
        self.learn_sigma, self.in_channels, self.patch_size = learn_sigma, in_channels, patch_size
        self.out_channels = in_channels * 2 if learn_sigma else in_channels
        self.dim, self.n_heads, self.n_layers = dim, n_heads, n_layers
        self.n_kv_heads = n_kv_heads or n_heads
        self.cap_feat_dim, self.norm_eps, self.multiple_of = cap_feat_dim, norm_eps, multiple_of
        self.qk_norm, self.scale_factor = qk_norm, scale_factor
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
        assert (dim // n_heads) % 4 == 0, "2d rope needs head dim to be divisible by 4"
        self.pad_token = nn.Parameter(torch.empty(dim))
        nn.init.normal_(self.pad_token, std=0.02)
        self._init_engine_state(max_tokens, max_cap_len, max_batch)

    # ------------------------------------------------------------------ engine plumbing
    def _check_supported(self) -> None:
        if self.ffn_dim % 128 != 0:
            raise NotImplementedError(f"the B200 engine needs a FeedForward width that is a multiple of 128 (got {self.ffn_dim})")
        if self.in_channels % 2 != 0 or not 2 <= self.in_channels <= 16:
            raise NotImplementedError(f"the B200 engine runs an even in_channels in 2..16 (got {self.in_channels})")

    def _ndit_config(self):
        # ffn_dim: the hidden width as Python computed it in __init__ (model.py:470-473, incl. ffn_dim_multiplier);
        # no_qk_norm: qk_norm=False -> q_norm / k_norm / ky_norm are nn.Identity (model.py:219-220)
        return _lib.NditConfig(self.dim, self.n_layers, self.n_heads, self.n_kv_heads, self.cap_feat_dim, self.in_channels,
                               self.patch_size, self.multiple_of, int(self.learn_sigma), float(self.norm_eps), *self._limits, 0, 0, 0, 0,
                               int(self.ffn_dim), 0 if self.qk_norm else 1)

    def _set_caption(self, lib, h, cap_feats: torch.Tensor, cap_mask: torch.Tensor, stream):
        key = self._tensor_key(cap_feats, cap_mask)
        if key is not None and key == self._cap_key:
            return
        cap = cap_feats.detach().to(torch.bfloat16).contiguous()
        mask = (cap_mask.detach() != 0).to(torch.uint8).contiguous()
        B, T, Cd = cap.shape
        if Cd != self.cap_feat_dim or tuple(mask.shape) != (B, T):
            raise ValueError(f"cap_feats {tuple(cap.shape)} / cap_mask {tuple(mask.shape)} do not match cap_feat_dim {self.cap_feat_dim}")
        self._ensure_capacity(lib, h, 0, T, B)
        _lib.check(lib.ndit_set_caption(h, C.c_void_p(cap.data_ptr()), C.c_void_p(mask.data_ptr()), B, T, stream), h)
        self._cap_key = key
        self._cap_keepalive = (cap_feats, cap_mask, cap, mask)   # keep the ids in the cache key alive

    @staticmethod
    def _step_params(cfg_scale, scale_factor, scale_watershed, base_seqlen, proportional_attn):
        if proportional_attn:
            assert base_seqlen is not None
        return _lib.NditStepParams(float(cfg_scale), float(scale_factor), float(scale_watershed), int(bool(proportional_attn)),
                                   int(base_seqlen) if base_seqlen is not None else 0)

    # ------------------------------------------------------------------ reference API
    def _remember_call_state(self, t_last: float, scale_factor, scale_watershed, base_seqlen, proportional_attn) -> None:
        """What the reference module keeps after forward_with_cfg (model.py:883-899): self.freqs_cis (as its two scaling factors)
        and layer.attention.{base_seqlen, proportional_attn}; a later plain forward() uses them."""
        self._freqs_state = (float(scale_factor), 1.0) if t_last < scale_watershed else (1.0, float(scale_factor))
        self._attn_state = (bool(proportional_attn), int(base_seqlen) if proportional_attn else None)

    @torch.no_grad()
    def forward(self, x, t, cap_feats, cap_mask):
        """model.py:836-864 (inference): x [N,C,H,W] (or a list of N tensors [C,H_i,W_i] of different sizes, model.py:789-834),
        t [N], cap_feats [N,T,cap_feat_dim], cap_mask [N,T]; returns the first C output channels, [N,C,H,W] or a list.
        No guidance, one timestep per row.  Tensor input is processed in groups of at most max_batch rows; a list is one call
        (its padded length is part of the result when proportional attention is on), so len(x) <= max_batch is grown on demand.
        Uses the RoPE table / attention scaling the module currently holds, like the reference."""
        if not isinstance(x, torch.Tensor):
            return self._forward_list(list(x), t, cap_feats, cap_mask)
        self._check_inputs(x, cap_feats, cap_mask)
        lib, h = self._engine(x.device)
        lin, ntk = getattr(self, "_freqs_state", (1.0, float(self.scale_factor)))
        prop, base = getattr(self, "_attn_state", (False, None))
        sp = _lib.NditStepParams(0.0, lin, 1.0, int(prop), int(base) if base is not None else 0, ntk)
        tv = (t.detach().float().reshape(-1).tolist() if isinstance(t, torch.Tensor) else [float(t)] * x.shape[0])
        if len(tv) == 1:
            tv = tv * x.shape[0]
        if len(tv) != x.shape[0]:
            raise ValueError(f"t has {len(tv)} entries for a batch of {x.shape[0]}")
        xb = x.detach().to(torch.bfloat16).contiguous()
        out = torch.empty_like(xb)
        Bn, _, Hh, Ww = xb.shape
        with torch.cuda.device(x.device):
            stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            self._ensure_capacity(lib, h, self._tokens_for(Hh, Ww), cap_feats.shape[1], 1)
            step = self._limits[2]
            for i in range(0, Bn, step):
                j = min(Bn, i + step)
                self._cap_key = None
                self._set_caption(lib, h, cap_feats[i:j], cap_mask[i:j], stream)
                ta = (C.c_float * (j - i))(*tv[i:j])
                _lib.check(lib.ndit_forward(h, C.c_void_p(xb[i:j].data_ptr()), ta, j - i, Hh, Ww, C.byref(sp),
                                            C.c_void_p(out[i:j].data_ptr()), stream), h)
            self._cap_key = None
        return out.to(x.dtype)

    @torch.no_grad()
    def _forward_list(self, xs, t, cap_feats, cap_mask):
        n = len(xs)
        if n == 0 or any(not isinstance(v, torch.Tensor) or v.dim() != 3 for v in xs):
            raise ValueError("list input: a non-empty list of [C, H, W] tensors")
        self._check_inputs(xs[0], cap_feats, cap_mask, *xs[1:])
        dev = xs[0].device
        lib, h = self._engine(dev)
        lin, ntk = getattr(self, "_freqs_state", (1.0, float(self.scale_factor)))
        prop, base = getattr(self, "_attn_state", (False, None))
        sp = _lib.NditStepParams(0.0, lin, 1.0, int(prop), int(base) if base is not None else 0, ntk)
        tv = (t.detach().float().reshape(-1).tolist() if isinstance(t, torch.Tensor) else [float(t)] * n)
        if len(tv) == 1:
            tv = tv * n
        if len(tv) != n or cap_feats.shape[0] != n:
            raise ValueError(f"t / cap_feats have {len(tv)} / {cap_feats.shape[0]} rows for a list of {n}")
        xb = [v.detach().to(torch.bfloat16).contiguous() for v in xs]
        outs = [torch.empty_like(v) for v in xb]
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            self._ensure_capacity(lib, h, max(self._tokens_for(v.shape[1], v.shape[2]) for v in xb), cap_feats.shape[1], n)
            self._cap_key = None
            self._set_caption(lib, h, cap_feats, cap_mask, stream)
            xp = (C.c_void_p * n)(*[v.data_ptr() for v in xb])
            op = (C.c_void_p * n)(*[v.data_ptr() for v in outs])
            hs = (C.c_int32 * n)(*[v.shape[1] for v in xb])
            ws = (C.c_int32 * n)(*[v.shape[2] for v in xb])
            ta = (C.c_float * n)(*tv)
            _lib.check(lib.ndit_forward_list(h, xp, hs, ws, ta, n, C.byref(sp), op, stream), h)
            self._cap_key = None
        return [o.to(v.dtype) for o, v in zip(outs, xs)]

    @torch.no_grad()
    def forward_with_cfg(self, x, t, cap_feats, cap_mask, cfg_scale, scale_factor=1.0, scale_watershed=1.0,
                         base_seqlen: Optional[int] = None, proportional_attn: bool = False):
        """model.py:866-913.  x [2n,C,H,W]; first half = cond, second half ignored on input."""
        if not isinstance(x, torch.Tensor):
            raise NotImplementedError("list-of-tensors (variable resolution) input is not supported by the B200 engine")
        self._check_inputs(x, cap_feats, cap_mask)
        lib, h = self._engine(x.device)
        with torch.cuda.device(x.device):
            stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            self._ensure_capacity(lib, h, self._tokens_for(x.shape[2], x.shape[3]), cap_feats.shape[1], x.shape[0])
            self._set_caption(lib, h, cap_feats, cap_mask, stream)
            sp = self._step_params(cfg_scale, scale_factor, scale_watershed, base_seqlen, proportional_attn)
            out = self._run_forward(lib, h, x, t, sp)
            self._remember_call_state(self._uniform_t(t), scale_factor, scale_watershed, base_seqlen, proportional_attn)
            return out

    @torch.no_grad()
    def sample_fixed_grid(self, z, t_grid, method: str, cap_feats, cap_mask, cfg_scale, scale_factor=1.0, scale_watershed=1.0,
                          base_seqlen: Optional[int] = None, proportional_attn: bool = False, return_trajectory: bool = True):
        """Whole fixed-grid ODE solve inside the engine (used by transport.Sampler for euler / midpoint)."""
        self._check_inputs(z, cap_feats, cap_mask)
        lib, h = self._engine(z.device)
        with torch.cuda.device(z.device):
            stream = C.c_void_p(torch.cuda.current_stream(z.device).cuda_stream)
            self._ensure_capacity(lib, h, self._tokens_for(z.shape[2], z.shape[3]), cap_feats.shape[1], z.shape[0])
            self._set_caption(lib, h, cap_feats, cap_mask, stream)
            sp = self._step_params(cfg_scale, scale_factor, scale_watershed, base_seqlen, proportional_attn)
            out = self._run_sample(lib, h, z, t_grid, method, sp, return_trajectory)
            grid = [float(v) for v in t_grid]
            t_last = grid[-1] if method == "rk4" else (grid[-2] if method == "euler" else 0.5 * (grid[-2] + grid[-1]))
            self._remember_call_state(t_last, scale_factor, scale_watershed, base_seqlen, proportional_attn)
            return out


def _sample_sde_loop(self, z, pts, dt, sqrt_dt, half_dt, noise, method: int, cap_feats, cap_mask, cfg_scale, scale_factor=1.0,
                     scale_watershed=1.0, base_seqlen: Optional[int] = None, proportional_attn: bool = False):
    """The stochastic loop of transport.Sampler.sample_sde inside the engine (ndit_sample_sde); returns the list of states after
    each step.  ``pts``: (t, ratio, var, diffusion, sqrt(2 diffusion)) per evaluation point, ``noise`` [n_steps, *z.shape] bf16."""
    self._check_inputs(z, cap_feats, cap_mask, noise)
    lib, h = self._engine(z.device)
    n = noise.shape[0]
    assert len(pts) == n * (2 if method == 1 else 1) and tuple(noise.shape[1:]) == tuple(z.shape)
    with torch.cuda.device(z.device):
        stream = C.c_void_p(torch.cuda.current_stream(z.device).cuda_stream)
        self._ensure_capacity(lib, h, self._tokens_for(z.shape[2], z.shape[3]), cap_feats.shape[1], z.shape[0])
        self._set_caption(lib, h, cap_feats, cap_mask, stream)
        sp = self._step_params(cfg_scale, scale_factor, scale_watershed, base_seqlen, proportional_attn)
        zb = z.detach().to(torch.bfloat16).contiguous()
        nb = noise.detach().to(torch.bfloat16).contiguous()
        traj = torch.empty_like(nb)
        arr = (_lib.NditSdePoint * len(pts))(*[_lib.NditSdePoint(*p) for p in pts])
        B, _, Hh, Ww = zb.shape
        _lib.check(lib.ndit_sample_sde(h, C.c_void_p(zb.data_ptr()), B, Hh, Ww, n, int(method), arr, float(dt), float(sqrt_dt), float(half_dt),
                                       C.c_void_p(nb.data_ptr()), C.byref(sp), C.c_void_p(traj.data_ptr()), stream), h)
        self._remember_call_state(pts[-1][0], scale_factor, scale_watershed, base_seqlen, proportional_attn)
    return [traj[i].to(z.dtype) for i in range(n)]


NextDiT.sample_sde_loop = _sample_sde_loop


def NextDiT_2B_patch2(**kwargs):
    """model.py:994-995."""
    return NextDiT(patch_size=2, dim=2304, n_layers=24, n_heads=32, **kwargs)


def NextDiT_2B_GQA_patch2(**kwargs):
    """model.py:998-999."""
    return NextDiT(patch_size=2, dim=2304, n_layers=24, n_heads=32, n_kv_heads=8, **kwargs)
