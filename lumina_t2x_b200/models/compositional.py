"""Drop-in mirror of ``lumina_next_compositional_generation/models`` (``NextDiT_2B_GQA_patch2`` / ``NextDiT_2B_patch2`` of
``models/model.py:1044-1049``): the Lumina-Next-T2I NextDiT whose caption cross-attention is region-masked (``Attention.forward``
:421-446) and whose adaLN conditioning pools a separate global caption (``forward`` :852-899).

Same constructor, state-dict keys and ``forward_with_cfg`` signature as the reference (:902-953), so the reference's ``demo.py``
(:185-250: ``sample_fn(z, model.forward_with_cfg, cap_feats=..., cap_mask=..., global_cap_feats=..., global_cap_mask=..., h_split_num=...,
w_split_num=..., ...)``) works unchanged.  All compute is in libndit_b200.so (``ndit_set_caption_regions`` + the usual
``ndit_forward_cfg`` / ``ndit_sample``); there is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from .. import _lib
from .nextdit import NextDiT as _BaseNextDiT


class NextDiT(_BaseNextDiT):
    """B200 engine behind the compositional ``NextDiT`` API (model.py:680-1040)."""

    def _check_supported(self) -> None:
        super()._check_supported()
        if self.dim // self.n_heads != 72:
            raise NotImplementedError("region-masked cross-attention is built for head_dim 72 (the 2B models of the reference)")

    def _set_caption_regions(self, lib, h, cap_feats, cap_mask, global_cap_feats, global_cap_mask, h_split_num, w_split_num, stream):
        hs, ws = int(h_split_num), int(w_split_num)
        key = self._tensor_key(cap_feats, cap_mask, global_cap_feats, global_cap_mask)
        key = None if key is None else key + (("regions", hs, ws),)
        if key is not None and key == self._cap_key:
            return
        cap = cap_feats.detach().to(torch.bfloat16).contiguous()
        mask = (cap_mask.detach() != 0).to(torch.uint8).contiguous()
        gcap = global_cap_feats.detach().to(torch.bfloat16).contiguous()
        gmask = (global_cap_mask.detach() != 0).to(torch.uint8).contiguous()
        R, T, Cd = cap.shape
        if Cd != self.cap_feat_dim or tuple(mask.shape) != (R, T):
            raise ValueError(f"cap_feats {tuple(cap.shape)} / cap_mask {tuple(mask.shape)} do not match cap_feat_dim {self.cap_feat_dim}")
        if gcap.dim() != 3 or gcap.shape[0] != 1 or gcap.shape[2] != Cd or tuple(gmask.shape) != tuple(gcap.shape[:2]):
            raise ValueError(f"global_cap_feats {tuple(gcap.shape)} / global_cap_mask {tuple(gmask.shape)}: one global caption row [1, T, {Cd}]")
        if R < 2:
            raise ValueError("cap_feats holds the region captions followed by the unconditional caption: at least 2 rows")
        if hs < 1 or ws < 1 or hs * ws - 1 >= R:
            # model.py:879-883 indexes region_mask[(h + 1) * (w + 1) - 1] on a tensor with cap_feats.shape[0] rows
            raise IndexError(f"h_split_num x w_split_num = {hs} x {ws} needs region id {hs * ws - 1} < {R} caption rows")
        self._ensure_capacity(lib, h, 0, max(T, gcap.shape[1]), 2)
        _lib.check(lib.ndit_set_caption_regions(h, C.c_void_p(cap.data_ptr()), C.c_void_p(mask.data_ptr()), R, T, C.c_void_p(gcap.data_ptr()),
                                                C.c_void_p(gmask.data_ptr()), gcap.shape[1], hs, ws, stream), h)
        self._cap_key = key
        self._cap_keepalive = (cap_feats, cap_mask, global_cap_feats, global_cap_mask, cap, mask, gcap, gmask)

    @staticmethod
    def _need_globals(global_cap_feats, global_cap_mask):
        if global_cap_feats is None or global_cap_mask is None:
            # the reference dereferences global_cap_mask unconditionally (model.py:866)
            raise AttributeError("global_cap_feats / global_cap_mask are required by the compositional NextDiT (model.py:866)")

    @torch.no_grad()
    def forward(self, x, t, cap_feats, cap_mask, global_cap_feats=None, global_cap_mask=None, h_split_num=1, w_split_num=1):
        """model.py:852-899 (inference): x [2, C, H, W] - row 0 attends to the region captions, row 1 to the last caption (:421-446 pairs the
        FIRST row with every caption but the last and the LAST row with the last one, so the reference itself only makes sense for a pair);
        t [2]; returns the first C output channels [2, C, H, W], no guidance.  Uses the RoPE table / attention scaling the module holds."""
        self._need_globals(global_cap_feats, global_cap_mask)
        if not isinstance(x, torch.Tensor) or x.dim() != 4 or x.shape[0] != 2:
            raise ValueError("the compositional forward takes a pair of rows: x [2, C, H, W]")
        self._check_inputs(x, cap_feats, cap_mask, global_cap_feats, global_cap_mask)
        lib, h = self._engine(x.device)
        lin, ntk = getattr(self, "_freqs_state", (1.0, float(self.scale_factor)))
        prop, base = getattr(self, "_attn_state", (False, None))
        sp = _lib.NditStepParams(0.0, lin, 1.0, int(prop), int(base) if base is not None else 0, ntk)
        tv = (t.detach().float().reshape(-1).tolist() if isinstance(t, torch.Tensor) else [float(t)] * 2)
        if len(tv) == 1:
            tv = tv * 2
        if len(tv) != 2:
            raise ValueError(f"t has {len(tv)} entries for a batch of 2")
        xb = x.detach().to(torch.bfloat16).contiguous()
        out = torch.empty_like(xb)
        with torch.cuda.device(x.device):
            stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            self._ensure_capacity(lib, h, self._tokens_for(x.shape[2], x.shape[3]), max(cap_feats.shape[1], global_cap_feats.shape[1]), 2)
            self._set_caption_regions(lib, h, cap_feats, cap_mask, global_cap_feats, global_cap_mask, h_split_num, w_split_num, stream)
            ta = (C.c_float * 2)(*tv)
            _lib.check(lib.ndit_forward(h, C.c_void_p(xb.data_ptr()), ta, 2, x.shape[2], x.shape[3], C.byref(sp), C.c_void_p(out.data_ptr()), stream), h)
        return out.to(x.dtype)

    @torch.no_grad()
    def forward_with_cfg(self, x, t, cap_feats, cap_mask, cfg_scale, scale_factor=1.0, scale_watershed=1.0, base_seqlen: Optional[int] = None,
                         proportional_attn: bool = False, global_cap_feats=None, global_cap_mask=None, h_split_num=1, w_split_num=1):
        """model.py:902-953.  x [2, C, H, W] (cond, uncond; the second row is ignored on input); cap_feats [R + 1, T, C] = region captions
        followed by the unconditional one."""
        self._need_globals(global_cap_feats, global_cap_mask)
        if not isinstance(x, torch.Tensor) or x.dim() != 4 or x.shape[0] != 2:
            raise ValueError("the compositional forward_with_cfg takes one latent as a cond / uncond pair: x [2, C, H, W]")
        self._check_inputs(x, cap_feats, cap_mask, global_cap_feats, global_cap_mask)
        lib, h = self._engine(x.device)
        with torch.cuda.device(x.device):
            stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            self._ensure_capacity(lib, h, self._tokens_for(x.shape[2], x.shape[3]), max(cap_feats.shape[1], global_cap_feats.shape[1]), 2)
            self._set_caption_regions(lib, h, cap_feats, cap_mask, global_cap_feats, global_cap_mask, h_split_num, w_split_num, stream)
            sp = self._step_params(cfg_scale, scale_factor, scale_watershed, base_seqlen, proportional_attn)
            out = self._run_forward(lib, h, x, t, sp)
            self._remember_call_state(self._uniform_t(t), scale_factor, scale_watershed, base_seqlen, proportional_attn)
            return out

    @torch.no_grad()
    def sample_fixed_grid(self, z, t_grid, method: str, cap_feats, cap_mask, cfg_scale, scale_factor=1.0, scale_watershed=1.0,
                          base_seqlen: Optional[int] = None, proportional_attn: bool = False, global_cap_feats=None, global_cap_mask=None,
                          h_split_num=1, w_split_num=1, return_trajectory: bool = True):
        """Whole fixed-grid ODE solve inside the engine (transport.Sampler routes euler / midpoint / rk4 here)."""
        self._need_globals(global_cap_feats, global_cap_mask)
        if z.dim() != 4 or z.shape[0] != 2:
            raise ValueError("the compositional model samples one latent as a cond / uncond pair: z [2, C, H, W]")
        self._check_inputs(z, cap_feats, cap_mask, global_cap_feats, global_cap_mask)
        lib, h = self._engine(z.device)
        with torch.cuda.device(z.device):
            stream = C.c_void_p(torch.cuda.current_stream(z.device).cuda_stream)
            self._ensure_capacity(lib, h, self._tokens_for(z.shape[2], z.shape[3]), max(cap_feats.shape[1], global_cap_feats.shape[1]), 2)
            self._set_caption_regions(lib, h, cap_feats, cap_mask, global_cap_feats, global_cap_mask, h_split_num, w_split_num, stream)
            sp = self._step_params(cfg_scale, scale_factor, scale_watershed, base_seqlen, proportional_attn)
            out = self._run_sample(lib, h, z, t_grid, method, sp, return_trajectory)
            grid = [float(v) for v in t_grid]
            t_last = grid[-1] if method == "rk4" else (grid[-2] if method == "euler" else 0.5 * (grid[-2] + grid[-1]))
            self._remember_call_state(t_last, scale_factor, scale_watershed, base_seqlen, proportional_attn)
            return out

    @torch.no_grad()
    def sample_sde_loop(self, z, pts, dt, sqrt_dt, half_dt, noise, method: int, cap_feats, cap_mask, cfg_scale, scale_factor=1.0,
                        scale_watershed=1.0, base_seqlen: Optional[int] = None, proportional_attn: bool = False, global_cap_feats=None,
                        global_cap_mask=None, h_split_num=1, w_split_num=1):
        """The stochastic loop of transport.Sampler.sample_sde inside the engine (ndit_sample_sde), region-masked captions."""
        self._need_globals(global_cap_feats, global_cap_mask)
        if z.dim() != 4 or z.shape[0] != 2:
            raise ValueError("the compositional model samples one latent as a cond / uncond pair: z [2, C, H, W]")
        self._check_inputs(z, cap_feats, cap_mask, global_cap_feats, global_cap_mask, noise)
        lib, h = self._engine(z.device)
        n = noise.shape[0]
        assert len(pts) == n * (2 if method == 1 else 1) and tuple(noise.shape[1:]) == tuple(z.shape)
        with torch.cuda.device(z.device):
            stream = C.c_void_p(torch.cuda.current_stream(z.device).cuda_stream)
            self._ensure_capacity(lib, h, self._tokens_for(z.shape[2], z.shape[3]), max(cap_feats.shape[1], global_cap_feats.shape[1]), 2)
            self._set_caption_regions(lib, h, cap_feats, cap_mask, global_cap_feats, global_cap_mask, h_split_num, w_split_num, stream)
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


def NextDiT_2B_patch2(**kwargs):
    """model.py:1044-1045."""
    return NextDiT(patch_size=2, dim=2304, n_layers=24, n_heads=32, **kwargs)


def NextDiT_2B_GQA_patch2(**kwargs):
    """model.py:1048-1049."""
    return NextDiT(patch_size=2, dim=2304, n_layers=24, n_heads=32, n_kv_heads=8, **kwargs)
