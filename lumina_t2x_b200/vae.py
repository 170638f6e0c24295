"""Reference-side mirror of the VAE-decode end: ``diffusers.AutoencoderKL`` as lumina_next_t2i/sample.py uses it
(sample.py:117-120, :238),

    vae = AutoencoderKL.from_pretrained("stabilityai/sdxl-vae", torch_dtype=torch.float32).cuda()
    samples = vae.decode(samples / factor).sample

behind the C ABI of include/ndit_vae.h (nvae_*).  Same constructor keywords as AutoencoderKL's config (``latent_channels``,
``out_channels``, ``block_out_channels``, ``layers_per_block``, ``norm_num_groups``, ``scaling_factor``), same state-dict keys for
the decode half (``post_quant_conv.*``, ``decoder.*``); ``load_state_dict`` also takes a full AutoencoderKL state dict and drops
the ``encoder.*`` / ``quant_conv.*`` entries, which the sampling path never touches.  Only ``decode`` is implemented.
Arithmetic is the reference's autocast(bf16) run of the fp32 module: bf16 convolutions with fp32 accumulation, fp32 GroupNorm /
SiLU, bf16 output.  No PyTorch fallback: without the CUDA library / an sm_100 device the call raises."""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Dict, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib


def _decoder_shapes(latent_channels: int, out_channels: int, block_out_channels: Sequence[int], layers_per_block: int) -> Dict[str, Tuple[int, ...]]:
    ch = tuple(reversed(tuple(block_out_channels)))
    S: Dict[str, Tuple[int, ...]] = {}

    def conv(pre, cin, cout, k):
        S[pre + ".weight"], S[pre + ".bias"] = (cout, cin, k, k), (cout,)

    def norm(pre, c):
        S[pre + ".weight"], S[pre + ".bias"] = (c,), (c,)

    def res(pre, cin, cout):
        norm(pre + ".norm1", cin)
        conv(pre + ".conv1", cin, cout, 3)
        norm(pre + ".norm2", cout)
        conv(pre + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(pre + ".conv_shortcut", cin, cout, 1)

    conv("post_quant_conv", latent_channels, latent_channels, 1)
    conv("decoder.conv_in", latent_channels, ch[0], 3)
    res("decoder.mid_block.resnets.0", ch[0], ch[0])
    res("decoder.mid_block.resnets.1", ch[0], ch[0])
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", ch[0])
    for name in ("to_q", "to_k", "to_v", "to_out.0"):
        S[f"{a}.{name}.weight"], S[f"{a}.{name}.bias"] = (ch[0], ch[0]), (ch[0],)
    for i in range(len(ch)):
        cin = ch[0] if i == 0 else ch[i - 1]
        for j in range(layers_per_block + 1):
            res(f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else ch[i], ch[i])
        if i < len(ch) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", ch[i], ch[i], 3)
    norm("decoder.conv_norm_out", ch[-1])
    conv("decoder.conv_out", ch[-1], out_channels, 3)
    return S


_OLD_ATTN = {".query.": ".to_q.", ".key.": ".to_k.", ".value.": ".to_v.", ".proj_attn.": ".to_out.0."}


class AutoencoderKL(nn.Module):
    """Decode half of diffusers' AutoencoderKL on the B200 engine.  Parameters live in a flat ParameterDict-like table keyed by the
    diffusers state-dict names (dots replaced for nn.Module registration); ``state_dict()`` / ``load_state_dict()`` use the real names."""

    def __init__(self, in_channels: int = 3, out_channels: int = 3, latent_channels: int = 4,
                 block_out_channels: Sequence[int] = (128, 256, 512, 512), layers_per_block: int = 2, norm_num_groups: int = 32,
                 scaling_factor: float = 0.13025, **unused):
        super().__init__()
        if len(block_out_channels) != 4:
            raise NotImplementedError("the engine implements the four-level decoder of sdxl-vae / sd-vae-ft")
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels, latent_channels=latent_channels,
                                      block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                                      norm_num_groups=norm_num_groups, scaling_factor=scaling_factor)
        self._shapes = _decoder_shapes(latent_channels, out_channels, block_out_channels, layers_per_block)
        g = torch.Generator().manual_seed(0)
        for k, shp in self._shapes.items():
            if "norm" in k:
                t = torch.ones(shp) if k.endswith("weight") else torch.zeros(shp)
            elif k.endswith("bias"):
                t = torch.zeros(shp)
            else:
                fan_in = shp[1] * (shp[2] * shp[3] if len(shp) == 4 else 1)
                t = torch.randn(shp, generator=g) * fan_in ** -0.5
            self.register_parameter(k.replace(".", "__"), nn.Parameter(t, requires_grad=False))
        self._handle, self._dirty = None, True

    # ------------------------------------------------------------------ state dict under the diffusers names
    def state_dict(self, *a, **k):
        return {n.replace("__", "."): p for n, p in super().state_dict(*a, **k).items()}

    def load_state_dict(self, state_dict, strict: bool = True, **k):
        sd = {}
        for key, v in state_dict.items():
            if key.startswith("encoder.") or key.startswith("quant_conv."):
                continue
            if "attentions.0" in key:
                for old, new in _OLD_ATTN.items():
                    key = key.replace(old, new)
            want = self._shapes.get(key)
            if want is not None and len(want) == 2 and v.dim() == 4:     # Linear stored as a 1x1 convolution by old checkpoints
                v = v.reshape(want)
            sd[key.replace(".", "__")] = v
        self._dirty = True
        return super().load_state_dict(sd, strict=strict, **k)

    def _apply(self, fn, *a, **k):
        self._dirty = True
        return super()._apply(fn, *a, **k)

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().nvae_destroy(self._handle)
        except Exception:
            pass

    @staticmethod
    def _check(lib, h, rc):
        if rc != 0:
            raise RuntimeError(f"nvae error {rc}: {(lib.nvae_last_error(h) or b'?').decode()}")

    def _engine(self, device):
        lib = _lib.load()
        if self._handle is not None and not self._dirty:
            return lib, self._handle
        if self._handle is not None:
            lib.nvae_destroy(self._handle)
            self._handle = None
        c = self.config
        cfg = _lib.NvaeConfig(c.latent_channels, c.out_channels, (C.c_int32 * 4)(*c.block_out_channels), c.layers_per_block, c.norm_num_groups)
        h = C.c_void_p()
        with torch.cuda.device(device):
            rc = lib.nvae_create(C.byref(cfg), C.byref(h))
            if rc != 0:
                raise RuntimeError(f"nvae_create failed ({rc}): {(lib.nvae_last_error(None) or b'?').decode()}")
            stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
            for key, p in self.state_dict().items():
                t = p.detach()
                if t.device != device:
                    raise RuntimeError(f"parameter {key} is on {t.device}, expected {device}")
                if t.dtype == torch.bfloat16:
                    dt = _lib.NDIT_BF16
                else:
                    t, dt = t.float(), _lib.NDIT_F32
                t = t.contiguous()
                shape = (C.c_int64 * t.dim())(*t.shape)
                self._check(lib, h, lib.nvae_set_weight(h, key.encode(), C.c_void_p(t.data_ptr()), shape, t.dim(), dt, stream))
                del t
            torch.cuda.current_stream(device).synchronize()
            self._check(lib, h, lib.nvae_finalize_weights(h, stream))
        self._handle, self._dirty = h, False
        return lib, h

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True, **unused):
        """vae.decode(z) -> DecoderOutput-like object with ``.sample`` [B, out_channels, 8h, 8w] (bf16, the dtype autocast returns)."""
        if not z.is_cuda:
            raise RuntimeError("AutoencoderKL (B200 engine) needs a CUDA latent; there is no CPU path")
        if z.dim() != 4 or z.shape[1] != self.config.latent_channels:
            raise ValueError(f"latent must be [B, {self.config.latent_channels}, h, w], got {tuple(z.shape)}")
        dev = z.device
        lib, h = self._engine(dev)
        zb = z.detach().to(torch.bfloat16).contiguous()
        B, _, lh, lw = zb.shape
        out = torch.empty(B, self.config.out_channels, 8 * lh, 8 * lw, dtype=torch.bfloat16, device=dev)
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            self._check(lib, h, lib.nvae_decode(h, C.c_void_p(zb.data_ptr()), B, lh, lw, C.c_void_p(out.data_ptr()), stream))
        if not return_dict:
            return (out,)
        return SimpleNamespace(sample=out)

    def forward(self, *a, **k):
        raise NotImplementedError("only AutoencoderKL.decode is on the sampling path (sample.py:238); encode / forward are out of scope")
