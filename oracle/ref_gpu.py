"""Run the UNMODIFIED reference NextDiT on the GPU next to the engine.

TEST INFRASTRUCTURE (only tests/ and bench.py's baseline legs import this).  The reference sources come from
``oracle/harness/ref_import.py`` (``/root/reference`` here, the byte-identical ``oracle/_ref`` copies on the GPU box).

Two precisions of the same unmodified module (lumina_next_t2i_mini/models/nextdit.py; the canonical fairscale flavour
lumina_next_t2i/models/model.py is available through ``flavour="full"``):
  * ``fp32``: parameters and inputs in float32, TF32 off -> the SDPA branch of Attention.forward (nextdit.py:358-373);
  * ``bf16``: parameters in bfloat16 under ``torch.autocast("cuda", torch.bfloat16)`` -> ``flash_attn_varlen_func``
    (nextdit.py:327-355) - exactly what sample.py runs (sample.py:125-129,177-188) - the "stock CUDA path".
"""
from __future__ import annotations

import contextlib
import math

import torch

from oracle.harness import ref_import


def randomize_(module: torch.nn.Module, seed: int = 0) -> None:
    """Deterministic non-degenerate weights for a NextDiT-shaped module (ours or the reference: same parameter names).
    The reference zero-initialises adaLN / final layer / cap_embedder / gates, which makes the output identically 0."""
    dev = next(module.parameters()).device
    g = torch.Generator(device=dev).manual_seed(seed)
    with torch.no_grad():
        for k, p in module.named_parameters():
            if p.dim() == 2:
                std = (0.5 if "adaLN" in k else 1.0) / math.sqrt(p.shape[1])
                p.copy_(torch.randn(p.shape, generator=g, device=dev) * std)
            elif k.endswith("gate"):
                p.copy_(0.5 * torch.randn(p.shape, generator=g, device=dev))
            elif ("norm" in k and k.endswith("weight")) or k == "cap_embedder.0.weight":
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g, device=dev))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g, device=dev))


def build_reference(state_dict, *, dim, n_layers, n_heads, n_kv_heads, cap_feat_dim, dtype, device="cuda", flavour="mini",
                    use_flash_attn=True):
    """Construct the unmodified reference NextDiT, strict-load `state_dict` (pins key names and shapes), cast to `dtype`."""
    if flavour == "mini":
        mod = ref_import.import_reference_mini()[0].nextdit
        extra = dict(use_flash_attn=use_flash_attn)
    elif flavour == "compositional":     # lumina_next_compositional_generation/models/model.py (region-masked cross-attention)
        mod = ref_import.import_reference_compositional_model()
        extra = {}
    else:
        mod = ref_import.import_reference_full_model()
        extra = {}
    with torch.device(device):
        m = mod.NextDiT(patch_size=2, in_channels=4, dim=dim, n_layers=n_layers, n_heads=n_heads, n_kv_heads=n_kv_heads, qk_norm=True,
                        cap_feat_dim=cap_feat_dim, **extra)
    m.load_state_dict({k: v.to(device) for k, v in state_dict.items()}, strict=True)
    return m.eval().to(device=device, dtype=dtype)


@contextlib.contextmanager
def precision_ctx(dtype):
    """fp32: TF32 off everywhere; bf16: the autocast context of sample.py:177."""
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            if dtype == torch.bfloat16:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    yield
            else:
                yield
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def ref_forward(m, z, t, cap, mask, **kw):
    dtype = next(m.parameters()).dtype
    with precision_ctx(dtype):
        return m.forward_with_cfg(z.to(dtype), t, cap.to(dtype), mask, **kw)


def ref_sample(m, z, cap, mask, num_steps, method, time_shifting_factor, **kw):
    """The reference's own ODE class (lumina_next_t2i_mini/transport.py:57-111) around its own forward_with_cfg."""
    transport = ref_import.import_reference_mini()[1]
    dtype = next(m.parameters()).dtype
    with precision_ctx(dtype):
        return transport.ODE(num_steps, method, time_shifting_factor).sample(z.to(dtype), m.forward_with_cfg, cap_feats=cap.to(dtype),
                                                                            cap_mask=mask, **kw)


def rel_linf(a, b) -> float:
    return ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()
