"""VAE-decode end (include/ndit_vae.h) on the GPU against oracle/vae_oracle.py (parity unpinned: diffusers is absent
from the reference tree and the image, see the oracle's header)."""
import os
import sys
import time

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

from oracle import vae_oracle as VO  # noqa: E402


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()


def _rms_rel(a, b):
    return ((a.float() - b.float()).pow(2).mean().sqrt() / b.float().pow(2).mean().sqrt()).item()


def _model(cfg, W):
    from lumina_t2x_b200.vae import AutoencoderKL
    m = AutoencoderKL(latent_channels=cfg.latent_channels, out_channels=cfg.out_channels, block_out_channels=cfg.block_out_channels,
                      layers_per_block=cfg.layers_per_block, norm_num_groups=cfg.norm_num_groups)
    m.load_state_dict(W, strict=True)
    return m.cuda()


@pytest.mark.parametrize("shape", [(1, 8, 8), (2, 12, 20)])
def test_decode_tiny_vs_oracle(shape):
    """Small decoder (128, 128, 256, 256; one resnet + 1 per block): every kernel class of the walk (conv_in, GroupNorm, implicit
    3x3 GEMM with N = 128 / 256, 1x1 shortcut, attention, upsample, conv_out) against the fp32 oracle and its autocast emulation."""
    cfg = VO.config_tiny()
    W = VO.synthetic_weights(cfg, seed=3)
    B, lh, lw = shape
    z = torch.randn(B, 4, lh, lw, generator=torch.Generator().manual_seed(5)).to(torch.bfloat16)
    ref32 = VO.decode(cfg, W, z.float(), "fp32")
    ref16 = VO.decode(cfg, W, z.float(), "bf16")
    m = _model(cfg, W)
    out = m.decode(z.cuda()).sample
    assert out.shape == ref32.shape and out.dtype == torch.bfloat16
    out = out.float().cpu()
    assert torch.isfinite(out).all()
    floor = _rel(ref16, ref32)
    print(f"\n  vae tiny {shape}: engine vs fp32 {_rel(out, ref32):.3e}  oracle-bf16 vs fp32 {floor:.3e}  engine vs oracle-bf16 {_rel(out, ref16):.3e}"
          f"  rms {_rms_rel(out, ref32):.3e}")
    assert _rel(out, ref32) < 1.5 * floor + 5e-3
    assert _rms_rel(out, ref32) < 1.5 * _rms_rel(ref16, ref32) + 2e-3
    # a second call at the same shape replays the cached plans; a different shape rebuilds the workspace
    out2 = m.decode(z.cuda()).sample.float().cpu()
    assert torch.equal(out, out2)


def test_decode_state_dict_surface():
    """Full AutoencoderKL state dicts load (encoder / quant_conv entries dropped, old attention names mapped, strict otherwise)."""
    cfg = VO.config_tiny()
    W = VO.synthetic_weights(cfg, seed=1)
    full = dict(W)
    full["encoder.conv_in.weight"] = torch.zeros(128, 3, 3, 3)
    full["quant_conv.weight"] = torch.zeros(8, 8, 1, 1)
    a = "decoder.mid_block.attentions.0"
    for new, old in (("to_q", "query"), ("to_k", "key"), ("to_v", "value"), ("to_out.0", "proj_attn")):
        full[f"{a}.{old}.weight"] = full.pop(f"{a}.{new}.weight")[:, :, None, None]
        full[f"{a}.{old}.bias"] = full.pop(f"{a}.{new}.bias")
    m = _model(cfg, full)
    assert set(m.state_dict()) == set(W)
    z = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16)
    out = m.decode(z.cuda()).sample.float().cpu()
    out_b = _model(cfg, W).decode(z.cuda()).sample.float().cpu()
    assert torch.equal(out, out_b)
    bad = dict(W)
    bad.pop("decoder.conv_out.bias")
    from lumina_t2x_b200.vae import AutoencoderKL
    with pytest.raises(RuntimeError):
        AutoencoderKL(block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block).load_state_dict(bad, strict=True)


def test_decode_sdxl_shape_properties_and_time():
    """sdxl-vae architecture (128, 256, 512, 512; 2 + 1 resnets per block) at the headline size: one 128 x 128 latent -> 1024 x 1024.
    The CPU oracle at this size takes minutes, so: finite, deterministic, batch rows independent of each other (decode of a batch
    of two equals the two single decodes), and the convolution path is translation-covariant away from the borders."""
    cfg = VO.VaeCfg()
    W = VO.synthetic_weights(cfg, seed=0)
    m = _model(cfg, W)
    g = torch.Generator().manual_seed(9)
    z = torch.randn(2, 4, 32, 32, generator=g).to(torch.bfloat16).cuda()
    both = m.decode(z).sample
    one = m.decode(z[1:]).sample
    assert torch.isfinite(both.float()).all()
    assert torch.equal(both[1:], one)
    # mid-size parity against the oracle (a 32 x 32 latent is ~0.6 TFLOP: seconds on the CPU)
    ref16 = VO.decode(cfg, W, z[:1].float().cpu(), "bf16")
    ref32 = VO.decode(cfg, W, z[:1].float().cpu(), "fp32")
    out = both[:1].float().cpu()
    floor = _rel(ref16, ref32)
    print(f"\n  vae sdxl 32x32: engine vs fp32 {_rel(out, ref32):.3e}  oracle-bf16 vs fp32 {floor:.3e}  rms {_rms_rel(out, ref32):.3e} / {_rms_rel(ref16, ref32):.3e}")
    assert _rel(out, ref32) < 1.5 * floor + 5e-3
    z1 = torch.randn(1, 4, 128, 128, generator=g).to(torch.bfloat16).cuda()
    o1 = m.decode(z1).sample
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        o2 = m.decode(z1).sample
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    assert o1.shape == (1, 3, 1024, 1024) and torch.isfinite(o1.float()).all() and torch.equal(o1, o2)
    print(f"  vae sdxl 128x128 -> 1024x1024: {dt * 1e3:.2f} ms per decode ({10.3 / dt / 1e3:.2f} PFLOP/s of ~10.3 TFLOP)")
