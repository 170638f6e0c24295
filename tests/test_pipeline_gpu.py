"""The whole sampling pipeline of lumina_next_t2i/sample.py (:23-53, :177-240) on the three engines, in the reference's call sequence
(examples/sample_pipeline.py::generate): caption encoder -> hidden_states[-2] -> cap_feats of the ODE solve around forward_with_cfg ->
[-1][:1] -> vae.decode(samples / factor).sample -> (x + 1) / 2 clamped.  Every stage is checked against ITS oracle on the tensors the
previous engine stage produced (the data formats on either side of the hot path are the hand-over points: bf16 [2, T, C] caption
features + int64 [2, T] mask in, bf16 [1, 4, h/8, w/8] latent out), with the tolerances of the per-stage test files.
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))

pytestmark = pytest.mark.gpu


def _rel(a, b, sel=None):
    a, b = a.float(), b.float()
    if sel is not None:
        a, b = a[sel], b[sel]
    return ((a - b).abs().max() / b.abs().max()).item()


def test_sample_py_pipeline_stage_by_stage():
    import sample_pipeline as SP
    from oracle import gemma_oracle as G
    from oracle import nextdit_oracle as O
    from oracle import vae_oracle as VO
    (enc, gcfg, GW), (dit, dcfg, DW), (vae, vcfg, VW) = SP.build_tiny()
    ids, mask = SP.tiny_prompt(gcfg)
    z = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(1))
    st = {}
    kw = dict(cfg_scale=2.0, num_sampling_steps=5, sampling_method="midpoint", time_shifting_factor=4.0, train_image_size=128)
    img = SP.generate(enc, dit, vae, ids, mask, z, stages=st, **kw)
    assert img.shape == (1, 3, 128, 128) and img.dtype == torch.bfloat16
    assert float(img.min()) >= 0.0 and float(img.max()) <= 1.0 and torch.isfinite(img.float()).all()

    # stage 1: caption features = hidden_states[-2] (valid positions: the denoiser masks the padded ones out)
    cap = st["cap_feats"].float().cpu()
    assert cap.shape == (2, ids.shape[1], gcfg.hidden_size) and st["cap_feats"].dtype == torch.bfloat16
    orc_cap = G.hidden_states_m2(gcfg, GW, ids, mask, "bf16")
    assert _rel(cap, orc_cap, mask.bool()) < 2e-2, _rel(cap, orc_cap, mask.bool())

    # stage 2: the ODE solve on the ENGINE's caption features (what sample.py hands over), last grid point, first row
    mk = st["model_kwargs"]
    assert mk["proportional_attn"] is True and mk["base_seqlen"] == (128 // 16) ** 2 and mk["scale_factor"] == 1.0
    lat = st["latent"].float().cpu()
    assert lat.shape == (1, 4, 16, 16)
    zb = z.to(torch.bfloat16).repeat(2, 1, 1, 1)
    orc_traj = O.sample_ode(dcfg, DW, zb, st["cap_feats"].cpu(), mask, num_steps=5, method="midpoint", time_shifting_factor=4.0, cfg_scale=2.0,
                            scale_factor=1.0, scale_watershed=1.0, base_seqlen=64, proportional_attn=True, precision="bf16")
    assert _rel(lat, orc_traj[-1][:1]) < 4e-2, _rel(lat, orc_traj[-1][:1])     # 8 model calls; per-call bound 2e-2 (test_model_gpu.py)

    # stage 3: VAE decode of the ENGINE's latent divided by the scaling factor, then the reference's post-processing
    dec = st["decoded"].float().cpu()
    zin = (st["latent"] / 0.13025).float().cpu()
    ref32, ref16 = VO.decode(vcfg, VW, zin, "fp32"), VO.decode(vcfg, VW, zin, "bf16")
    assert dec.shape == ref32.shape
    assert _rel(dec, ref32) < 1.5 * _rel(ref16, ref32) + 5e-3, (_rel(dec, ref32), _rel(ref16, ref32))
    post = ((st["decoded"] + 1.0) / 2.0).clamp_(0.0, 1.0)
    assert torch.equal(img, post)

    # the same call again: cached caption state, graph-replayed solve, cached VAE plans - same bits
    img2 = SP.generate(enc, dit, vae, ids, mask, z, **kw)
    assert torch.equal(img, img2)
