"""GPU parity of the region-masked (compositional) NextDiT: ``lumina_t2x_b200.models.compositional`` -> ndit_set_caption_regions ->
attention_fused_kernel<72, REGION> against the fixtures recorded from the unmodified
``lumina_next_compositional_generation/models/model.py`` (tests/golden/comp_*.pt, fp32 on CPU), the oracle's bf16-rounding mode,
the engine's own CUDA-core attention kernel, and - where oracle/_ref is present - the real reference on the same GPU.

Tolerances (relative L-inf): engine vs the reference's fp32 output within 1.5x the distance of a bf16 pipeline from it (the oracle's
bf16 mode on CPU / the reference's own autocast + flash-attn path on the GPU) + 2e-3; engine vs oracle(bf16 rounding points) 3e-2: two
bf16 pipelines that are each ~1.6e-2 from the fp32 truth (fixtures) sit up to the sum apart - measured 2.1e-2 on comp_2x2 (guidance scale
3 amplifies cond - uncond differences), against > 5e-2 .. 1 for a wrong region assignment (test_regions_matter_and_state_switches).
"""
import dataclasses
import os

import pytest
import torch

from oracle import compositional_oracle as CO
from oracle import nextdit_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max()).item()


def _build(cfg, W, **kw):
    from lumina_t2x_b200.models import compositional
    m = compositional.NextDiT(dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, n_kv_heads=cfg.n_kv_heads, qk_norm=True,
                              cap_feat_dim=cfg.cap_feat_dim, **kw)
    m.load_state_dict(W, strict=True)
    return m.eval().to("cuda", dtype=torch.bfloat16)


def _call(m, z, t, cap, mask, gcap, gmask, hs, ws, **kw):
    return m.forward_with_cfg(z.cuda(), t.cuda(), cap.cuda(), mask.cuda(), global_cap_feats=gcap.cuda(), global_cap_mask=gmask.cuda(),
                              h_split_num=hs, w_split_num=ws, **kw).float().cpu()


@pytest.fixture(scope="module")
def tiny_mha():
    cfg = dataclasses.replace(O.config_tiny(n_layers=2), n_kv_heads=O.config_tiny().n_heads)
    W = O.synthetic_weights(cfg, seed=0)
    return cfg, W, _build(cfg, W, max_tokens=512, max_cap_len=32)


@pytest.mark.parametrize("name", ["comp_2x2", "comp_1x3", "comp_1x1"])
@pytest.mark.parametrize("attn", ["tcgen05", "refkernel"])
def test_forward_with_cfg_vs_reference_fixture(tiny_mha, name, attn):
    cfg, W, m = tiny_mha
    fx = torch.load(os.path.join(GOLD, f"{name}.pt"), map_location="cpu", weights_only=False)
    z, cap, mask, gcap, gmask = CO.synthetic_inputs(cfg, tuple(fx["hw"]), fx["n_regions"], fx["T"], seed=fx["input_seed"])
    t = torch.full((2,), fx["t"])
    m.set_option("attn_ref", 1 if attn == "refkernel" else 0)
    out = _call(m, z, t, cap, mask, gcap, gmask, fx["hs"], fx["ws"], **fx["kw"])
    m.set_option("attn_ref", 0)
    assert out.shape == z.shape and torch.isfinite(out).all()
    orc = CO.forward_with_cfg(cfg, W, z, t, cap, mask, precision="bf16", global_cap_feats=gcap, global_cap_mask=gmask,
                              h_split_num=fx["hs"], w_split_num=fx["ws"], **fx["kw"])
    ref32 = fx["out_fp32"]
    floor = _rel(orc, ref32)
    print("COMPOSITIONAL", name, attn, dict(engine_vs_oracle_bf16=_rel(out, orc), engine_vs_ref_fp32=_rel(out, ref32), oracle_bf16_vs_ref_fp32=floor))
    assert _rel(out, orc) < 3e-2, (_rel(out, orc), floor)
    assert _rel(out, ref32) < 1.5 * floor + 2e-3, (_rel(out, ref32), floor)
    assert torch.equal(out[0, :3], out[1, :3])          # CFG structure (model.py:944-951)


def test_regions_matter_and_state_switches(tiny_mha):
    """The region assignment is live (another split changes the output well beyond bf16 noise), a repeated call is bit-identical, and
    the handle goes back and forth between region-masked and plain captions (ndit_set_caption clears the region state)."""
    cfg, W, m = tiny_mha
    z, cap, mask, gcap, gmask = CO.synthetic_inputs(cfg, (32, 32), 4, 16, seed=5)
    t = torch.full((2,), 0.4)
    kw = dict(cfg_scale=2.0, scale_factor=1.0, scale_watershed=1.0, base_seqlen=64, proportional_attn=True)
    a = _call(m, z, t, cap, mask, gcap, gmask, 2, 2, **kw)
    b = _call(m, z, t, cap, mask, gcap, gmask, 1, 2, **kw)
    a2 = _call(m, z, t, cap, mask, gcap, gmask, 2, 2, **kw)
    assert torch.equal(a, a2)
    assert _rel(b, a) > 5e-2
    orc_b = CO.forward_with_cfg(cfg, W, z, t, cap, mask, precision="bf16", global_cap_feats=gcap, global_cap_mask=gmask, h_split_num=1,
                                w_split_num=2, **kw)
    assert _rel(b, orc_b) < 3e-2, _rel(b, orc_b)
    # plain captions on the same handle: the base-class path (per-row pooled caption, no regions), then regions again
    from lumina_t2x_b200.models.nextdit import NextDiT as Base
    zp, capp, maskp = O.synthetic_inputs(cfg, (32, 32), 16, 8, seed=6)
    plain = Base.forward_with_cfg(m, zp.cuda(), t.cuda(), capp.cuda(), maskp.cuda(), **kw).float().cpu()
    orc_p = O.forward_with_cfg(cfg, W, zp, t, capp, maskp, precision="bf16", **kw)
    assert _rel(plain, orc_p) < 2e-2
    a3 = _call(m, z, t, cap, mask, gcap, gmask, 2, 2, **kw)
    assert torch.equal(a, a3)
    with pytest.raises(IndexError):
        _call(m, z, t, cap, mask, gcap, gmask, 2, 3, **kw)      # region id 5 on 5 caption rows: the reference raises IndexError


@pytest.mark.parametrize("attn", ["tcgen05", "refkernel"])
def test_gqa_long_captions_ragged_tokens(attn):
    """GQA (n_kv_heads = 2), captions longer than one 128-row kv block (T = 136 -> two blocks per caption, 7 captions -> the caption
    buffers grow past max_batch rows), a token count that is not a multiple of the 256-row CTA tile (28 x 40 latent = 280 tokens),
    3 x 2 regions with a row remainder (14 // 3 = 4: token rows 12, 13 belong to no region), the NTK branch of the time-aware RoPE.  Engine vs the oracle's bf16 mode."""
    cfg = O.config_tiny(n_layers=2)
    W = O.synthetic_weights(cfg, seed=2)
    m = _build(cfg, W, max_tokens=256, max_cap_len=16)          # T = 136 > 16: the mirror grows the workspace
    z, cap, mask, gcap, gmask = CO.synthetic_inputs(cfg, (28, 40), 6, 136, seed=9)
    mask[2, :] = 0
    mask[2, 130:134] = 1                                        # a caption whose only valid keys sit in its SECOND kv block
    t = torch.full((2,), 0.65)
    kw = dict(cfg_scale=3.0, scale_factor=2.0, scale_watershed=0.3, base_seqlen=64, proportional_attn=True)
    m.set_option("attn_ref", 1 if attn == "refkernel" else 0)
    out = _call(m, z, t, cap, mask, gcap, gmask, 3, 2, **kw)
    m.set_option("attn_ref", 0)
    orc = CO.forward_with_cfg(cfg, W, z, t, cap, mask, precision="bf16", global_cap_feats=gcap, global_cap_mask=gmask, h_split_num=3,
                              w_split_num=2, **kw)
    assert torch.isfinite(out).all()
    print("COMPOSITIONAL gqa_long", attn, _rel(out, orc))
    assert _rel(out, orc) < 3e-2, _rel(out, orc)


def test_plain_forward_pair(tiny_mha):
    """NextDiT.forward of the compositional model (model.py:852-899): a pair of DIFFERENT rows with different timesteps, row 0 on the region
    captions, row 1 on the last caption, no guidance; on a fresh module (default RoPE table, no proportional attention)."""
    cfg, W, _ = tiny_mha
    m = _build(cfg, W, max_tokens=256, max_cap_len=16)
    z, cap, mask, gcap, gmask = CO.synthetic_inputs(cfg, (24, 32), 4, 16, seed=17)
    x = torch.randn(2, cfg.in_channels, 24, 32, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16)
    t = torch.tensor([0.3, 0.8])
    out = m(x.cuda(), t.cuda(), cap.cuda(), mask.cuda(), global_cap_feats=gcap.cuda(), global_cap_mask=gmask.cuda(), h_split_num=2,
            w_split_num=2).float().cpu()
    orc = CO.forward(cfg, W, x, t, cap, mask, gcap, gmask, 2, 2, precision="bf16")
    ref = CO.forward(cfg, W, x.float(), t, cap.float(), mask, gcap.float(), gmask, 2, 2, precision="fp32")
    assert out.shape == x.shape and torch.isfinite(out).all()
    assert _rel(out, orc) < 3e-2, _rel(out, orc)
    assert _rel(out, ref) < 1.5 * _rel(orc, ref) + 2e-3, (_rel(out, ref), _rel(orc, ref))


@pytest.mark.parametrize("method", ["Euler", "Heun"])
def test_sde_loop_in_engine_equals_host_loop(tiny_mha, method):
    """transport.Sampler.sample_sde with the compositional kwargs: the in-engine stochastic loop (ndit_sample_sde over region-masked
    captions) gives the bits of the mirror's host loop around the same forward_with_cfg."""
    from lumina_t2x_b200 import transport
    cfg, W, m = tiny_mha
    z, cap, mask, gcap, gmask = CO.synthetic_inputs(cfg, (32, 32), 4, 16, seed=19)
    kw = dict(cap_feats=cap.cuda(), cap_mask=mask.cuda(), cfg_scale=2.0, scale_factor=1.0, scale_watershed=1.0, base_seqlen=64,
              proportional_attn=True, global_cap_feats=gcap.cuda(), global_cap_mask=gmask.cuda(), h_split_num=2, w_split_num=2)
    tr = transport.create_transport("Linear", "velocity", None, None, None)
    fn = transport.Sampler(tr).sample_sde(sampling_method=method, diffusion_form="sigma", diffusion_norm=1.0, last_step="Mean",
                                          last_step_size=0.04, num_steps=5)
    zb = z.cuda().to(torch.bfloat16)
    torch.manual_seed(7)
    n0 = m.launch_count()
    fused = fn(zb, m.forward_with_cfg, **kw)
    n1 = m.launch_count()
    torch.manual_seed(7)
    host = fn(zb, lambda x, t, **k: m.forward_with_cfg(x, t, **k), **kw)
    assert len(fused) == len(host) == 5 and n1 > n0
    for i, (a, b) in enumerate(zip(fused, host)):
        assert torch.equal(a, b), (method, i, (a.float() - b.float()).abs().max().item())


@pytest.mark.parametrize("method", ["euler", "midpoint"])
def test_sampler_fused_solve_equals_generic_loop(tiny_mha, method):
    """transport.Sampler.sample_ode with the compositional kwargs (demo.py:185-250): the in-engine solve (graph replay on the second
    call) gives the bits of the generic per-step loop through forward_with_cfg."""
    from lumina_t2x_b200 import transport
    cfg, W, m = tiny_mha
    z, cap, mask, gcap, gmask = CO.synthetic_inputs(cfg, (32, 32), 4, 16, seed=11)
    kw = dict(cap_feats=cap.cuda(), cap_mask=mask.cuda(), cfg_scale=4.0, scale_factor=1.0, scale_watershed=1.0, base_seqlen=64,
              proportional_attn=True, global_cap_feats=gcap.cuda(), global_cap_mask=gmask.cuda(), h_split_num=2, w_split_num=2)
    tr = transport.create_transport("Linear", "velocity", None, None, None)
    fn = transport.Sampler(tr).sample_ode(sampling_method=method, num_steps=5, atol=1e-6, rtol=1e-3, reverse=False, time_shifting_factor=4.0)
    zc = z.cuda()
    r0 = m.graph_replay_count()
    fused = fn(zc, m.forward_with_cfg, **kw)
    fused2 = fn(zc, m.forward_with_cfg, **kw)                   # second solve: captured and launched as one CUDA graph
    assert m.graph_replay_count() == r0 + 1
    generic = transport._fixed_grid_torch(lambda tt, xx: m.forward_with_cfg(xx, torch.ones(2, device="cuda") * tt, **kw), zc,
                                          transport._time_grid(0, 1, 5, 4.0).cuda(), method)
    assert fused.shape == (5,) + tuple(z.shape) and torch.isfinite(fused).all()
    assert torch.equal(fused, fused2)
    assert torch.equal(fused, generic)


def test_against_the_real_reference_on_the_gpu():
    """The unmodified compositional reference on the same B200 (oracle/_ref): fp32 (TF32 off) = the truth, autocast(bf16) +
    flash_attn_varlen_func = its stock path, at the 2B widths (MHA: the reference's fp32 SDPA branch cannot run GQA), 4 layers,
    1024 x 1024 (2 x 4096 tokens), 2 x 2 regions, T = 128.  Bar: the engine is as close to the fp32 truth as the reference's own bf16 path."""
    from oracle import ref_gpu
    from oracle.harness import ref_import
    if not os.path.isfile(os.path.join(ref_import.REF_ROOT, "lumina_next_compositional_generation", "models", "model.py")):
        pytest.skip("oracle/_ref holds no compositional reference")
    from lumina_t2x_b200.models import compositional
    with torch.device("cuda"):
        m = compositional.NextDiT(patch_size=2, dim=2304, n_layers=4, n_heads=32, n_kv_heads=None, qk_norm=True, cap_feat_dim=2048,
                                  max_tokens=4096, max_cap_len=128)
    ref_gpu.randomize_(m, 4)
    m = m.eval().to("cuda", dtype=torch.bfloat16)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(8)
    z = torch.randn(1, 4, 128, 128, generator=g).to(torch.bfloat16).repeat(2, 1, 1, 1).cuda()
    cap = torch.randn(5, 128, 2048, generator=g).to(torch.bfloat16).cuda()
    mask = torch.zeros(5, 128, dtype=torch.int64)
    for r, n in enumerate((128, 96, 40, 77, 8)):
        mask[r, :n] = 1
    mask = mask.cuda()
    gcap = torch.randn(1, 128, 2048, generator=g).to(torch.bfloat16).cuda()
    gmask = torch.ones(1, 128, dtype=torch.int64).cuda()
    t = torch.full((2,), 0.45, device="cuda")
    kw = dict(cfg_scale=4.0, scale_factor=1.0, scale_watershed=1.0, base_seqlen=4096, proportional_attn=True, global_cap_feats=gcap,
              global_cap_mask=gmask, h_split_num=2, w_split_num=2)
    eng = m.forward_with_cfg(z, t, cap, mask, **kw).float()
    dims = dict(dim=2304, n_layers=4, n_heads=32, n_kv_heads=None, cap_feat_dim=2048)

    def run(dtype):
        ref = ref_gpu.build_reference(sd, dtype=dtype, flavour="compositional", **dims)
        with ref_gpu.precision_ctx(dtype):
            k2 = dict(kw, global_cap_feats=gcap.to(dtype))
            return ref.forward_with_cfg(z.to(dtype), t, cap.to(dtype), mask, **k2).float()

    out16 = run(torch.bfloat16)
    # the same bf16 run with PyTorch's math SDPA backend for the masked caption cross-attention (flash_attn_varlen_func of the self-attention
    # is unaffected): separates bf16 rounding noise from what the fused SDPA backends do with fully masked (token, caption) rows
    from torch.nn.attention import SDPBackend, sdpa_kernel
    with sdpa_kernel(SDPBackend.MATH):
        out16_math = run(torch.bfloat16)
    out32 = run(torch.float32)
    floor, floor_math, mine, cross = _rel(out16, out32), _rel(out16_math, out32), _rel(eng, out32), _rel(eng, out16_math)
    rec = dict(test="compositional_mha_4layers_2x2", ref_bf16_vs_fp32=floor, ref_bf16_math_sdpa_vs_fp32=floor_math, engine_vs_fp32=mine,
               engine_vs_ref_bf16_math_sdpa=cross, torch=torch.__version__)
    print("COMPOSITIONAL_REFERENCE_PARITY", rec)
    try:
        import json
        out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "reference_parity.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
    assert torch.isfinite(eng).all()
    # Measured on torch 2.11: the reference's default autocast path is 0.67 (relative L-inf) away from its own fp32 output on this
    # input - far outside bf16 noise (1.6e-2 for the plain model at the same size, profiles/r02_reference_parity_full_size.jsonl) -
    # while the engine is 2.0e-2 from the fp32 truth.  So the bar is the tighter of the two bf16 runs, and an absolute cap.
    best = min(floor, floor_math)
    assert mine <= max(1.5 * best + 5e-3, 2.5e-2), (mine, floor, floor_math)
