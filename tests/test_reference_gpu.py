"""Full-size GPU parity against the REAL reference (unmodified sources: /root/reference here, oracle/_ref on the GPU box).

Three implementations of the same forward_with_cfg on the same weights and inputs, on the same B200:
  ref32   the reference in fp32 (TF32 off, SDPA branch)                              = the truth
  ref16   the reference under autocast(bf16) with flash_attn_varlen_func             = what sample.py runs (stock CUDA path)
  engine  this repo (models.NextDiT -> C ABI -> sm_100a kernels), bf16 with fp32 accumulation

Bar ("next round" item 1): rel(engine, ref32) <= 1.5 * rel(ref16, ref32) [+ 2e-3], i.e. the engine must be
as close to the fp32 truth as the reference's own bf16 path is (relative L-inf = max|a-b| / max|b|).  The north-star's
1e-3 is below what any bf16 pipeline delivers: the measured floors are printed by every test and recorded in DESIGN.md.
"""
import gc
import os

import pytest
import torch

from oracle.harness import ref_import

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_import.reference_available(), reason="oracle/_ref not built")]

RESULTS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "reference_parity.jsonl")


def _record(**kw):
    import json
    try:
        os.makedirs(os.path.dirname(RESULTS), exist_ok=True)
        with open(RESULTS, "a") as f:
            f.write(json.dumps(kw) + "\n")
    except OSError:
        pass
    print("REFERENCE_PARITY", json.dumps(kw))


def _engine(n_layers, max_tokens, seed=0, n_kv_heads=8):
    from lumina_t2x_b200 import models
    from oracle import ref_gpu
    with torch.device("cuda"):
        m = models.NextDiT(patch_size=2, dim=2304, n_layers=n_layers, n_heads=32, n_kv_heads=n_kv_heads, qk_norm=True, cap_feat_dim=2048,
                           max_tokens=max_tokens, max_cap_len=128)
    ref_gpu.randomize_(m, seed)
    m = m.eval().to("cuda", dtype=torch.bfloat16)
    return m, {k: v.detach().clone() for k, v in m.state_dict().items()}


def _inputs(hw, T=128, ul=8, seed=1):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(1, 4, *hw, generator=g).to(torch.bfloat16).repeat(2, 1, 1, 1).cuda()
    cap = torch.randn(2, T, 2048, generator=g).to(torch.bfloat16).cuda()
    mask = torch.zeros(2, T, dtype=torch.int64)
    mask[0, :] = 1
    mask[1, :ul] = 1
    return z, cap, mask.cuda()


def _free():
    gc.collect()
    torch.cuda.empty_cache()


def test_config2_full_size_forward_and_trajectory():
    """BASELINE config 2 itself: NextDiT_2B_GQA_patch2, 24 layers, 2 x 4096 tokens, T = 128, proportional attention.
    One forward_with_cfg and the full 30-point Euler solve (29 model calls) against the real reference."""
    from lumina_t2x_b200 import transport
    from oracle import ref_gpu
    m, sd = _engine(24, 4096)
    z, cap, mask = _inputs((128, 128))
    kw = dict(cfg_scale=2.0, scale_factor=1.0, scale_watershed=1.0, base_seqlen=4096, proportional_attn=True)
    t = torch.full((2,), 0.3, device="cuda")
    dims = dict(dim=2304, n_layers=24, n_heads=32, n_kv_heads=8, cap_feat_dim=2048)

    eng = m.forward_with_cfg(z, t, cap, mask, **kw).float()
    tr = transport.create_transport("Linear", "velocity", None, None, None)
    fn = transport.Sampler(tr).sample_ode(sampling_method="euler", num_steps=30, atol=1e-6, rtol=1e-3, reverse=False, time_shifting_factor=1.0)
    eng_traj = fn(z, m.forward_with_cfg, cap_feats=cap, cap_mask=mask, **kw)[-1].float()

    ref16 = ref_gpu.build_reference(sd, dtype=torch.bfloat16, **dims)
    out16 = ref_gpu.ref_forward(ref16, z, t, cap, mask, **kw).float()
    traj16 = ref_gpu.ref_sample(ref16, z, cap, mask, 30, "euler", 1.0, **kw)[-1].float()
    # the canonical (fairscale) flavour on the same weights: same bits expected under bf16 + flash-attn
    full16 = ref_gpu.build_reference(sd, dtype=torch.bfloat16, flavour="full", **dims)
    out16_full = ref_gpu.ref_forward(full16, z, t, cap, mask, **kw).float()
    del ref16, full16
    _free()

    ref32 = ref_gpu.build_reference(sd, dtype=torch.float32, **dims)
    out32 = ref_gpu.ref_forward(ref32, z, t, cap, mask, **kw)
    traj32 = ref_gpu.ref_sample(ref32, z, cap, mask, 30, "euler", 1.0, **kw)[-1]
    del ref32
    _free()

    floor, mine, cross = ref_gpu.rel_linf(out16, out32), ref_gpu.rel_linf(eng, out32), ref_gpu.rel_linf(eng, out16)
    tfloor, tmine, tcross = ref_gpu.rel_linf(traj16, traj32), ref_gpu.rel_linf(eng_traj, traj32), ref_gpu.rel_linf(eng_traj, traj16)
    _record(test="config2_full", forward=dict(ref_bf16_vs_fp32=floor, engine_vs_fp32=mine, engine_vs_ref_bf16=cross),
            trajectory_30pt_euler=dict(ref_bf16_vs_fp32=tfloor, engine_vs_fp32=tmine, engine_vs_ref_bf16=tcross),
            canonical_vs_mini_bf16=ref_gpu.rel_linf(out16_full, out16))
    assert torch.isfinite(eng).all() and torch.isfinite(eng_traj).all()
    assert ref_gpu.rel_linf(out16_full, out16) < 1e-6          # the two reference flavours agree
    assert mine <= 1.5 * floor + 2e-3, (mine, floor)
    assert tmine <= 1.5 * tfloor + 2e-3, (tmine, tfloor)


def test_config2_mha_variant_one_forward():
    """NextDiT_2B_patch2 (MHA, model.py:994-995) at the config-2 shape, 4 layers: the canonical fairscale flavour can run
    this one in fp32 too (its SDPA branch has no GQA repeat)."""
    from oracle import ref_gpu
    m, sd = _engine(4, 4096, seed=3, n_kv_heads=None)
    z, cap, mask = _inputs((128, 128), seed=5)
    kw = dict(cfg_scale=4.0, scale_factor=1.0, scale_watershed=1.0, base_seqlen=None, proportional_attn=False)
    t = torch.full((2,), 0.7, device="cuda")
    dims = dict(dim=2304, n_layers=4, n_heads=32, n_kv_heads=None, cap_feat_dim=2048)
    eng = m.forward_with_cfg(z, t, cap, mask, **kw).float()
    out16 = ref_gpu.ref_forward(ref_gpu.build_reference(sd, dtype=torch.bfloat16, flavour="full", **dims), z, t, cap, mask, **kw).float()
    out32 = ref_gpu.ref_forward(ref_gpu.build_reference(sd, dtype=torch.float32, flavour="full", **dims), z, t, cap, mask, **kw)
    floor, mine = ref_gpu.rel_linf(out16, out32), ref_gpu.rel_linf(eng, out32)
    _record(test="config2_mha_4layers_canonical", ref_bf16_vs_fp32=floor, engine_vs_fp32=mine, engine_vs_ref_bf16=ref_gpu.rel_linf(eng, out16))
    _free()
    assert mine <= 1.5 * floor + 2e-3, (mine, floor)


@pytest.mark.parametrize("tval", [0.1, 0.8])
def test_config3_full_width_2048(tval):
    """BASELINE config 3 shape: 2048 x 2048 image = 2 x 16384 tokens, time-aware scaled RoPE (scale_factor 2, watershed 0.3:
    t = 0.1 takes the linear-interpolation branch, t = 0.8 the NTK branch), proportional attention, 2 layers at the 2B widths."""
    from oracle import ref_gpu
    m, sd = _engine(2, 16384, seed=1)
    z, cap, mask = _inputs((256, 256), seed=3)
    kw = dict(cfg_scale=4.0, scale_factor=2.0, scale_watershed=0.3, base_seqlen=4096, proportional_attn=True)
    t = torch.full((2,), tval, device="cuda")
    dims = dict(dim=2304, n_layers=2, n_heads=32, n_kv_heads=8, cap_feat_dim=2048)
    eng = m.forward_with_cfg(z, t, cap, mask, **kw).float()
    ref16 = ref_gpu.build_reference(sd, dtype=torch.bfloat16, **dims)
    out16 = ref_gpu.ref_forward(ref16, z, t, cap, mask, **kw).float()
    del ref16
    _free()
    cross = ref_gpu.rel_linf(eng, out16)
    floor = mine = None
    free_b, _ = torch.cuda.mem_get_info()
    if free_b > 150 * 2**30:       # the fp32 SDPA branch materialises a [2,32,16384,16384] fp32 mask (69 GB) plus workspace
        ref32 = ref_gpu.build_reference(sd, dtype=torch.float32, **dims)
        out32 = ref_gpu.ref_forward(ref32, z, t, cap, mask, **kw)
        del ref32
        floor, mine = ref_gpu.rel_linf(out16, out32), ref_gpu.rel_linf(eng, out32)
        del out32
    _free()
    _record(test="config3_2layers", t=tval, ref_bf16_vs_fp32=floor, engine_vs_fp32=mine, engine_vs_ref_bf16=cross)
    assert torch.isfinite(eng).all()
    assert cross < 2e-2, cross                 # two bf16 implementations of a 2-layer model
    if floor is not None:
        assert mine <= 1.5 * floor + 2e-3, (mine, floor)
