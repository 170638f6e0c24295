"""Drop-in boundary on the GPU: the call sequence of the reference's sample.py executed against the mirror, workspace growth,
CUDA-graph replay of repeated solves, dtype contract.  (Reference lines: lumina_next_t2i/sample.py:125-142,177-234.)"""
import math
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu


def _randomize(m, seed=0):
    from oracle import ref_gpu
    ref_gpu.randomize_(m, seed)
    return m


@pytest.fixture(scope="module")
def flagship():
    """sample.py:125-142 verbatim: factory looked up by name in the models module, DEFAULT ctor limits, .eval().to(cuda, dtype),
    strict load_state_dict of a checkpoint with the reference key names."""
    from lumina_t2x_b200 import models
    dtype = torch.bfloat16
    model = models.__dict__["NextDiT_2B_GQA_patch2"](qk_norm=True, cap_feat_dim=2048)
    model.eval().to("cuda", dtype=dtype)
    with torch.device("cuda"):
        donor = models.__dict__["NextDiT_2B_GQA_patch2"](qk_norm=True, cap_feat_dim=2048)
    ckpt = {k: v.to("cpu", torch.bfloat16) for k, v in _randomize(donor, 7).state_dict().items()}
    del donor
    model.load_state_dict(ckpt, strict=True)
    return model


@pytest.mark.parametrize("res", ["1024:1024x1024", "2048:2048x2048"])
def test_sample_py_call_sequence(flagship, res):
    from lumina_t2x_b200.transport import Sampler, create_transport
    model, dtype, image_size = flagship, torch.bfloat16, 1024
    with torch.autocast("cuda", dtype):
        transport = create_transport("Linear", "velocity", None, None, None)
        sampler = Sampler(transport)
        sample_fn = sampler.sample_ode(sampling_method="euler", num_steps=4, atol=1e-6, rtol=1e-3, reverse=False, time_shifting_factor=1.0)
        torch.random.manual_seed(1)
        res_cat, resolution = res.split(":")
        do_extrapolation = int(res_cat) > 1024
        w, h = (int(v) for v in resolution.split("x"))
        latent_w, latent_h = w // 8, h // 8
        z = torch.randn([1, 4, latent_w, latent_h], device="cuda").to(dtype)
        z = z.repeat(2, 1, 1, 1)
        cap_feats = torch.randn(2, 40, 2048, device="cuda").to(dtype)
        cap_mask = torch.zeros(2, 40, dtype=torch.int64, device="cuda")
        cap_mask[0, :33] = 1
        cap_mask[1, :8] = 1
        model_kwargs = dict(cap_feats=cap_feats, cap_mask=cap_mask, cfg_scale=4.0)
        model_kwargs["proportional_attn"] = True
        model_kwargs["base_seqlen"] = (image_size // 16) ** 2
        if do_extrapolation:
            model_kwargs["scale_factor"] = math.sqrt(w * h / image_size ** 2)
            model_kwargs["scale_watershed"] = 0.3
        else:
            model_kwargs["scale_factor"] = 1.0
            model_kwargs["scale_watershed"] = 1.0
        samples = sample_fn(z, model.forward_with_cfg, **model_kwargs)[-1]
        samples = samples[:1]
    assert samples.shape == (1, 4, latent_w, latent_h) and samples.dtype == dtype
    assert torch.isfinite(samples.float()).all()
    # a second identical solve replays the captured graph: same bits
    with torch.autocast("cuda", dtype):
        again = sample_fn(z, model.forward_with_cfg, **model_kwargs)[-1][:1]
        third = sample_fn(z, model.forward_with_cfg, **model_kwargs)[-1][:1]
    assert torch.equal(samples, again) and torch.equal(samples, third)


def _tiny(**kw):
    from lumina_t2x_b200 import models
    from oracle import nextdit_oracle as O
    cfg = O.config_tiny(n_layers=2)
    W = O.synthetic_weights(cfg, seed=0)
    m = models.NextDiT(dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, n_kv_heads=cfg.n_kv_heads, qk_norm=True,
                       cap_feat_dim=cfg.cap_feat_dim, **kw)
    m.load_state_dict(W, strict=True)
    return cfg, W, m.eval().to("cuda", dtype=torch.bfloat16)


def test_workspace_grows_instead_of_failing():
    """More tokens / caption tokens than the handle was created for: ndit_reserve re-creates the workspace (weights stay
    packed) and the result equals the one of an engine sized for the call from the start."""
    from oracle import nextdit_oracle as O
    cfg, W, small = _tiny(max_tokens=64, max_cap_len=8)
    _, _, big = _tiny(max_tokens=1024, max_cap_len=64)
    kw = dict(cfg_scale=2.0, scale_factor=1.0, scale_watershed=1.0, base_seqlen=64, proportional_attn=True)
    for hw, T in (((16, 16), 8), ((32, 48), 24), ((16, 16), 8)):
        z, cap, mask = O.synthetic_inputs(cfg, hw, T, 4, seed=2)
        t = torch.full((2,), 0.4, device="cuda")
        a = small.forward_with_cfg(z.cuda(), t, cap.cuda(), mask.cuda(), **kw)
        b = big.forward_with_cfg(z.cuda(), t, cap.cuda(), mask.cuda(), **kw)
        assert torch.equal(a, b), hw


def test_graph_replay_equals_direct_launches_and_survives_other_shapes():
    from lumina_t2x_b200 import transport
    from oracle import nextdit_oracle as O
    cfg, W, m = _tiny(max_tokens=1024, max_cap_len=64)
    kw = dict(cfg_scale=2.0, scale_factor=2.0, scale_watershed=0.3, base_seqlen=64, proportional_attn=True)   # both RoPE branches
    fn = transport.Sampler(transport.create_transport("Linear", "velocity", None, None, None)).sample_ode(
        sampling_method="midpoint", num_steps=6, atol=1e-6, rtol=1e-3, reverse=False, time_shifting_factor=4.0)
    zA, capA, maskA = (v.cuda() for v in O.synthetic_inputs(cfg, (32, 32), 24, 8, seed=4))
    zB, capB, maskB = (v.cuda() for v in O.synthetic_inputs(cfg, (16, 48), 16, 8, seed=5))
    m.set_option("graph", 0)
    refA = fn(zA, m.forward_with_cfg, cap_feats=capA, cap_mask=maskA, **kw)
    refB = fn(zB, m.forward_with_cfg, cap_feats=capB, cap_mask=maskB, **kw)
    m.set_option("graph", 1)
    launches = []
    r0 = m.graph_replay_count()
    for i in range(3):            # 1st: direct (registers the solve), 2nd: capture + launch, 3rd: replay
        n0 = m.launch_count()
        a = fn(zA, m.forward_with_cfg, cap_feats=capA, cap_mask=maskA, **kw)
        launches.append(m.launch_count() - n0)
        assert torch.equal(a, refA), i
        b = fn(zB, m.forward_with_cfg, cap_feats=capB, cap_mask=maskB, **kw)     # another shape in between
        assert torch.equal(b, refB), i
    assert launches[0] > 0 and launches[2] >= launches[0] - 8        # replays are counted with the launches they contain
    # the graphs really ran: both shapes were captured on the 2nd pass and replayed on the 3rd (PyTorch's current stream is the legacy
    # default stream, which cannot be captured - the engine records on its own stream and launches the graph on the caller's)
    assert m.graph_replay_count() - r0 == 4, m.graph_replay_count() - r0
    # a forward_with_cfg of yet another shape right after a replay (V^T layout / RoPE slots are shared state)
    zC, capC, maskC = (v.cuda() for v in O.synthetic_inputs(cfg, (24, 24), 16, 8, seed=6))
    t = torch.full((2,), 0.7, device="cuda")
    c1 = m.forward_with_cfg(zC, t, capC, maskC, **kw)
    m.set_option("graph", 0)
    c2 = m.forward_with_cfg(zC, t, capC, maskC, **kw)
    assert torch.equal(c1, c2)


def test_dtype_contract_and_timestep_checks():
    from lumina_t2x_b200 import transport
    from oracle import nextdit_oracle as O
    cfg, W, m = _tiny(max_tokens=256, max_cap_len=32)
    z, cap, mask = (v.cuda() for v in O.synthetic_inputs(cfg, (16, 16), 16, 8, seed=3))
    kw = dict(cfg_scale=2.0)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        out32 = m.forward_with_cfg(z.float(), torch.full((2,), 0.5, device="cuda"), cap, mask, **kw)
    assert out32.dtype == torch.float32 and any("computes in bfloat16" in str(w.message) for w in rec)
    out16 = m.forward_with_cfg(z, torch.full((2,), 0.5, device="cuda"), cap, mask, **kw)
    assert torch.equal(out32, out16.float())
    with pytest.raises(ValueError):          # per-row timesteps are not what forward_with_cfg is called with (transport.py:106)
        m.forward_with_cfg(z, torch.tensor([0.1, 0.9], device="cuda"), cap, mask, **kw)
    with pytest.raises(RuntimeError):        # conditioning on another device
        m.forward_with_cfg(z, torch.full((2,), 0.5, device="cuda"), cap.cpu(), mask, **kw)
    with torch.inference_mode():             # inference tensors have no version counter (the caption cache must cope)
        zi, ci, mi = z.clone(), cap.clone(), mask.clone()
        o = m.forward_with_cfg(zi, torch.full((2,), 0.5, device="cuda"), ci, mi, **kw)
        assert torch.equal(o, out16)
    # an fp32 state keeps fp32 time stepping (generic loop): t is not rounded to bf16 as it is for a bf16 state
    fn = transport.Sampler(transport.create_transport("Linear", "velocity", None, None, None)).sample_ode(
        sampling_method="euler", num_steps=4, atol=1e-6, rtol=1e-3, reverse=False, time_shifting_factor=1.0)
    tr32 = fn(z.float(), m.forward_with_cfg, cap_feats=cap, cap_mask=mask, **kw)
    assert tr32.dtype == torch.float32 and torch.isfinite(tr32).all()


def test_rk4_in_engine_equals_generic_loop():
    """sampling_method="rk4" (torchdiffeq's fixed-grid 3/8 rule): the in-engine solve (ndit_sample, NDIT_RK4) must give the bits
    of the PyTorch-driven loop over the same engine (transport._fixed_grid_torch: one forward_with_cfg per stage, every
    tensor op in bf16)."""
    from lumina_t2x_b200 import transport
    from oracle import nextdit_oracle as O
    cfg, W, m = _tiny(max_tokens=1024, max_cap_len=64)
    z, cap, mask = (v.cuda() for v in O.synthetic_inputs(cfg, (32, 32), 24, 8, seed=4))
    kw = dict(cfg_scale=2.0, scale_factor=1.0, scale_watershed=1.0, base_seqlen=64, proportional_attn=True)
    fn = transport.Sampler(transport.create_transport("Linear", "velocity", None, None, None)).sample_ode(
        sampling_method="rk4", num_steps=5, atol=1e-6, rtol=1e-3, reverse=False, time_shifting_factor=4.0)
    n0 = m.launch_count()
    fused = fn(z, m.forward_with_cfg, cap_feats=cap, cap_mask=mask, **kw)
    assert m.launch_count() - n0 > 4 * 4 * 20          # 4 intervals x 4 model calls inside the engine
    generic = transport._fixed_grid_torch(
        lambda t, x: m.forward_with_cfg(x, torch.ones(2, device="cuda") * t, cap, mask, **kw),
        z, transport._time_grid(0, 1, 5, 4.0).cuda(), "rk4")
    assert fused.shape == generic.shape == (5, 2, 4, 32, 32)
    assert torch.equal(fused, generic)


def test_packed_weight_file_round_trip(tmp_path):
    """checkpoint.save_packed / load_packed (ndit_save_packed / ndit_load_packed): a cold start from the packed file gives the
    bits of the engine that was built from the state dict; a file of another architecture is refused."""
    from lumina_t2x_b200 import checkpoint, models
    from oracle import nextdit_oracle as O
    cfg, W, m = _tiny(max_tokens=256, max_cap_len=32)
    z, cap, mask = (v.cuda() for v in O.synthetic_inputs(cfg, (16, 16), 16, 8, seed=3))
    t = torch.full((2,), 0.5, device="cuda")
    ref = m.forward_with_cfg(z, t, cap, mask, 2.0)
    path = str(tmp_path / "tiny.nditpk")
    checkpoint.save_packed(m, path)
    cold = models.NextDiT(dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, n_kv_heads=cfg.n_kv_heads, qk_norm=True,
                          cap_feat_dim=cfg.cap_feat_dim, max_tokens=64, max_cap_len=8)        # never sees a state dict
    checkpoint.load_packed(cold, path)
    n0 = cold.launch_count()
    out = cold.forward_with_cfg(z, t, cap, mask, 2.0)
    assert torch.equal(out, ref)
    assert n0 == 0                                       # no re-packing launches on the cold start
    cold.to("cuda")                                      # a move re-creates the engine from the file, not from the (unset) parameters
    assert torch.equal(cold.forward_with_cfg(z, t, cap, mask, 2.0), ref)
    other = models.NextDiT(dim=cfg.dim, n_layers=cfg.n_layers + 1, n_heads=cfg.n_heads, n_kv_heads=cfg.n_kv_heads, qk_norm=True,
                           cap_feat_dim=cfg.cap_feat_dim)
    with pytest.raises(RuntimeError):
        checkpoint.load_packed(other, path)
