"""CPU tests: the C ABI loads and exports every symbol include/ndit.h declares (no compute calls without a
GPU), the host-side transport mirror behaves like the reference's, and the data-parallel sharding helpers work
at world_size 2 over gloo."""
import os
import re
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_loads_and_exports_header_symbols():
    import __graft_entry__ as g
    g.build()
    from lumina_t2x_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "ndit.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ndit_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 18
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ndit_abi_version() == 5
    # the caption-encoder end has its own header
    theader = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ndit_text.h")).read(), flags=re.S)
    tdeclared = set(re.findall(r"\b(ntxt_[a-z_0-9]+)\s*\(", theader))
    assert tdeclared == set(_lib.TEXT_SIGNATURES), tdeclared ^ set(_lib.TEXT_SIGNATURES)
    for name in tdeclared:
        assert hasattr(lib, name), name
    # and so does the VAE-decode end
    vheader = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ndit_vae.h")).read(), flags=re.S)
    vdeclared = set(re.findall(r"\b(nvae_[a-z_0-9]+)\s*\(", vheader))
    assert vdeclared == set(_lib.VAE_SIGNATURES), vdeclared ^ set(_lib.VAE_SIGNATURES)
    for name in vdeclared:
        assert hasattr(lib, name), name


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_no_silent_fallback_without_gpu():
    import ctypes as C
    from lumina_t2x_b200 import _lib, models
    lib = _lib.load()
    cfg = _lib.NditConfig(576, 2, 8, 2, 256, 4, 2, 256, 1, 1e-5, 256, 32, 2)
    h = C.c_void_p()
    assert lib.ndit_create(C.byref(cfg), C.byref(h)) != 0          # error code, never abort()
    assert lib.ndit_last_error(None)
    m = models.NextDiT(dim=576, n_layers=1, n_heads=8, n_kv_heads=2, qk_norm=True, cap_feat_dim=256)
    with pytest.raises(RuntimeError):
        m.forward_with_cfg(torch.zeros(2, 4, 16, 16), torch.zeros(2), torch.zeros(2, 8, 256), torch.ones(2, 8), 2.0)


def test_state_dict_keys_match_reference_inventory():
    from lumina_t2x_b200 import models
    from oracle import nextdit_oracle as O
    cfg = O.config_tiny(2)
    m = models.NextDiT(dim=cfg.dim, n_layers=2, n_heads=cfg.n_heads, n_kv_heads=cfg.n_kv_heads, qk_norm=True, cap_feat_dim=cfg.cap_feat_dim)
    want = O.state_dict_shapes(cfg)          # pinned against the reference by make_golden (strict load)
    have = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert have == want
    assert m.parameter_count() == sum(int(torch.tensor(s).prod()) for s in want.values())
    assert hasattr(models, "NextDiT_2B_GQA_patch2") and hasattr(models, "NextDiT_2B_patch2")


def test_class_conditional_state_dict_keys_match_reference_inventory():
    from lumina_t2x_b200 import models
    from oracle import dit_llama_oracle as DL
    cfg = DL.config_tiny48()
    m = models.DiT_Llama(input_size=16, dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, num_classes=cfg.num_classes, qk_norm=True)
    want = DL.state_dict_shapes(cfg)         # pinned against the reference by make_golden.make_imagenet (strict load)
    have = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert have == want
    assert hasattr(models, "DiT_Llama_600M_patch2") and hasattr(models, "DiT_Llama_2B_patch2")
    with pytest.raises(RuntimeError):        # no GPU / no CUDA library -> loud failure, never a CPU fallback
        m.forward_with_cfg(torch.zeros(2, 4, 16, 16), torch.zeros(2), torch.tensor([1, cfg.num_classes]), 2.0)


@pytest.mark.parametrize("moe", ["time", "space", "both"])
def test_moe_state_dict_keys_match_reference_inventory(moe):
    from lumina_t2x_b200 import models
    from lumina_t2x_b200.models import moe as moe_models
    from oracle import dit_llama_oracle as DL
    cfg = DL.config_tiny_moe(moe)
    m = models.DiT_Llama(input_size=16, dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, num_classes=cfg.num_classes,
                         qk_norm=True, moe=moe)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == DL.state_dict_shapes(cfg)   # pinned by make_golden.make_moe
    for f in ("DiT_Llama_600M_patch2", "DiT_Llama_600M_patch2_Spatial", "DiT_Llama_600M_patch2_Both"):
        assert hasattr(moe_models, f)


def test_flag_dit_state_dict_keys_match_reference_inventory():
    from lumina_t2x_b200.models import lumina_t2i as models
    from oracle import flag_dit_oracle as FD
    cfg = FD.config_tiny96()
    m = models.DiT_Llama(dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, qk_norm=True, cap_feat_dim=cfg.cap_feat_dim)
    want = FD.state_dict_shapes(cfg)         # pinned against the reference by make_golden.make_flag_dit (strict load)
    have = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert have == want
    assert hasattr(models, "DiT_Llama_5B_patch2")
    with pytest.raises(RuntimeError):
        m.forward_with_cfg(torch.zeros(2, 4, 16, 16), torch.zeros(2), torch.zeros(2, 8, cfg.cap_feat_dim), torch.ones(2, 8), 2.0)


def test_create_transport_and_grid_semantics():
    from lumina_t2x_b200 import transport as T
    from oracle import nextdit_oracle as O
    tr = T.create_transport("Linear", "velocity", None, None, None)
    assert tr.train_eps == 0 and tr.sample_eps == 0 and tr.model_type == T.ModelType.VELOCITY
    assert tr.check_interval(0, 0, sde=False, eval=True, reverse=False, last_step_size=0.0) == (0, 1)
    tr2 = T.create_transport("Linear", "noise")
    assert tr2.train_eps == 1e-3 and tr2.sample_eps is None          # the reference's quirk, pinned by transport_table.pt
    for n, s in ((30, 1.0), (50, 4.0), (5, None)):
        assert torch.equal(T._time_grid(0, 1, n, s), O.time_grid(n, s))
    g = T._time_grid(0, 1, 30, 4.0)
    assert len(g) == 30 and g[0] == 0 and abs(g[-1] - 1) < 1e-6      # 30 points = 29 integration steps
    with pytest.raises(NotImplementedError):       # noise / score parameterisations are outside the mirrored ODE path
        T.Sampler(T.create_transport("Linear", "noise", None, 1e-3, 1e-3)).sample_ode(sampling_method="euler")


@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
def test_generic_fixed_grid_loop(method):
    """y' = -y: the PyTorch-driven loop (used for model functions that are not the engine) reproduces the
    textbook update formulas and returns every grid state."""
    from lumina_t2x_b200 import transport as T
    tr = T.create_transport("Linear", "velocity")
    calls = []

    def model(x, t, scale=1.0):
        calls.append(float(t[0]))
        return -scale * x

    fn = T.Sampler(tr).sample_ode(sampling_method=method, num_steps=11, time_shifting_factor=None)
    x0 = torch.ones(2, 3, dtype=torch.float64)
    out = fn(x0, model, scale=1.0)
    assert out.shape == (11, 2, 3)
    h = 0.1
    fac = {"euler": 1 - h, "midpoint": 1 - h + h * h / 2, "rk4": 1 - h + h**2 / 2 - h**3 / 6 + h**4 / 24}[method]
    assert torch.allclose(out[-1], torch.full_like(x0, fac ** 10), atol=1e-6)
    assert len(calls) == {"euler": 10, "midpoint": 20, "rk4": 40}[method]
    ode = T.ODE(11, method, None)
    assert torch.allclose(ode.sample(x0, model, scale=1.0), out)
    with pytest.raises(NotImplementedError):
        T.Sampler(tr).sample_ode(sampling_method="dopri5", num_steps=5)(x0, model)


def test_model_mirrors_keep_the_reference_signatures():
    """tests/golden/signatures.json (recorded from the reference by oracle/make_golden.py::make_signatures): every constructor
    and forward_with_cfg parameter of the reference classes exists in the mirror at the same position with the same default
    (the mirrors only append engine limits such as max_tokens), and every factory name resolves."""
    import inspect
    import json
    from lumina_t2x_b200 import models
    from lumina_t2x_b200.models import compositional, dit_llama, lumina_t2i, moe
    table = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "signatures.json")))
    mirrors = {"compositional": (compositional.NextDiT, compositional), "next_t2i_mini": (models.NextDiT, models), "imagenet": (dit_llama.DiT_Llama, models), "lumina_t2i": (lumina_t2i.DiT_Llama, lumina_t2i),
               "moe_time": (dit_llama.DiT_Llama, moe), "moe_space": (dit_llama.DiT_Llama, moe), "moe_both": (dit_llama.DiT_Llama, moe)}
    for name, ref in table.items():
        cls, pkg = mirrors[name]
        for meth in ("init", "forward_with_cfg"):
            mine = [p for p in inspect.signature(getattr(cls, "__init__" if meth == "init" else meth)).parameters.values()]
            for i, (pname, pdef) in enumerate(ref[meth]):
                assert mine[i].name == pname, (name, meth, i, pname, mine[i].name)
                if pdef is not None:
                    assert repr(mine[i].default) == pdef, (name, meth, pname, pdef, mine[i].default)
        for f in ref["factories"]:
            assert hasattr(pkg, f), (name, f)


def test_create_transport_and_check_interval_table_matches_reference():
    """3456 combinations of create_transport(path, prediction, loss_weight, train_eps, sample_eps) x check_interval(...) recorded
    from the unmodified reference package (tests/golden/transport_table.pt): same eps defaults (including the reference's
    sample_eps-stays-None quirk), same intervals, same error type where the reference itself fails."""
    from lumina_t2x_b200 import transport as T
    rows = torch.load(os.path.join(os.path.dirname(__file__), "golden", "transport_table.pt"), map_location="cpu", weights_only=False)
    assert len(rows) == 3456
    for r in rows:
        tr = T.create_transport(*r["args"])
        assert (tr.train_eps, tr.sample_eps) == (r["train_eps"], r["sample_eps"]), r
        try:
            t0, t1 = tr.check_interval(tr.train_eps, tr.sample_eps, **r["kw"])
            got = (None if t0 is None else float(t0), None if t1 is None else float(t1), None)
        except Exception as ex:
            got = (None, None, type(ex).__name__)
        assert got == (r["t0"], r["t1"], r["error"]), (r, got)


def test_sde_sampler_matches_reference_fixture():
    """Sampler.sample_sde (Euler-Maruyama / Heun, four diffusion forms, all last-step variants) against trajectories of the
    unmodified reference transport package on a toy velocity (tests/golden/toy_sde.pt, oracle/make_golden.py::make_sde):
    same formulas AND same consumption order of the global torch RNG -> bit-identical."""
    from lumina_t2x_b200 import transport as T
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "toy_sde.pt"), map_location="cpu", weights_only=False)

    def toy(x, t, **kw):
        return torch.sin(3.0 * x) * (1.0 + t.view(-1, 1, 1, 1)) - 0.5 * x

    for c in fx["cases"]:
        fn = T.Sampler(T.create_transport("Linear", "velocity")).sample_sde(**c["kw"])
        torch.manual_seed(c["seed"])
        xs = fn(fx["z"].clone(), toy)
        assert len(xs) == c["kw"]["num_steps"] and torch.isfinite(c["xs"]).all()
        assert torch.equal(torch.stack(xs), c["xs"]), c["kw"]
    with pytest.raises(NotImplementedError):
        T.Sampler(T.create_transport("Linear", "noise")).sample_sde()


def test_shard_ranges():
    from lumina_t2x_b200.parallel import shard_range, shard_sizes
    for total in (0, 1, 7, 8, 9, 64):
        for world in (1, 2, 4, 8):
            rs = [shard_range(total, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(shard_sizes(total, world)) - min(shard_sizes(total, world)) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lumina_t2x_b200.parallel import sample_sharded, shard_range

    def sample_one(i):                      # stands in for one ODE solve; depends only on the sample index
        g = torch.Generator().manual_seed(100 + i)
        return torch.randn(4, 8, 8, generator=g)

    out = sample_sharded(sample_one, total)
    q.put((rank, out, shard_range(total, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [5, 4, 1])
def test_data_parallel_gather_world2_gloo(total):
    """Rank-sharded results equal the single-process results for the same per-sample seeds (independent
    units => exact equality), including ragged shards."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = torch.stack([torch.randn(4, 8, 8, generator=torch.Generator().manual_seed(100 + i)) for i in range(total)])
    for rank, out, _ in res:
        assert torch.equal(out, expect), rank


def test_checkpoint_tooling_safetensors_and_convert(tmp_path):
    """checkpoint.convert mirrors `lumina_next convert` (entry_point.py:115-156); the mmap safetensors reader / writer written
    here must interoperate with the safetensors library the reference uses."""
    from safetensors.torch import load_file, save_file
    from lumina_t2x_b200 import checkpoint
    g = torch.Generator().manual_seed(0)
    sd = {"layers.0.attention.wq.weight": torch.randn(48, 32, generator=g).to(torch.bfloat16), "pad_token": torch.randn(32, generator=g),
          "layers.0.attention.gate": torch.zeros(4, dtype=torch.float16), "step": torch.tensor([7], dtype=torch.int64), "empty": torch.empty(0, 3)}
    p_lib, p_ours = str(tmp_path / "a.safetensors"), str(tmp_path / "b.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, p_lib)
    checkpoint.write_safetensors(sd, p_ours)
    for path, reader in ((p_lib, checkpoint.read_safetensors), (p_ours, lambda p: load_file(p, device="cpu"))):
        got = reader(path)
        assert set(got) == set(sd)
        for k, v in sd.items():
            assert got[k].dtype == v.dtype and got[k].shape == v.shape and torch.equal(got[k], v), k
    torch.save(sd, str(tmp_path / "consolidated.00-of-01.pth"))
    out = checkpoint.convert(str(tmp_path / "consolidated.00-of-01.pth"), str(tmp_path / "conv"))
    assert out.endswith("consolidated.00-of-01.safetensors")
    back = checkpoint.convert(out, str(tmp_path / "conv2"))
    assert back.endswith("consolidated.00-of-01.pth")
    sd2 = torch.load(back, map_location="cpu", weights_only=True)
    assert all(torch.equal(sd2[k], sd[k]) for k in sd)
    with pytest.raises(ValueError):
        checkpoint.convert(str(tmp_path / "x.bin"), str(tmp_path))


def test_compositional_mirror_host_logic():
    """models.compositional: same parameter inventory as the plain NextDiT (the reference class differs only in forward code), the
    sampler routes its bound forward_with_cfg to the in-engine solve with exactly the reference's keyword set (signatures.json), the
    reference's own failure modes are kept (missing global caption, region id beyond the caption rows), and there is no CPU path."""
    import inspect
    import json
    from lumina_t2x_b200 import models, transport
    from lumina_t2x_b200.models import compositional
    from oracle import nextdit_oracle as O
    cfg = O.config_tiny(2)
    kw = dict(dim=cfg.dim, n_layers=2, n_heads=cfg.n_heads, n_kv_heads=cfg.n_kv_heads, qk_norm=True, cap_feat_dim=cfg.cap_feat_dim)
    m = compositional.NextDiT(**kw)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == O.state_dict_shapes(cfg)
    assert isinstance(m, models.NextDiT) and transport._engine_of(m.forward_with_cfg) is m
    allowed, required = transport._ENGINE_KW[compositional.NextDiT]
    table = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "signatures.json")))
    ref_kw = [p for p, _ in table["compositional"]["forward_with_cfg"]][3:]          # after self, x, t
    assert list(allowed) == ref_kw and set(required) <= set(allowed)
    assert [p for p in inspect.signature(m.forward).parameters] == ["x", "t", "cap_feats", "cap_mask", "global_cap_feats", "global_cap_mask",
                                                                  "h_split_num", "w_split_num"]
    x, t = torch.zeros(2, 4, 16, 16), torch.zeros(2)
    cap, mask = torch.zeros(3, 8, cfg.cap_feat_dim), torch.ones(3, 8)
    with pytest.raises(AttributeError):        # the reference dereferences global_cap_mask unconditionally (model.py:866)
        m.forward_with_cfg(x, t, cap, mask, 2.0)
    with pytest.raises(ValueError):            # one cond / uncond pair
        m.forward_with_cfg(torch.zeros(4, 4, 16, 16), t, cap, mask, 2.0, global_cap_feats=cap[:1], global_cap_mask=mask[:1])
    with pytest.raises(RuntimeError):          # CPU tensors: no fallback
        m.forward_with_cfg(x, t, cap, mask, 2.0, global_cap_feats=cap[:1], global_cap_mask=mask[:1])
    with pytest.raises(NotImplementedError):   # head_dim 48: the region kernel is instantiated for head_dim 72
        compositional.NextDiT(dim=384, n_layers=1, n_heads=8, qk_norm=True, cap_feat_dim=64)._check_supported()


def test_nextdit_ctor_variants_host_logic():
    """qk_norm=False drops the *_norm keys, ffn_dim_multiplier follows model.py:470-473 in Python arithmetic, in_channels sizes the
    embedder / final layer; the engine config carries them (ffn_dim, no_qk_norm)."""
    import dataclasses
    from lumina_t2x_b200 import models
    from oracle import nextdit_oracle as O
    for over in (dict(qk_norm=False), dict(in_channels=16, ffn_dim_multiplier=1.3), dict(ffn_dim_multiplier=0.5)):
        cfg = dataclasses.replace(O.config_tiny(2), **over)
        m = models.NextDiT(dim=cfg.dim, n_layers=2, n_heads=cfg.n_heads, n_kv_heads=cfg.n_kv_heads, qk_norm=cfg.qk_norm, in_channels=cfg.in_channels,
                           ffn_dim_multiplier=cfg.ffn_dim_multiplier, cap_feat_dim=cfg.cap_feat_dim)
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == O.state_dict_shapes(cfg), over
        c = m._ndit_config()
        assert c.ffn_dim == cfg.ffn_dim and c.no_qk_norm == (0 if cfg.qk_norm else 1) and c.in_channels == cfg.in_channels
        m._check_supported()
    with pytest.raises(NotImplementedError):
        models.NextDiT(dim=576, n_layers=1, n_heads=8, qk_norm=True, cap_feat_dim=64, in_channels=3)._check_supported()


def test_region_assignment_formula_of_the_kernel_matches_the_reference_mask():
    """attention_fused_kernel<72, REGION> / attention_ref_kernel pick ONE caption per query token with integer arithmetic
    (attention_tcgen05.cu: hs = (row / Wp) / hp, ws = (row % Wp) / wp, id = (hs + 1) * (ws + 1) - 1, valid iff hs < h_split, ws < w_split,
    id < n_cond; hp = Hp / h_split, wp = Wp / w_split as set up in engine.cu).  Restated here and checked against the boolean region
    mask of the oracle (= model.py:872-887, pinned by the comp_*.pt fixtures) over many grids: for every cond caption row r and token n,
    mask[r, n] == (caption_of(n) == r); at most one caption per token; the last row is all ones."""
    from hypothesis import given, settings
    from hypothesis import strategies as st
    from oracle import compositional_oracle as CO

    def caption_of(n, Hp, Wp, hs, ws, n_cond):
        hp, wp = Hp // hs, Wp // ws
        i, j = (n // Wp) // hp, (n % Wp) // wp
        rid = (i + 1) * (j + 1) - 1
        return rid if (i < hs and j < ws and rid < n_cond) else -1

    @settings(max_examples=150, deadline=None)
    @given(st.integers(1, 24), st.integers(1, 24), st.integers(1, 4), st.integers(1, 4), st.integers(0, 3))
    def check(Hp, Wp, hs, ws, extra):
        if hs > Hp or ws > Wp:
            return                                        # the engine rejects regions that do not fit the token grid
        n_caps = hs * ws + extra                          # the smallest the reference accepts is h_split * w_split rows (region id < rows)
        n_cond = n_caps - 1
        mask = CO.region_mask(n_caps, Hp, Wp, hs, ws)
        assert mask[-1].all()
        owner = torch.tensor([caption_of(n, Hp, Wp, hs, ws, n_cond) for n in range(Hp * Wp)])
        for r in range(n_cond):
            assert torch.equal(mask[r], owner == r), (Hp, Wp, hs, ws, n_caps, r)
        assert (mask[:-1].sum(0) <= 1).all()

    check()
