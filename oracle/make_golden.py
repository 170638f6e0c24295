"""Generate tests/golden/*.pt by running the UNMODIFIED reference on CPU.

Run in the authoring container only (needs /root/reference):

    python oracle/make_golden.py

Each fixture stores the seeds/config needed to rebuild inputs with
``oracle.nextdit_oracle.synthetic_*`` plus the reference's outputs, so the committed
files stay small (outputs only).  Reference entry points exercised:
``models.NextDiT(...)`` ctor + ``load_state_dict(strict=True)`` (pins the state-dict key
names/shapes), ``forward_with_cfg`` (models/nextdit.py:838-885) in fp32 and under
``torch.autocast("cpu", bf16)``, per-block outputs via forward hooks, and
``transport.ODE(...).sample`` (transport.py:57-111) for euler and midpoint.
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import nextdit_oracle as O  # noqa: E402
from oracle import dit_llama_oracle as DL  # noqa: E402
from oracle.harness.ref_import import import_reference_imagenet, import_reference_mini  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = {
    # name: (latent_hw, T, uncond_len, kwargs for forward_with_cfg, t value)
    "tiny_prop": dict(hw=(16, 16), T=16, ul=8, t=0.25,
                      kw=dict(cfg_scale=2.0, scale_factor=1.0, scale_watershed=1.0, base_seqlen=64, proportional_attn=True)),
    "tiny_ntk": dict(hw=(16, 24), T=24, ul=8, t=0.6,
                     kw=dict(cfg_scale=4.0, scale_factor=2.0, scale_watershed=0.3, base_seqlen=16, proportional_attn=True)),
    "tiny_lin": dict(hw=(24, 16), T=16, ul=16, t=0.1,
                     kw=dict(cfg_scale=1.0, scale_factor=2.0, scale_watershed=0.3, base_seqlen=None, proportional_attn=False)),
}


def build_ref(models, cfg: O.NextDiTConfig, W):
    m = models.nextdit.NextDiT(patch_size=cfg.patch_size, in_channels=cfg.in_channels, dim=cfg.dim, n_layers=cfg.n_layers,
                       n_heads=cfg.n_heads, n_kv_heads=cfg.n_kv_heads, qk_norm=True, cap_feat_dim=cfg.cap_feat_dim,
                       use_flash_attn=False)
    m.load_state_dict({k: v.clone() for k, v in W.items()}, strict=True)
    return m.eval()


def main() -> None:
    models, transport = import_reference_mini()
    os.makedirs(OUT, exist_ok=True)
    cfg = O.config_tiny(n_layers=2)
    W = O.synthetic_weights(cfg, seed=0, dtype=torch.bfloat16)
    torch.set_grad_enabled(False)

    ref32 = build_ref(models, cfg, W).float()
    ref16 = build_ref(models, cfg, W).to(torch.bfloat16)

    for name, c in CASES.items():
        z, cap, mask = O.synthetic_inputs(cfg, c["hw"], c["T"], c["ul"], seed=1)
        t = torch.full((2,), c["t"], dtype=torch.float32)
        taps = {}
        hooks = [l.register_forward_hook(lambda mod, i, o, idx=idx: taps.__setitem__(f"block{idx}", o.clone()))
                 for idx, l in enumerate(ref32.layers)]
        out32 = ref32.forward_with_cfg(z.float(), t, cap.float(), mask, **c["kw"])
        for h in hooks:
            h.remove()
        with torch.autocast("cpu", dtype=torch.bfloat16):
            out16 = ref16.forward_with_cfg(z, t, cap, mask, **c["kw"])
        fx = dict(case=name, cfg=dict(dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, n_kv_heads=cfg.n_kv_heads,
                                      cap_feat_dim=cfg.cap_feat_dim),
                  hw=c["hw"], T=c["T"], ul=c["ul"], t=c["t"], kw=c["kw"], weight_seed=0, input_seed=1,
                  out_fp32=out32.clone(), out_autocast_cpu_bf16=out16.float().to(torch.bfloat16),
                  taps={k: v.to(torch.bfloat16) for k, v in taps.items()})
        torch.save(fx, os.path.join(OUT, f"fwd_{name}.pt"))
        print(name, "out32 absmax", out32.abs().max().item(), "autocast diff", (out16.float() - out32).abs().max().item())

    # trajectories through the reference's ODE class (fp32 state)
    c = CASES["tiny_prop"]
    z, cap, mask = O.synthetic_inputs(cfg, c["hw"], c["T"], c["ul"], seed=1)
    for method, steps, shift in (("euler", 5, 1.0), ("midpoint", 4, 4.0)):
        ode = transport.ODE(steps, method, shift)
        traj = ode.sample(z.float(), ref32.forward_with_cfg, cap_feats=cap.float(), cap_mask=mask, **c["kw"])
        fx = dict(method=method, num_steps=steps, time_shifting_factor=shift, hw=c["hw"], T=c["T"], ul=c["ul"], kw=c["kw"],
                  weight_seed=0, input_seed=1, grid=ode.t.clone(), traj_fp32=traj.clone())
        torch.save(fx, os.path.join(OUT, f"traj_{method}.pt"))
        print(method, "traj", tuple(traj.shape), traj[-1].abs().max().item())

    # stepping semantics with a bf16 state and a toy velocity (pins the t / dt dtype handling of the solver glue)
    def toy(x, t, **kw):
        return (torch.sin(3.0 * x.float()) * (1.0 + t.float().view(-1, 1, 1, 1))).to(x.dtype)

    zb = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(7)).to(torch.bfloat16)
    for method, steps, shift in (("euler", 7, 4.0), ("midpoint", 6, 4.0)):
        traj = transport.ODE(steps, method, shift).sample(zb, toy)
        torch.save(dict(method=method, num_steps=steps, time_shifting_factor=shift, z=zb, traj_bf16=traj),
                   os.path.join(OUT, f"toy_{method}_bf16.pt"))
        print("toy", method, traj.dtype, tuple(traj.shape))

    make_imagenet()


def make_imagenet() -> None:
    """Class-conditional Next-DiT (BASELINE config 1 / SURVEY 8a14): unmodified Next-DiT-ImageNet/models/models.py (fp32,
    CPU, fairscale at world size 1).  Fixtures: two tiny models (head_dim 48 and 72, default and scaled RoPE) and the
    config-1 case itself - DiT_Llama_600M_patch2, 256x256 image = latent 32x32, label 207, one forward_with_cfg."""
    ref = import_reference_imagenet()
    torch.set_grad_enabled(False)
    cases = {
        "imagenet_tiny48": dict(cfg=DL.config_tiny48(), hw=(16, 16), labels=(3, 7), t=0.35, cfg_scale=3.0, rope=None),
        "imagenet_tiny72_rope": dict(cfg=DL.config_tiny72(), hw=(16, 24), labels=(1,), t=0.8, cfg_scale=1.5, rope=(2.0, 1.5)),
        "imagenet_600m_config1": dict(cfg=DL.config_600m(), hw=(32, 32), labels=(207,), t=0.0, cfg_scale=4.0, rope=None),
    }
    for name, c in cases.items():
        cfg = c["cfg"]
        W = DL.synthetic_weights(cfg, seed=0)
        m = ref.DiT_Llama(input_size=c["hw"][0], patch_size=2, dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads,
                          num_classes=cfg.num_classes, qk_norm=True)
        m.load_state_dict({k: v.float() for k, v in W.items()}, strict=True)
        m = m.eval().float()
        z, y = DL.synthetic_inputs(cfg, c["hw"], c["labels"], seed=1)
        t = torch.full((len(z),), c["t"])
        kw = {} if c["rope"] is None else dict(rope_scaling_factor=c["rope"][0], ntk_factor=c["rope"][1])
        out = m.forward_with_cfg(z.float(), t, y, c["cfg_scale"], **kw)
        fx = dict(case=name, cfg=dict(dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, num_classes=cfg.num_classes),
                  hw=c["hw"], labels=c["labels"], t=c["t"], cfg_scale=c["cfg_scale"], rope=c["rope"], weight_seed=0, input_seed=1,
                  out_fp32=out.clone())
        torch.save(fx, os.path.join(OUT, f"{name}.pt"))
        print(name, tuple(out.shape), "absmax", out.abs().max().item())
        del m, W


def make_imagenet_plain_forward() -> None:
    """DiT_Llama.forward of the class-conditional model (models.py:920-944): odd batch, one timestep and label per row, fresh module
    and after a forward_with_cfg call that left scaled rope factors on the module."""
    ref = import_reference_imagenet()
    torch.set_grad_enabled(False)
    cfg = DL.config_tiny72()
    W = DL.synthetic_weights(cfg, seed=0)
    g = torch.Generator().manual_seed(17)
    x = torch.randn(3, cfg.in_channels, 16, 24, generator=g).to(torch.bfloat16)
    t = torch.tensor([0.1, 0.55, 0.9])
    y = torch.tensor([2, 9, cfg.num_classes])
    fx = dict(cfg=dict(dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, num_classes=cfg.num_classes), weight_seed=0, x=x, t=t, y=y,
              sticky_call=dict(rope_scaling_factor=2.0, ntk_factor=1.5, cfg_scale=2.0, hw=(16, 16), labels=(1,), seed=5, t=0.3))
    for state in ("fresh", "sticky"):
        m = ref.DiT_Llama(input_size=16, patch_size=2, dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, num_classes=cfg.num_classes, qk_norm=True)
        m.load_state_dict({k: v.float() for k, v in W.items()}, strict=True)
        m = m.eval().float()
        kw = {}
        if state == "sticky":
            z2, y2 = DL.synthetic_inputs(cfg, (16, 16), (1,), seed=5)
            m.forward_with_cfg(z2.float(), torch.full((len(z2),), 0.3), y2, 2.0, rope_scaling_factor=2.0, ntk_factor=1.5)
            kw = dict(rope_scaling_factor=2.0, ntk_factor=1.5)
        out = m(x.float(), t, y)
        o = DL.forward(cfg, W, x.float(), t, y, precision="fp32", **kw)
        ob = DL.forward(cfg, W, x, t, y, precision="bf16", **kw)
        fx[state] = dict(out_fp32=out.clone())
        print(state, tuple(out.shape), "absmax", out.abs().max().item(), "oracle fp32 rel", ((o - out).abs().max() / out.abs().max()).item(),
              "oracle bf16 rel", ((ob - out).abs().max() / out.abs().max()).item())
    torch.save(fx, os.path.join(OUT, "imagenet_plain_forward.pt"))


def make_moe() -> None:
    """Next-DiT-MoE (BASELINE config 5 / SURVEY 8a16): the unmodified time / space / both MoE models (fp32, CPU)."""
    from oracle.harness.ref_import import import_reference_moe
    torch.set_grad_enabled(False)
    cases = {
        "moe_tiny_time": dict(moe="time", file="models", hw=(16, 16), labels=(3,), t=0.35, cfg_scale=3.0),
        "moe_tiny_space": dict(moe="space", file="models1", hw=(16, 16), labels=(7,), t=0.6, cfg_scale=2.0),
        "moe_tiny_both": dict(moe="both", file="models2", hw=(16, 24), labels=(1,), t=0.8, cfg_scale=1.5),
    }
    for name, c in cases.items():
        ref = import_reference_moe(c["file"])
        cfg = DL.config_tiny_moe(c["moe"])
        W = DL.synthetic_weights(cfg, seed=0)
        m = ref.DiT_Llama(input_size=c["hw"][0], patch_size=2, dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads,
                          num_classes=cfg.num_classes, qk_norm=True)
        m.load_state_dict({k: v.float() for k, v in W.items()}, strict=True)
        m = m.eval().float()
        z, y = DL.synthetic_inputs(cfg, c["hw"], c["labels"], seed=1)
        t = torch.full((len(z),), c["t"])
        out = m.forward_with_cfg(z.float(), t, y, c["cfg_scale"])
        fx = dict(case=name, cfg=dict(dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, num_classes=cfg.num_classes, moe=cfg.moe),
                  hw=c["hw"], labels=c["labels"], t=c["t"], cfg_scale=c["cfg_scale"], rope=None, weight_seed=0, input_seed=1,
                  out_fp32=out.clone())
        torch.save(fx, os.path.join(OUT, f"{name}.pt"))
        o = DL.forward_with_cfg(cfg, W, z.float(), t, y, c["cfg_scale"], precision="fp32")
        ob = DL.forward_with_cfg(cfg, W, z, t, y, c["cfg_scale"], precision="bf16")
        print(name, tuple(out.shape), "absmax", out.abs().max().item(), "oracle fp32 rel", ((o - out).abs().max() / out.abs().max()).item(),
              "bf16-mode rel", ((ob - out).abs().max() / out.abs().max()).item())
        del m, W


def make_canonical() -> None:
    """The canonical fairscale flavour ``lumina_next_t2i/models/model.py`` (what sample.py / demo.py import), fp32 on CPU.
    Its fp32 SDPA branch does not repeat the kv heads (model.py:407-417), so it only runs MHA models in fp32: a tiny MHA
    model, proportional attention + time-aware RoPE.  Also asserts that the mini flavour gives the same bits."""
    import dataclasses
    from oracle.harness.ref_import import import_reference_full_model
    full = import_reference_full_model()
    mini, _ = import_reference_mini()
    torch.set_grad_enabled(False)
    cfg = dataclasses.replace(O.config_tiny(n_layers=2), n_kv_heads=O.config_tiny().n_heads)
    W = O.synthetic_weights(cfg, seed=0, dtype=torch.bfloat16)
    outs = []
    for cls in (full.NextDiT, mini.nextdit.NextDiT):
        m = cls(patch_size=2, in_channels=4, dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, n_kv_heads=cfg.n_kv_heads,
                qk_norm=True, cap_feat_dim=cfg.cap_feat_dim)
        m.load_state_dict({k: v.clone() for k, v in W.items()}, strict=True)
        m = m.eval().float()
        z, cap, mask = O.synthetic_inputs(cfg, (16, 24), 24, 8, seed=1)
        t = torch.full((2,), 0.45)
        kw = dict(cfg_scale=3.0, scale_factor=2.0, scale_watershed=0.3, base_seqlen=32, proportional_attn=True)
        outs.append(m.forward_with_cfg(z.float(), t, cap.float(), mask, **kw))
    assert torch.equal(outs[0], outs[1]), "canonical and mini flavours differ"
    fx = dict(case="canon_tiny_mha", cfg=dict(dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, n_kv_heads=cfg.n_kv_heads,
                                              cap_feat_dim=cfg.cap_feat_dim),
              hw=(16, 24), T=24, ul=8, t=0.45, kw=kw, weight_seed=0, input_seed=1, out_fp32=outs[0].clone())
    torch.save(fx, os.path.join(OUT, "canon_tiny_mha.pt"))
    print("canon_tiny_mha", tuple(outs[0].shape), "absmax", outs[0].abs().max().item(), "mini == canonical: True")


def make_sde() -> None:
    """Stochastic sampler of the full transport package (transport.py:285-344) with a toy velocity function: pins the step
    formulas, the noise consumption order of the global torch RNG and the last-step variants."""
    from oracle.harness.ref_import import import_reference_full_transport
    T = import_reference_full_transport()

    def toy(x, t, **kw):
        return torch.sin(3.0 * x) * (1.0 + t.view(-1, 1, 1, 1)) - 0.5 * x

    z = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(11))
    cases = [dict(sampling_method="Euler", diffusion_form="sigma", diffusion_norm=1.0, last_step="Mean", last_step_size=0.04, num_steps=8),
             dict(sampling_method="Heun", diffusion_form="linear", diffusion_norm=0.5, last_step="Euler", last_step_size=0.02, num_steps=6),
             dict(sampling_method="Euler", diffusion_form="decreasing", diffusion_norm=1.0, last_step="Tweedie", last_step_size=0.05, num_steps=5),
             dict(sampling_method="Euler", diffusion_form="inccreasing-decreasing", diffusion_norm=0.3, last_step=None, last_step_size=0.04,
                  num_steps=5)]      # ("constant" makes the reference itself fail: th.sqrt of a Python float)
    out = []
    for i, c in enumerate(cases):
        fn = T.Sampler(T.create_transport("Linear", "velocity")).sample_sde(**c)
        torch.manual_seed(100 + i)
        xs = fn(z.clone(), toy)
        out.append(dict(kw=c, seed=100 + i, xs=torch.stack(xs)))
        print("sde", c["sampling_method"], c["diffusion_form"], c["last_step"], tuple(out[-1]["xs"].shape), out[-1]["xs"][-1].abs().max().item())
    torch.save(dict(z=z, cases=out), os.path.join(OUT, "toy_sde.pt"))


def make_transport_table() -> None:
    """Host logic of the full transport package: create_transport's eps selection and Transport.check_interval for every
    combination the mirror accepts (transport/__init__.py:4-66, transport.py:67-93)."""
    import itertools
    from oracle.harness.ref_import import import_reference_full_transport
    T = import_reference_full_transport()
    rows = []
    for path, pred, lw, te, se in itertools.product(("Linear", "GVP", "VP"), ("velocity", "noise", "score"), (None, "velocity", "likelihood"),
                                                    (None, 1e-4), (None, 2e-3)):
        tr = T.create_transport(path, pred, lw, te, se)
        for form, sde, rev, ev, lss in itertools.product(("SBDM", "sigma"), (False, True), (False, True), (False, True), (0.0, 0.04)):
            row = dict(args=(path, pred, lw, te, se), kw=dict(diffusion_form=form, sde=sde, reverse=rev, eval=ev, last_step_size=lss),
                       train_eps=tr.train_eps, sample_eps=tr.sample_eps)
            try:        # some combinations make the reference itself fail (eps left at None): the error type is part of the contract
                t0, t1 = tr.check_interval(tr.train_eps, tr.sample_eps, **row["kw"])
                row.update(t0=None if t0 is None else float(t0), t1=None if t1 is None else float(t1), error=None)
            except Exception as ex:
                row.update(t0=None, t1=None, error=type(ex).__name__)
            rows.append(row)
    torch.save(rows, os.path.join(OUT, "transport_table.pt"))
    print("transport table", len(rows), "rows")


def make_signatures() -> None:
    """Constructor / forward_with_cfg signatures and factory names of the four reference model packages (drop-in surface)."""
    import inspect
    import json
    from oracle.harness.ref_import import import_reference_compositional_model, import_reference_flag_dit, import_reference_moe
    mods = {"compositional": import_reference_compositional_model(), "next_t2i_mini": import_reference_mini()[0].nextdit, "imagenet": import_reference_imagenet(), "lumina_t2i": import_reference_flag_dit(),
            "moe_time": import_reference_moe("models"), "moe_space": import_reference_moe("models1"), "moe_both": import_reference_moe("models2")}

    def sig(fn):
        out = []
        for p in inspect.signature(fn).parameters.values():
            if p.kind in (p.VAR_KEYWORD, p.VAR_POSITIONAL):
                continue
            out.append([p.name, None if p.default is p.empty else repr(p.default)])
        return out

    table = {}
    for name, m in mods.items():
        cls = m.NextDiT if hasattr(m, "NextDiT") else m.DiT_Llama
        table[name] = {"class": cls.__name__, "init": sig(cls.__init__), "forward_with_cfg": sig(cls.forward_with_cfg),
                       "factories": sorted(n for n in dir(m) if n.startswith(cls.__name__ + "_") and callable(getattr(m, n)))}
    with open(os.path.join(OUT, "signatures.json"), "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print("signatures", {k: len(v["init"]) for k, v in table.items()})


def make_flag_dit() -> None:
    """Flag-DiT (Lumina-T2I, BASELINE config 4 / SURVEY 8a15): unmodified lumina_t2i/models/model.py (fp32, CPU, fairscale
    at world size 1).  Tiny models with the flagship head_dim 96: default call, proportional attention + NTK factor (the
    demo's high-resolution kwargs, lumina_t2i/demo.py:169-178), rope scaling on a non-square latent."""
    from oracle import flag_dit_oracle as FD
    from oracle.harness.ref_import import import_reference_flag_dit
    ref = import_reference_flag_dit()
    torch.set_grad_enabled(False)
    cases = {
        "flagdit_tiny_default": dict(hw=(16, 16), T=16, t=0.3, cfg_scale=4.0, kw={}),
        "flagdit_tiny_prop_ntk": dict(hw=(24, 24), T=24, t=0.7, cfg_scale=2.0,
                                      kw=dict(proportional_attn=True, base_seqlen=8 * 8 + 8 * 2, ntk_factor=2.25)),
        "flagdit_tiny_ropescale": dict(hw=(16, 32), T=8, t=0.05, cfg_scale=1.5, kw=dict(rope_scaling_factor=2.0, ntk_factor=1.0)),
    }
    cfg = FD.config_tiny96()
    W = FD.synthetic_weights(cfg, seed=0)
    for name, c in cases.items():
        m = ref.DiT_Llama(patch_size=2, dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, qk_norm=True,
                          cap_feat_dim=cfg.cap_feat_dim)
        m.load_state_dict({k: v.float() for k, v in W.items()}, strict=True)
        m = m.eval().float()
        z, cap, mask = FD.synthetic_inputs(cfg, c["hw"], T=c["T"], uncond_len=4, seed=1)
        t = torch.full((2,), c["t"])
        out = m.forward_with_cfg(z.float(), t, cap.float(), mask, c["cfg_scale"], **c["kw"])
        fx = dict(case=name, cfg=dict(dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, n_kv_heads=cfg.n_kv_heads,
                                      cap_feat_dim=cfg.cap_feat_dim),
                  hw=c["hw"], T=c["T"], uncond_len=4, t=c["t"], cfg_scale=c["cfg_scale"], kw=c["kw"], weight_seed=0, input_seed=1,
                  out_fp32=out.clone())
        torch.save(fx, os.path.join(OUT, f"{name}.pt"))
        o = FD.forward_with_cfg(cfg, W, z.float(), t, cap.float(), mask, c["cfg_scale"], precision="fp32", **c["kw"])
        ob = FD.forward_with_cfg(cfg, W, z, t, cap, mask, c["cfg_scale"], precision="bf16", **c["kw"])
        print(name, tuple(out.shape), "absmax", out.abs().max().item(), "oracle fp32 rel", ((o - out).abs().max() / out.abs().max()).item(),
              "bf16-mode rel", ((ob - out).abs().max() / out.abs().max()).item())
        del m


def make_plain_forward() -> None:
    """NextDiT.forward without guidance (nextdit.py:808-836 == lumina_next_t2i/models/model.py:836-864): odd batch, one timestep per
    row, (a) on a freshly constructed module (default rope table) and (b) after a forward_with_cfg call whose scale_factor /
    proportional attention settings stay on the module (self.freqs_cis, layer.attention.base_seqlen)."""
    models, _ = import_reference_mini()
    cfg = O.config_tiny(n_layers=2)
    W = O.synthetic_weights(cfg, seed=0, dtype=torch.bfloat16)
    torch.set_grad_enabled(False)
    g = torch.Generator().manual_seed(11)
    B, T, hw = 3, 24, (16, 24)
    x = torch.randn(B, cfg.in_channels, *hw, generator=g).to(torch.bfloat16)
    cap = torch.randn(B, T, cfg.cap_feat_dim, generator=g).to(torch.bfloat16)
    mask = torch.zeros(B, T, dtype=torch.int32)
    for b, n in enumerate((24, 9, 1)):
        mask[b, :n] = 1
    t = torch.tensor([0.15, 0.6, 0.95])
    sticky = dict(cfg_scale=2.0, scale_factor=2.0, scale_watershed=0.3, base_seqlen=64, proportional_attn=True)
    z2, cap2, mask2 = O.synthetic_inputs(cfg, (16, 16), 16, 4, seed=3)
    fx = dict(hw=hw, T=T, weight_seed=0, x=x, cap=cap, mask=mask, t=t, sticky_call=dict(kw=sticky, t=0.2, hw=(16, 16), T=16, ul=4, seed=3))
    for state in ("fresh", "sticky"):
        ref32 = build_ref(models, cfg, W).float()
        ref16 = build_ref(models, cfg, W).to(torch.bfloat16)
        if state == "sticky":
            for m_, dt in ((ref32, torch.float32), (ref16, torch.bfloat16)):
                m_.forward_with_cfg(z2.to(dt), torch.full((2,), 0.2), cap2.to(dt), mask2, **sticky)
        out32 = ref32(x.float(), t, cap.float(), mask)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            out16 = ref16(x, t, cap, mask)
        fx[state] = dict(out_fp32=out32.clone(), out_autocast_cpu_bf16=out16.float().to(torch.bfloat16))
        kw = dict() if state == "fresh" else dict(scale_factor=2.0, scale_watershed=0.3, rope_timestep=0.2, base_seqlen=64, proportional_attn=True)
        o = O.forward(cfg, W, x.float(), t, cap.float(), mask, precision="fp32", **kw)
        print(state, tuple(out32.shape), "absmax", out32.abs().max().item(), "oracle fp32 rel", ((o - out32).abs().max() / out32.abs().max()).item(),
              "ref bf16 vs fp32", ((out16.float() - out32).abs().max() / out32.abs().max()).item())
    torch.save(fx, os.path.join(OUT, "plain_forward.pt"))


def make_list_forward() -> None:
    """NextDiT.forward with a LIST of latents of different sizes (nextdit.py:761-806): pad token, per-image rope grid, masked keys.
    Recorded on a fresh module and with proportional attention left on the module by a forward_with_cfg call (the scale then sees
    the PADDED sequence length)."""
    models, _ = import_reference_mini()
    cfg = O.config_tiny(n_layers=2)
    W = O.synthetic_weights(cfg, seed=0, dtype=torch.bfloat16)
    torch.set_grad_enabled(False)
    g = torch.Generator().manual_seed(13)
    sizes = [(16, 16), (24, 32), (16, 24)]
    T = 16
    xs = [torch.randn(cfg.in_channels, hh, ww, generator=g).to(torch.bfloat16) for hh, ww in sizes]
    cap = torch.randn(len(sizes), T, cfg.cap_feat_dim, generator=g).to(torch.bfloat16)
    mask = torch.zeros(len(sizes), T, dtype=torch.int32)
    for b, n in enumerate((16, 5, 11)):
        mask[b, :n] = 1
    t = torch.tensor([0.3, 0.8, 0.55])
    sticky = dict(cfg_scale=2.0, scale_factor=2.0, scale_watershed=0.3, base_seqlen=64, proportional_attn=True)
    z2, cap2, mask2 = O.synthetic_inputs(cfg, (16, 16), 16, 4, seed=3)
    fx = dict(sizes=sizes, T=T, weight_seed=0, xs=xs, cap=cap, mask=mask, t=t, sticky_call=dict(kw=sticky, t=0.2, hw=(16, 16), T=16, ul=4, seed=3))
    for state in ("fresh", "sticky"):
        ref32 = build_ref(models, cfg, W).float()
        ref16 = build_ref(models, cfg, W).to(torch.bfloat16)
        if state == "sticky":
            for m_, dt in ((ref32, torch.float32), (ref16, torch.bfloat16)):
                m_.forward_with_cfg(z2.to(dt), torch.full((2,), 0.2), cap2.to(dt), mask2, **sticky)
        out32 = ref32([v.float() for v in xs], t, cap.float(), mask)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            out16 = ref16(xs, t, cap, mask)
        fx[state] = dict(out_fp32=[o.clone() for o in out32], out_autocast_cpu_bf16=[o.float().to(torch.bfloat16) for o in out16])
        kw = dict() if state == "fresh" else dict(scale_factor=2.0, scale_watershed=0.3, rope_timestep=0.2, base_seqlen=64, proportional_attn=True)
        o = O.forward_list(cfg, W, [v.float() for v in xs], t, cap.float(), mask, precision="fp32", **kw)
        for i, (a, b, c) in enumerate(zip(o, out32, out16)):
            print(state, i, tuple(b.shape), "absmax", b.abs().max().item(), "oracle fp32 rel", ((a - b).abs().max() / b.abs().max()).item(),
                  "ref bf16 vs fp32", ((c.float() - b).abs().max() / b.abs().max()).item())
    torch.save(fx, os.path.join(OUT, "list_forward.pt"))


def make_gemma() -> None:
    """Caption-encoder end: transformers' own GemmaModel (the reference's third-party text encoder, sample.py:46-50,111) with the
    oracle's seeded weights, small config with the real head_dim / activation; hidden_states[-2] in fp32 and in bf16 (CPU)."""
    from transformers import GemmaConfig, GemmaModel
    import transformers
    from oracle import gemma_oracle as G
    cfg = G.config_tiny()
    W = G.synthetic_weights(cfg, seed=0)
    hc = GemmaConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_hidden_layers,
                     num_attention_heads=cfg.num_attention_heads, num_key_value_heads=cfg.num_key_value_heads, head_dim=cfg.head_dim,
                     intermediate_size=cfg.intermediate_size, hidden_act="gelu_pytorch_tanh", rms_norm_eps=cfg.rms_norm_eps,
                     max_position_embeddings=1024)
    torch.set_grad_enabled(False)
    m32 = GemmaModel(hc).eval()
    m32.load_state_dict({k: v.float() for k, v in W.items()}, strict=True)
    m16 = GemmaModel(hc).eval().to(torch.bfloat16)
    m16.load_state_dict(W, strict=True)
    cases = []
    for B, T, seed in ((2, 24, 1), (3, 40, 2), (1, 8, 3)):
        ids, mask = G.synthetic_inputs(cfg, B, T, seed)
        h32 = m32(input_ids=ids, attention_mask=mask, output_hidden_states=True).hidden_states[-2]
        h16 = m16(input_ids=ids, attention_mask=mask, output_hidden_states=True).hidden_states[-2]
        o32 = G.hidden_states_m2(cfg, W, ids, mask, "fp32")
        o16 = G.hidden_states_m2(cfg, W, ids, mask, "bf16")
        valid = mask.bool()
        rel = lambda a, b: ((a.float() - b.float())[valid].abs().max() / b.float()[valid].abs().max()).item()
        print(f"gemma B={B} T={T}: |h| max {h32.abs().max().item():.3f}  oracle fp32 vs HF fp32 {rel(o32, h32):.2e}  HF bf16 vs fp32 {rel(h16, h32):.2e}  "
              f"oracle bf16 vs HF fp32 {rel(o16, h32):.2e}")
        cases.append(dict(ids=ids, mask=mask, h_fp32=h32.clone(), h_bf16_cpu=h16.clone()))
    torch.save(dict(cfg=cfg.__dict__, weight_seed=0, transformers=transformers.__version__, cases=cases), os.path.join(OUT, "gemma_tiny.pt"))


VARIANT_CASES = {
    # ctor variants of the same NextDiT (models/nextdit.py:607-690): name -> (config overrides, call)
    "variant_noqknorm": dict(cfg=dict(qk_norm=False), hw=(16, 24), T=24, ul=8, t=0.6,
                             kw=dict(cfg_scale=4.0, scale_factor=2.0, scale_watershed=0.3, base_seqlen=16, proportional_attn=True)),
    # sd3-VAE latents (in_channels 16, lumina_next_t2i/train.py:323) with a custom FeedForward width (ffn_dim_multiplier 1.3:
    # int(1.3 * 1536) = 1996 -> 2048 instead of 1536)
    "variant_c16_ffn13": dict(cfg=dict(in_channels=16, ffn_dim_multiplier=1.3), hw=(16, 16), T=16, ul=8, t=0.25,
                              kw=dict(cfg_scale=2.0, scale_factor=1.0, scale_watershed=1.0, base_seqlen=64, proportional_attn=True)),
}


def make_variants() -> None:
    """qk_norm=False, in_channels=16 and ffn_dim_multiplier through the unmodified mini reference (fp32 and CPU-autocast bf16)."""
    import dataclasses
    models, _ = import_reference_mini()
    torch.set_grad_enabled(False)
    for name, c in VARIANT_CASES.items():
        cfg = dataclasses.replace(O.config_tiny(n_layers=2), **c["cfg"])
        W = O.synthetic_weights(cfg, seed=0, dtype=torch.bfloat16)
        outs = {}
        for dt in (torch.float32, torch.bfloat16):
            m = models.nextdit.NextDiT(patch_size=cfg.patch_size, in_channels=cfg.in_channels, dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads,
                                       n_kv_heads=cfg.n_kv_heads, ffn_dim_multiplier=cfg.ffn_dim_multiplier, qk_norm=cfg.qk_norm,
                                       cap_feat_dim=cfg.cap_feat_dim, use_flash_attn=False)
            m.load_state_dict({k: v.clone() for k, v in W.items()}, strict=True)       # pins the key set (no *_norm keys without qk_norm)
            m = m.eval().to(dt)
            z, cap, mask = O.synthetic_inputs(cfg, c["hw"], c["T"], c["ul"], seed=1)
            t = torch.full((2,), c["t"])
            if dt == torch.float32:
                outs[dt] = m.forward_with_cfg(z.float(), t, cap.float(), mask, **c["kw"])
            else:
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    outs[dt] = m.forward_with_cfg(z, t, cap, mask, **c["kw"])
        o = O.forward_with_cfg(cfg, W, z.float(), t, cap.float(), mask, precision="fp32", **c["kw"])
        out32, out16 = outs[torch.float32], outs[torch.bfloat16]
        fx = dict(case=name, cfg_overrides=c["cfg"], hw=c["hw"], T=c["T"], ul=c["ul"], t=c["t"], kw=c["kw"], weight_seed=0, input_seed=1,
                  out_fp32=out32.clone(), out_autocast_cpu_bf16=out16.float().to(torch.bfloat16))
        torch.save(fx, os.path.join(OUT, f"{name}.pt"))
        print(name, tuple(out32.shape), "absmax", out32.abs().max().item(), "oracle fp32 rel", ((o - out32).abs().max() / out32.abs().max()).item(),
              "ref bf16 vs fp32", ((out16.float() - out32).abs().max() / out32.abs().max()).item())


COMPOSITIONAL_CASES = {
    # name: latent hw, region captions, split, caption length, t, kwargs
    "comp_2x2": dict(hw=(32, 32), n_regions=4, hs=2, ws=2, T=16, t=0.35,
                     kw=dict(cfg_scale=3.0, scale_factor=1.0, scale_watershed=1.0, base_seqlen=64, proportional_attn=True)),
    # 1 x 3 split of a non-square latent whose width is not divisible by 3 (rightmost token column has no region -> zeros), NTK branch
    "comp_1x3": dict(hw=(16, 40), n_regions=3, hs=1, ws=3, T=24, t=0.7,
                     kw=dict(cfg_scale=2.0, scale_factor=2.0, scale_watershed=0.3, base_seqlen=32, proportional_attn=True)),
    # single region = plain text-to-image through the compositional code path, linear-interpolation RoPE branch
    "comp_1x1": dict(hw=(24, 16), n_regions=1, hs=1, ws=1, T=16, t=0.1,
                     kw=dict(cfg_scale=4.0, scale_factor=2.0, scale_watershed=0.3, base_seqlen=None, proportional_attn=False)),
}


def make_compositional() -> None:
    """Region-masked cross-attention: the UNMODIFIED ``lumina_next_compositional_generation/models/model.py`` in fp32 on CPU (its
    fp32 SDPA branch does not repeat kv heads, model.py:407-417, so n_kv_heads = n_heads), pinned against ``oracle/compositional_oracle.py``.
    2 x 2 regions exercise the reference's region_id formula ((i + 1) * (j + 1) - 1: rectangles (0, 1) and (1, 0) share caption 1,
    caption 2 owns nothing, rectangle (1, 1) -> caption 3)."""
    import dataclasses
    from oracle import compositional_oracle as CO
    from oracle.harness.ref_import import import_reference_compositional_model
    ref = import_reference_compositional_model()
    torch.set_grad_enabled(False)
    cfg = dataclasses.replace(O.config_tiny(n_layers=2), n_kv_heads=O.config_tiny().n_heads)
    W = O.synthetic_weights(cfg, seed=0, dtype=torch.bfloat16)
    m = ref.NextDiT(patch_size=2, in_channels=4, dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, n_kv_heads=cfg.n_kv_heads,
                    qk_norm=True, cap_feat_dim=cfg.cap_feat_dim)
    m.load_state_dict({k: v.clone() for k, v in W.items()}, strict=True)
    m = m.eval().float()
    for name, c in COMPOSITIONAL_CASES.items():
        z, cap, mask, gcap, gmask = CO.synthetic_inputs(cfg, c["hw"], c["n_regions"], c["T"], seed=21)
        t = torch.full((2,), c["t"])
        kw = dict(c["kw"], global_cap_feats=gcap.float(), global_cap_mask=gmask, h_split_num=c["hs"], w_split_num=c["ws"])
        out = m.forward_with_cfg(z.float(), t, cap.float(), mask, **kw)
        assert torch.isfinite(out).all()
        o = CO.forward_with_cfg(cfg, W, z.float(), t, cap.float(), mask, precision="fp32", **kw)
        o16 = CO.forward_with_cfg(cfg, W, z.float(), t, cap.float(), mask, precision="bf16", **kw)
        fx = dict(case=name, cfg=dict(dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, n_kv_heads=cfg.n_kv_heads, cap_feat_dim=cfg.cap_feat_dim),
                  hw=c["hw"], n_regions=c["n_regions"], hs=c["hs"], ws=c["ws"], T=c["T"], t=c["t"], kw=c["kw"], weight_seed=0, input_seed=21,
                  out_fp32=out.clone())
        torch.save(fx, os.path.join(OUT, f"{name}.pt"))
        print(name, tuple(out.shape), "absmax", out.abs().max().item(), "oracle fp32 rel", ((o - out).abs().max() / out.abs().max()).item(),
              "oracle bf16 vs ref fp32", ((o16 - out).abs().max() / out.abs().max()).item())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "compositional":
        make_compositional()
    elif len(sys.argv) > 1 and sys.argv[1] == "signatures":
        make_signatures()
    elif len(sys.argv) > 1 and sys.argv[1] == "variants":
        make_variants()
    elif len(sys.argv) > 1 and sys.argv[1] == "gemma":
        make_gemma()
    elif len(sys.argv) > 1 and sys.argv[1] == "imagenet_plain_forward":
        make_imagenet_plain_forward()
    elif len(sys.argv) > 1 and sys.argv[1] == "plain_forward":
        make_plain_forward()
    elif len(sys.argv) > 1 and sys.argv[1] == "list_forward":
        make_list_forward()
    else:
        main()
        make_plain_forward()
        make_list_forward()
        make_gemma()
        make_imagenet_plain_forward()
