"""Drop-in mirror of the reference ``transport`` package for ODE sampling.

Mirrors ``lumina_next_t2i/transport/__init__.py:4-66`` (create_transport), ``transport.py:221-391``
(Sampler.sample_ode), ``integrators.py:79-116`` (the ``ode`` helper) and the mini flavour's
``lumina_next_t2i_mini/transport.py:57-111`` (``ODE``).  Fixed-grid ``euler`` / ``midpoint`` solves whose model
function is ``NextDiT.forward_with_cfg`` of the B200 engine run entirely inside libndit_b200.so
(``ndit_sample``); any other model function is driven by the same fixed-grid loop in PyTorch.
``Sampler.sample_sde`` (``transport.py:285-344``, ``integrators.py:5-76``, ``path.py`` ICPlan) is mirrored for the
velocity-prediction / Linear-path configuration as a host loop around the model function: the
stochastic sampler draws fresh noise on the host every step, so there is nothing to fuse.  Training losses and
likelihoods are out of scope.
"""
from __future__ import annotations

import enum
import math
from typing import Callable, Optional

import torch as th

from ..models.dit_llama import DiT_Llama
from ..models.lumina_t2i import DiT_Llama as FlagDiT
from ..models.nextdit import NextDiT
from ..models.compositional import NextDiT as CompositionalNextDiT

__all__ = ["create_transport", "Sampler", "Transport", "ModelType", "PathType", "WeightType", "ODE"]


class ModelType(enum.Enum):
    NOISE = enum.auto()
    SCORE = enum.auto()
    VELOCITY = enum.auto()


class PathType(enum.Enum):
    LINEAR = enum.auto()
    GVP = enum.auto()
    VP = enum.auto()


class WeightType(enum.Enum):
    NONE = enum.auto()
    VELOCITY = enum.auto()
    LIKELIHOOD = enum.auto()


class Transport:
    def __init__(self, *, model_type, path_type, loss_type, train_eps, sample_eps, snr_type="uniform"):
        self.model_type, self.path_type, self.loss_type = model_type, path_type, loss_type
        self.train_eps, self.sample_eps, self.snr_type = train_eps, sample_eps, snr_type

    def check_interval(self, train_eps, sample_eps, *, diffusion_form="SBDM", sde=False, reverse=False, eval=False,
                       last_step_size=0.0):
        """transport.py:67-93."""
        t0, t1 = 0, 1
        eps = train_eps if not eval else sample_eps
        if self.path_type == PathType.VP:
            t1 = 1 - eps if (not sde or last_step_size == 0) else 1 - last_step_size
        elif self.model_type != ModelType.VELOCITY or sde:
            t0 = eps if (diffusion_form == "SBDM" and sde) or self.model_type != ModelType.VELOCITY else 0
            t1 = 1 - eps if (not sde or last_step_size == 0) else 1 - last_step_size
        if reverse:
            t0, t1 = 1 - t0, 1 - t1
        return t0, t1

    def training_losses(self, *a, **k):
        raise NotImplementedError("training is out of scope for the B200 sampling engine")


def create_transport(path_type="Linear", prediction="velocity", loss_weight=None, train_eps=None, sample_eps=None,
                     snr_type="uniform") -> Transport:
    """transport/__init__.py:4-66 (same defaults and eps selection, including its use of ``train_eps`` in the
    ``sample_eps`` test)."""
    model_type = {"noise": ModelType.NOISE, "score": ModelType.SCORE}.get(prediction, ModelType.VELOCITY)
    loss_type = {"velocity": WeightType.VELOCITY, "likelihood": WeightType.LIKELIHOOD}.get(loss_weight, WeightType.NONE)
    ptype = {"Linear": PathType.LINEAR, "GVP": PathType.GVP, "VP": PathType.VP}[path_type]
    # the statements are sequential on purpose: like the reference, the sample_eps line tests train_eps AFTER it has been
    # given its default, so a sample_eps left at None stays None (and check_interval(eval=True) then fails like the reference's)
    if ptype == PathType.VP:
        train_eps = 1e-5 if train_eps is None else train_eps
        sample_eps = 1e-3 if train_eps is None else sample_eps
    elif model_type != ModelType.VELOCITY:
        train_eps = 1e-3 if train_eps is None else train_eps
        sample_eps = 1e-3 if train_eps is None else sample_eps
    else:
        train_eps = sample_eps = 0
    return Transport(model_type=model_type, path_type=ptype, loss_type=loss_type, train_eps=train_eps,
                     sample_eps=sample_eps, snr_type=snr_type)


def _time_grid(t0, t1, num_steps, time_shifting_factor):
    """integrators.py:97-99: ``num_steps`` grid POINTS (num_steps-1 integration steps)."""
    t = th.linspace(t0, t1, num_steps)
    if time_shifting_factor:
        t = t / (t + time_shifting_factor - time_shifting_factor * t)
    return t


# kwargs the fused in-engine solve understands, and the ones it needs, per model mirror
_ENGINE_KW = {
    NextDiT: (("cap_feats", "cap_mask", "cfg_scale", "scale_factor", "scale_watershed", "base_seqlen", "proportional_attn"),
              ("cap_feats", "cap_mask", "cfg_scale")),
    CompositionalNextDiT: (("cap_feats", "cap_mask", "cfg_scale", "scale_factor", "scale_watershed", "base_seqlen", "proportional_attn",
                            "global_cap_feats", "global_cap_mask", "h_split_num", "w_split_num"),
                           ("cap_feats", "cap_mask", "cfg_scale", "global_cap_feats", "global_cap_mask")),
    DiT_Llama: (("y", "cfg_scale", "rope_scaling_factor", "ntk_factor"), ("y", "cfg_scale")),
    FlagDiT: (("cap_feats", "cap_mask", "cfg_scale", "rope_scaling_factor", "ntk_factor", "base_seqlen", "proportional_attn"),
              ("cap_feats", "cap_mask", "cfg_scale")),
}


def _engine_of(model_fn):
    owner = getattr(model_fn, "__self__", None)
    cls = type(owner)
    if cls in _ENGINE_KW and getattr(model_fn, "__func__", None) is cls.forward_with_cfg:
        return owner
    return None


def _fixed_grid_torch(fn: Callable, x: th.Tensor, t: th.Tensor, method: str) -> th.Tensor:
    """torchdiffeq fixed-grid euler / midpoint / rk4 semantics for an arbitrary model function
    (t is cast to the state dtype before the call, like torchdiffeq's _PerturbFunc)."""
    sol, y = [x], x

    def f(tt, yy):
        return fn(tt.to(yy.dtype), yy)

    for ta, tb in zip(t[:-1], t[1:]):
        dt = tb - ta
        if method == "euler":
            dy = dt * f(ta, y)
        elif method == "midpoint":
            half = 0.5 * dt
            dy = dt * f(ta + half, y + f(ta, y) * half)
        elif method == "rk4":     # torchdiffeq's fixed-grid rk4 = rk4_alt_step_func (3/8 rule), same expression order
            one_third, two_thirds = 1 / 3, 2 / 3
            k1 = f(ta, y)
            k2 = f(ta + dt * one_third, y + dt * k1 * one_third)
            k3 = f(ta + dt * two_thirds, y + dt * (k2 - k1 * one_third))
            k4 = f(tb, y + dt * (k1 - k2 + k3))
            dy = (k1 + 3 * (k2 + k3) + k4) * dt * 0.125
        else:
            raise NotImplementedError(
                f"sampling_method={method!r}: the B200 transport implements the fixed-grid euler/midpoint/rk4 solvers; "
                "adaptive solvers need torchdiffeq")
        y = y + dy
        sol.append(y)
    return th.stack(sol, dim=0)


def _solve(x, model_fn, t_grid, method, model_kwargs, wrap_drift=None):
    eng = _engine_of(model_fn) if wrap_drift is None else None
    # The in-engine solve keeps the state in bf16 and rounds t / dt the way torchdiffeq does for a bf16 state.  A state of
    # another dtype (fp32 latents with an fp32 checkpoint, the Next-DiT-MoE sample.py default) keeps the reference's
    # semantics through the generic loop below: fp32 state and time, the model call itself computes in bf16.
    if (eng is not None and method in ("euler", "midpoint", "rk4") and isinstance(x, th.Tensor) and x.is_cuda and x.dtype == th.bfloat16):
        allowed, required = _ENGINE_KW[type(eng)]
        if set(model_kwargs) <= set(allowed) and set(required) <= set(model_kwargs):
            return eng.sample_fixed_grid(x, t_grid.tolist(), method, **model_kwargs)
    device = x.device

    def _fn(t, xx):
        tv = th.ones(xx.size(0), device=device) * t
        if wrap_drift is not None:
            return wrap_drift(xx, tv, model_fn, **model_kwargs)
        return model_fn(xx, tv, **model_kwargs)

    return _fixed_grid_torch(_fn, x, t_grid.to(device), method)


class Sampler:
    """transport.py:221-391 (ODE part)."""

    def __init__(self, transport: Transport):
        self.transport = transport

    def sample_ode(self, *, sampling_method="dopri5", num_steps=50, atol=1e-6, rtol=1e-3, reverse=False,
                   time_shifting_factor=None):
        tr = self.transport
        t0, t1 = tr.check_interval(tr.train_eps, tr.sample_eps, sde=False, eval=True, reverse=reverse, last_step_size=0.0)
        if tr.model_type != ModelType.VELOCITY:
            raise NotImplementedError("the B200 transport implements velocity-prediction ODE sampling (the Lumina configuration); "
                                      "score / noise parameterisations are out of scope")
        # reverse=True: the reference evaluates the drift at 1 - t but also flips the interval to (1, 0), which its own ode class
        # rejects - same assertion here
        assert t0 < t1, "ODE sampler has to be in forward time"
        grid = _time_grid(t0, t1, num_steps, time_shifting_factor)
        wrap = (lambda xx, tv, fn, **kw: fn(xx, th.ones_like(tv) * (1 - tv), **kw)) if reverse else None

        def _sample(x, model, **model_kwargs):
            out = _solve(x, model, grid, sampling_method, model_kwargs, wrap_drift=wrap)
            assert out.shape[1:] == x.shape, "Output shape from ODE solver must match input shape"
            return out

        return _sample

    # ------------------------------------------------------------------ SDE sampling (velocity model, Linear path)
    @staticmethod
    def _expand(t, x):
        return t.view(t.size(0), *([1] * (x.dim() - 1)))

    def _diffusion(self, x, t, form, norm):
        """path.py ICPlan.compute_diffusion / compute_drift (alpha_t = t, sigma_t = 1 - t)."""
        t = self._expand(t, x)
        if form == "constant":
            return norm
        if form == "SBDM":
            sigma_t, d_sigma_t = 1 - t, -1
            return norm * ((1 / t) * (sigma_t ** 2) - sigma_t * d_sigma_t)
        if form in ("sigma", "linear"):
            return norm * (1 - t)
        if form == "decreasing":
            return 0.25 * (norm * th.cos(math.pi * t) + 1) ** 2
        if form == "inccreasing-decreasing":          # (sic) the reference's key
            return norm * th.sin(math.pi * t) ** 2
        raise NotImplementedError(f"Diffusion form {form} not implemented")

    @staticmethod
    def _score_terms(t):
        """The t-dependent factors of path.py ICPlan.get_score_from_velocity (alpha_t = t, sigma_t = 1 - t), same op sequence."""
        alpha_t, d_alpha_t, sigma_t, d_sigma_t = t, 1, 1 - t, -1
        reverse_alpha_ratio = alpha_t / d_alpha_t
        var = sigma_t ** 2 - reverse_alpha_ratio * d_sigma_t * sigma_t
        return reverse_alpha_ratio, var

    def _score(self, velocity, x, t):
        """path.py ICPlan.get_score_from_velocity."""
        reverse_alpha_ratio, var = self._score_terms(self._expand(t, x))
        return (reverse_alpha_ratio * velocity - x) / var

    def _sde_points(self, times, like, form, norm):
        """ndit_sde_point values for the in-engine SDE loop: the reference's own tensor ops on a one-element tensor of the state
        dtype, so that every scalar carries exactly the roundings the reference applies to its [B,1,1,1] tensors."""
        pts = []
        probe = th.zeros(1, 1, 1, 1, dtype=like.dtype, device=like.device)
        for tt in times:                                            # tt: one-element tensor of the state dtype (the `t` of the loop)
            te = self._expand(tt, probe)
            ratio, var = self._score_terms(te)
            diff = self._diffusion(probe, tt, form, norm)
            if not isinstance(diff, th.Tensor):
                return None                                         # "constant": th.sqrt(2 * float) fails in the reference as well
            pts.append((float(tt), float(ratio), float(var), float(diff), float(th.sqrt(2 * diff))))
        return pts

    def sample_sde(self, *, sampling_method="Euler", diffusion_form="SBDM", diffusion_norm=1.0, last_step="Mean",
                   last_step_size=0.04, num_steps=250):
        """transport.py:285-344 + integrators.py:5-76.  Returns ``fn(init, model, **kw) -> list of num_steps tensors``.
        The reference evaluates the model twice per drift (once for the drift, once for the score) on identical inputs;
        one evaluation is reused here."""
        tr = self.transport
        if tr.model_type != ModelType.VELOCITY or tr.path_type != PathType.LINEAR:
            raise NotImplementedError("the B200 transport implements SDE sampling for the velocity / Linear-path configuration")
        if sampling_method not in ("Euler", "Heun"):
            raise NotImplementedError("Smapler type not implemented.")
        if last_step not in (None, "Mean", "Tweedie", "Euler"):
            raise NotImplementedError()
        if last_step is None:
            last_step_size = 0.0
        t0, t1 = tr.check_interval(tr.train_eps, tr.sample_eps, diffusion_form=diffusion_form, sde=True, eval=True, reverse=False,
                                   last_step_size=last_step_size)
        assert t0 < t1, "SDE sampler has to be in forward time"
        grid = th.linspace(t0, t1, num_steps)
        dt = grid[1] - grid[0]

        def drift_and_score(x, t, model, kw):
            v = model(x, t, **kw)
            assert v.shape == x.shape, "Output shape from ODE solver must match input shape"
            return v, self._score(v, x, t)

        def sde_drift(x, t, model, kw):
            v, sc = drift_and_score(x, t, model, kw)
            return v + self._diffusion(x, t, diffusion_form, diffusion_norm) * sc

        def _fused_loop(init, model, kw):
            """The stochastic loop inside the engine (ndit_sample_sde) when `model` is the engine's forward_with_cfg and the state
            is bf16 on the GPU: noise drawn here exactly like the reference does (host RNG, one randn per step, same order), the
            t-dependent scalars computed with the reference's own ops.  Returns the list of states after each step, or None."""
            eng = _engine_of(model)
            if eng is None or not hasattr(eng, "sample_sde_loop") or not (isinstance(init, th.Tensor) and init.is_cuda and init.dtype == th.bfloat16):
                return None
            allowed, required = _ENGINE_KW[type(eng)]
            if not (set(kw) <= set(allowed) and set(required) <= set(kw)):
                return None
            # On the state's device, like the host loop: CUDA keeps a 0-dim CPU operand (ti, dt) in fp32 inside the op, the CPU
            # casts it to the tensor dtype first - `t + dt` differs by an ulp between the two, and the loop below must see what
            # the reference sees on this device.
            one = th.ones(1).to(init)
            times = []
            for ti in grid[:-1]:
                t = one * ti
                times.append(t)
                if sampling_method == "Heun":
                    times.append(t + dt)
            pts = self._sde_points(times, init, diffusion_form, diffusion_norm)
            if pts is None:
                return None
            noise = th.stack([th.randn(init.size()).to(init.dtype) for _ in grid[:-1]]).to(init.device)
            # dt, sqrt(dt), 0.5 * dt are 0-dim CPU tensors: a CUDA op takes them as fp32 scalars (no rounding to bf16)
            return eng.sample_sde_loop(init, pts, float(dt), float(th.sqrt(dt)), float(0.5 * dt), noise,
                                       0 if sampling_method == "Euler" else 1, **kw)

        def _sample(init, model, **kw):
            x, xs = init, []
            with th.no_grad():
                fused = _fused_loop(init, model, kw)
                for ti in (grid[:-1] if fused is None else []):
                    w_cur = th.randn(x.size()).to(x)
                    dw = w_cur * th.sqrt(dt)
                    t = th.ones(x.size(0)).to(x) * ti
                    if sampling_method == "Euler":      # Euler-Maruyama
                        drift = sde_drift(x, t, model, kw)
                        diffusion = self._diffusion(x, t, diffusion_form, diffusion_norm)
                        x = x + drift * dt + th.sqrt(2 * diffusion) * dw
                    else:                               # Heun
                        diffusion = self._diffusion(x, t, diffusion_form, diffusion_norm)
                        xhat = x + th.sqrt(2 * diffusion) * dw
                        k1 = sde_drift(xhat, t, model, kw)
                        xp = xhat + dt * k1
                        k2 = sde_drift(xp, t + dt, model, kw)
                        x = xhat + 0.5 * dt * (k1 + k2)
                    xs.append(x)
                if fused is not None:
                    xs = list(fused)
                ts = th.ones(init.size(0), device=init.device) * t1
                x = xs[-1]
                if last_step == "Mean":
                    x = x + sde_drift(x, ts, model, kw) * last_step_size
                elif last_step == "Euler":
                    x = x + model(x, ts, **kw) * last_step_size
                elif last_step == "Tweedie":
                    _, sc = drift_and_score(x, ts, model, kw)
                    x = x / ts[0] + ((1 - ts[0]) ** 2) / ts[0] * sc
                xs.append(x)
            assert len(xs) == num_steps, "Samples does not match the number of steps"
            return xs

        return _sample


class ODE:
    """lumina_next_t2i_mini/transport.py:57-111 (non-SD3 branch)."""

    def __init__(self, num_steps, sampler_type="euler", time_shifting_factor=None, t0=0.0, t1=1.0, use_sd3=False, strength=1.0):
        if use_sd3:
            raise NotImplementedError("SD3 sampling is out of scope")
        self.t = _time_grid(t0, t1, num_steps, time_shifting_factor)
        if strength != 1.0:
            self.t = self.t[int(num_steps * (1 - strength)):]
        self.sampler_type = sampler_type

    def sample(self, x, model, **model_kwargs):
        return _solve(x, model, self.t, self.sampler_type, model_kwargs)
