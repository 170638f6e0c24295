"""Import the UNMODIFIED reference modules from /root/reference (authoring container only).

TEST INFRASTRUCTURE.  Used by ``oracle/make_golden.py`` to pin the oracle against the
reference's own outputs.  /root/reference does not exist on the GPU box, so nothing
that runs there imports this file.

Shims that live here (not in the reference):
  * ``torchdiffeq.odeint`` stub: torchdiffeq is an unpinned third-party dependency
    (requirements.txt:7) that is absent from this image.  The stub restates its
    fixed-grid ``euler`` / ``midpoint`` solvers (FixedGridODESolver.integrate: for
    consecutive grid points t0,t1: dt=t1-t0; euler dy=dt*f(t0,y0); midpoint
    dy=dt*f(t0+dt/2, y0+f(t0,y0)*dt/2); outputs at the grid points) including
    ``_PerturbFunc``'s cast of ``t`` to the state dtype before the user function is called.
  * ``torch.Tensor.cuda`` -> identity when no GPU is present, because
    ``NextDiT.precompute_freqs_cis`` hard-codes ``.cuda()`` (models/nextdit.py:919).
"""
from __future__ import annotations

import importlib
import sys
import types

import torch

REF_ROOT = "/root/reference"


def _odeint(func, y0, t, *, method="euler", atol=None, rtol=None, **kw):
    assert method in ("euler", "midpoint"), method
    sol = [y0]
    y = y0

    def f(tt, yy):
        return func(tt.to(yy.abs().dtype), yy)

    for t0, t1 in zip(t[:-1], t[1:]):
        dt = t1 - t0
        if method == "euler":
            dy = dt * f(t0, y)
        else:
            half_dt = 0.5 * dt
            y_mid = y + f(t0, y) * half_dt
            dy = dt * f(t0 + half_dt, y_mid)
        y = y + dy
        sol.append(y)
    return torch.stack(sol, dim=0)


def install_shims() -> None:
    if "torchdiffeq" not in sys.modules:
        try:
            importlib.import_module("torchdiffeq")
        except ImportError:
            m = types.ModuleType("torchdiffeq")
            m.odeint = _odeint
            sys.modules["torchdiffeq"] = m
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # type: ignore[assignment]


def import_reference_mini():
    """Returns (models_module, transport_module) of lumina_next_t2i_mini, unmodified."""
    install_shims()
    root = REF_ROOT + "/lumina_next_t2i_mini"
    if root not in sys.path:
        sys.path.insert(0, root)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        models = importlib.import_module("models")
        transport = importlib.import_module("transport")
    return models, transport
