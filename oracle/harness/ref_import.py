"""Import the UNMODIFIED reference modules: from /root/reference in the authoring container, from the
byte-identical copies under ``oracle/_ref/`` (made by ``oracle/make_ref.py``, git-ignored, shipped with the
gpurun snapshot) on the GPU box.

TEST INFRASTRUCTURE.  Used by ``oracle/make_golden.py`` to pin the oracle against the reference's own outputs,
by ``tests/test_reference_gpu.py`` to run the real reference next to the engine on the B200, and by ``bench.py``
for the ``stock_cuda_baseline`` / ``--impl reference`` legs.  Never imported by ``lumina_t2x_b200/``.

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

import os

_VENDORED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_ref")
REF_ROOT = "/root/reference" if os.path.isdir("/root/reference") else _VENDORED


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "lumina_next_t2i_mini", "models", "nextdit.py"))


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


def _install_fairscale_stub() -> None:
    """fairscale (unpinned third-party dependency, requirements.txt:2) at model-parallel world size 1:
    Column/RowParallelLinear(in, out, bias, init_method=...) == nn.Linear with init_method(weight) and a zero bias,
    ParallelEmbedding == nn.Embedding, fs_init.get_model_parallel_world_size() == 1."""
    if "fairscale" in sys.modules:
        return
    import torch.nn as nn

    class _PLinear(nn.Linear):
        def __init__(self, in_features, out_features, bias=True, gather_output=True, input_is_parallel=False, init_method=None, **kw):
            super().__init__(in_features, out_features, bias=bias)
            if init_method is not None:
                init_method(self.weight)
            if bias:
                nn.init.zeros_(self.bias)

    class _PEmbedding(nn.Embedding):
        def __init__(self, num_embeddings, embedding_dim, init_method=None, **kw):
            super().__init__(num_embeddings, embedding_dim)
            if init_method is not None:
                init_method(self.weight)

    fs = types.ModuleType("fairscale")
    nnm = types.ModuleType("fairscale.nn")
    mp = types.ModuleType("fairscale.nn.model_parallel")
    init = types.ModuleType("fairscale.nn.model_parallel.initialize")
    layers = types.ModuleType("fairscale.nn.model_parallel.layers")
    init.get_model_parallel_world_size = lambda: 1
    init.get_model_parallel_rank = lambda: 0
    init.get_model_parallel_src_rank = lambda: 0
    init.get_model_parallel_group = lambda: None
    layers.ColumnParallelLinear = _PLinear
    layers.RowParallelLinear = _PLinear
    layers.ParallelEmbedding = _PEmbedding
    mp.initialize, mp.layers = init, layers
    nnm.model_parallel = mp
    fs.nn = nnm
    for name, mod in (("fairscale", fs), ("fairscale.nn", nnm), ("fairscale.nn.model_parallel", mp),
                      ("fairscale.nn.model_parallel.initialize", init), ("fairscale.nn.model_parallel.layers", layers)):
        sys.modules[name] = mod


def import_reference_imagenet():
    """Returns the unmodified Next-DiT-ImageNet ``models.models`` module (class-conditional DiT_Llama)."""
    import importlib.util
    import os
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")      # @torch.compile on the SiLU gate cannot build on this CPU box
    install_shims()
    _install_fairscale_stub()
    path = REF_ROOT + "/Next-DiT-ImageNet/models/models.py"
    spec = importlib.util.spec_from_file_location("ref_imagenet_models", path)
    mod = importlib.util.module_from_spec(spec)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    return mod


def import_reference_moe(which: str):
    """Unmodified Next-DiT-MoE model module: which = "models" (time MoE), "models1" (space MoE), "models2" (both)."""
    import importlib.util
    import os
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
    install_shims()
    _install_fairscale_stub()
    spec = importlib.util.spec_from_file_location("ref_moe_" + which, f"{REF_ROOT}/Next-DiT-MoE/models/{which}.py")
    mod = importlib.util.module_from_spec(spec)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    return mod


def import_reference_flag_dit():
    """Returns the unmodified lumina_t2i ``models.model`` module (Flag-DiT ``DiT_Llama`` / ``DiT_Llama_5B_patch2``)."""
    import importlib.util
    install_shims()
    _install_fairscale_stub()
    pkg_dir = REF_ROOT + "/lumina_t2i/models"
    # the file does ``from .components import RMSNorm``: load it as a submodule of a synthetic package
    pkg = types.ModuleType("ref_lumina_t2i_models")
    pkg.__path__ = [pkg_dir]
    sys.modules["ref_lumina_t2i_models"] = pkg
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = importlib.import_module("ref_lumina_t2i_models.model")
    return mod


def import_reference_full_model():
    """Unmodified canonical ``lumina_next_t2i/models/model.py`` (the fairscale flavour sample.py / demo.py import),
    with fairscale replaced by the world-size-1 stub above."""
    install_shims()
    _install_fairscale_stub()
    pkg = types.ModuleType("ref_lumina_next_t2i_models")
    pkg.__path__ = [REF_ROOT + "/lumina_next_t2i/models"]
    sys.modules["ref_lumina_next_t2i_models"] = pkg
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = importlib.import_module("ref_lumina_next_t2i_models.model")
    return mod


def import_reference_compositional_model():
    """Unmodified ``lumina_next_compositional_generation/models/model.py`` (region-masked cross-attention), fairscale replaced by the
    world-size-1 stub above."""
    install_shims()
    _install_fairscale_stub()
    pkg = types.ModuleType("ref_lumina_next_compositional_models")
    pkg.__path__ = [REF_ROOT + "/lumina_next_compositional_generation/models"]
    sys.modules["ref_lumina_next_compositional_models"] = pkg
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = importlib.import_module("ref_lumina_next_compositional_models.model")
    return mod


def import_reference_full_transport():
    """Unmodified ``lumina_next_t2i/transport`` package (create_transport, Sampler with sample_ode AND sample_sde)."""
    import importlib.util
    install_shims()
    d = REF_ROOT + "/lumina_next_t2i/transport"
    spec = importlib.util.spec_from_file_location("ref_full_transport", d + "/__init__.py", submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_full_transport"] = mod
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    return mod


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
