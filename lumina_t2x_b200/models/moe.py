"""Mirror of ``Next-DiT-MoE/models/__init__.py``: the class-conditional Next-DiT with a mixture-of-experts FFN
(``import lumina_t2x_b200.models.moe as models`` in ``Next-DiT-MoE/sample.py:17,87``).  Same state-dict keys as
``models.py`` (time-gated, 8 experts), ``models1.py`` (token-gated, 8 experts) and ``models2.py`` (both, 4 + 4)."""
from .dit_llama import DiT_Llama


def DiT_Llama_600M_patch2(**kwargs):
    """Next-DiT-MoE/models/models.py:1015-1018 (time-gated MoE)."""
    return DiT_Llama(patch_size=2, dim=1536, n_layers=16, n_heads=32, moe="time", **kwargs)


def DiT_Llama_2B_patch2(**kwargs):
    """Next-DiT-MoE/models/models.py:1027-1030 (time-gated MoE)."""
    return DiT_Llama(patch_size=2, dim=2304, n_layers=24, n_heads=32, moe="time", **kwargs)


def DiT_Llama_600M_patch2_Spatial(**kwargs):
    """Next-DiT-MoE/models/models1.py:1015-1018 (token-gated MoE)."""
    return DiT_Llama(patch_size=2, dim=1536, n_layers=16, n_heads=32, moe="space", **kwargs)


def DiT_Llama_600M_patch2_Both(**kwargs):
    """Next-DiT-MoE/models/models2.py:1063-1066 (time-gated then token-gated MoE)."""
    return DiT_Llama(patch_size=2, dim=1536, n_layers=16, n_heads=32, moe="both", **kwargs)


def DiT_Llama_3B_patch2(**kwargs):
    """Next-DiT-MoE/models/models.py:1033-1036 (time-gated MoE, head_dim 96)."""
    return DiT_Llama(patch_size=2, dim=3072, n_layers=32, n_heads=32, moe="time", **kwargs)


def DiT_Llama_7B_patch2(**kwargs):
    """Next-DiT-MoE/models/models.py:1039-1042 (head_dim 128: constructs, but is outside the attention kernel's head dims)."""
    return DiT_Llama(patch_size=2, dim=4096, n_layers=32, n_heads=32, moe="time", **kwargs)


def DiT_Llama_600M_GQA_patch2(**kwargs):
    """Next-DiT-MoE/models/models.py:1021-1024 (time-gated MoE, 8 kv heads; defined in all three reference files, not
    re-exported by the reference package)."""
    return DiT_Llama(patch_size=2, dim=1536, n_layers=16, n_heads=32, n_kv_heads=8, moe="time", **kwargs)
