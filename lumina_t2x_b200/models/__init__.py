"""Mirror of the reference ``models`` packages: ``lumina_next_t2i/models/__init__.py:1`` (NextDiT) and
``Next-DiT-ImageNet/models/__init__.py`` (DiT_Llama factories)."""
from .dit_llama import DiT_Llama, DiT_Llama_2B_patch2, DiT_Llama_3B_patch2, DiT_Llama_7B_patch2, DiT_Llama_600M_patch2  # noqa: F401
from .nextdit import NextDiT, NextDiT_2B_GQA_patch2, NextDiT_2B_patch2  # noqa: F401
