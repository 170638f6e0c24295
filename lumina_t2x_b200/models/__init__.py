"""Mirror of the reference ``models`` package (lumina_next_t2i/models/__init__.py:1)."""
from .nextdit import NextDiT, NextDiT_2B_GQA_patch2, NextDiT_2B_patch2  # noqa: F401
