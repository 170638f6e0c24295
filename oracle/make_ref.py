"""Vendor the UNMODIFIED reference model / sampler sources of the hot path into ``oracle/_ref/``.

TEST INFRASTRUCTURE.  ``/root/reference`` exists only in the authoring container; the GPU box gets whatever is in
the repository snapshot.  ``oracle/_ref/`` is git-ignored (reference sources never enter this repository's history)
but NOT gpurun-ignored, so the byte-identical copies made here travel to the B200 box, where

  * ``tests/test_reference_gpu.py`` runs the real reference (fp32, and ``autocast(bf16)`` + ``flash_attn_varlen_func``)
    next to the engine at BASELINE config-2 / config-3 sizes,
  * ``bench.py`` times it as ``stock_cuda_baseline`` and, on the host cores, as the
    ``--impl reference`` arm / ``cpu_baseline`` (kind "reference").

Nothing under ``lumina_t2x_b200/`` imports these files.  The copy is verified by SHA-256 against the source and the
manifest (relative path -> digest) is written to ``oracle/_ref/MANIFEST.json``.

    python oracle/make_ref.py            # (re)create oracle/_ref from /root/reference
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")

# relative paths under /root/reference; every file is on (or imported by) the hot path of SURVEY.md section 8
FILES = [
    "LICENSE",
    # Lumina-Next-T2I, fairscale-free flavour (BASELINE configs 2 and 3)
    "lumina_next_t2i_mini/models/__init__.py",
    "lumina_next_t2i_mini/models/components.py",
    "lumina_next_t2i_mini/models/nextdit.py",
    "lumina_next_t2i_mini/transport.py",
    # Lumina-Next-T2I, canonical fairscale flavour + the full transport package (ODE and SDE samplers)
    "lumina_next_t2i/models/__init__.py",
    "lumina_next_t2i/models/components.py",
    "lumina_next_t2i/models/model.py",
    "lumina_next_t2i/transport/__init__.py",
    "lumina_next_t2i/transport/integrators.py",
    "lumina_next_t2i/transport/path.py",
    "lumina_next_t2i/transport/transport.py",
    "lumina_next_t2i/transport/utils.py",
    # compositional generation: region-masked caption cross-attention
    "lumina_next_compositional_generation/models/__init__.py",
    "lumina_next_compositional_generation/models/components.py",
    "lumina_next_compositional_generation/models/model.py",
    # class-conditional Next-DiT (config 1), MoE variants (config 5), Flag-DiT (config 4)
    "Next-DiT-ImageNet/models/models.py",
    "Next-DiT-MoE/models/models.py",
    "Next-DiT-MoE/models/models1.py",
    "Next-DiT-MoE/models/models2.py",
    "lumina_t2i/models/__init__.py",
    "lumina_t2i/models/components.py",
    "lumina_t2i/models/model.py",
]


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def make(force: bool = False) -> str:
    """Copy FILES from /root/reference to oracle/_ref (byte-identical).  No-op when /root/reference is absent
    (GPU box: the prebuilt copy is used)."""
    if not os.path.isdir(SRC):
        if not os.path.isdir(DST):
            raise RuntimeError("neither /root/reference nor oracle/_ref exists")
        return DST
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if force or not os.path.exists(d) or _sha(s) != _sha(d):
            shutil.copyfile(s, d)
        assert _sha(s) == _sha(d), rel
        manifest[rel] = _sha(d)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "Alpha-VLLM/Lumina-T2X (unmodified files, see LICENSE)", "sha256": manifest}, f, indent=1, sort_keys=True)
    return DST


def verify() -> bool:
    """True when oracle/_ref matches its manifest (used on the GPU box, where the source tree is absent)."""
    p = os.path.join(DST, "MANIFEST.json")
    if not os.path.exists(p):
        return False
    man = json.load(open(p))["sha256"]
    return all(os.path.exists(os.path.join(DST, rel)) and _sha(os.path.join(DST, rel)) == dig for rel, dig in man.items())


if __name__ == "__main__":
    print(make(force="--force" in sys.argv), "verified" if verify() else "MISMATCH")
