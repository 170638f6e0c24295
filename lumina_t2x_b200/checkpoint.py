"""Checkpoint tooling for the B200 engine.

Mirrors ``lumina_next_t2i/entry_point.py:115-156`` (``lumina_next convert``: ``.pth`` <-> ``.safetensors``) and adds the
engine's own format:

  convert(weight_path, output_dir)      same behaviour as the reference command (.pth -> .safetensors and back)
  load_state_dict_file(path)            .pth / .safetensors -> {reference key: tensor}; the .safetensors reader is a
                                         zero-copy mmap parser written here (8-byte header length, JSON header, raw tensors),
                                         so nothing is read twice on a cold start
  save_packed(model, path)              the finalized engine weights in their GEMM-ready layout (ndit_save_packed)
  load_packed(model, path)              straight into a fresh engine (ndit_load_packed): no state dict, no re-packing kernels

The packed file is tied to the architecture and the ABI version; the module's own ``nn.Parameter`` tensors are NOT filled by
``load_packed`` (they are only needed for ``state_dict()`` / fine-tuning, not for sampling).
"""
from __future__ import annotations

import ctypes as C
import json
import mmap
import os
import struct
from typing import Dict

import torch

from . import _lib

_ST_DTYPES = {"BF16": torch.bfloat16, "F16": torch.float16, "F32": torch.float32, "F64": torch.float64, "I64": torch.int64,
              "I32": torch.int32, "I16": torch.int16, "I8": torch.int8, "U8": torch.uint8, "BOOL": torch.bool}
_ST_NAMES = {v: k for k, v in _ST_DTYPES.items()}


def read_safetensors(path: str) -> Dict[str, torch.Tensor]:
    """Zero-copy reader of the safetensors container: tensors are views of one read-only memory map."""
    f = open(path, "rb")
    mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    (hlen,) = struct.unpack("<Q", mm[:8])
    header = json.loads(mm[8:8 + hlen].decode("utf-8"))
    base = 8 + hlen
    buf = memoryview(mm)
    out = {}
    for key, meta in header.items():
        if key == "__metadata__":
            continue
        dt = _ST_DTYPES[meta["dtype"]]
        b0, b1 = meta["data_offsets"]
        shape = tuple(meta["shape"])
        n = 1
        for d in shape:
            n *= d
        if n == 0:
            out[key] = torch.empty(shape, dtype=dt)
            continue
        t = torch.frombuffer(buf, dtype=dt, count=n, offset=base + b0)
        assert (b1 - b0) == n * t.element_size(), key
        out[key] = t.view(shape)
    return out


def write_safetensors(tensors: Dict[str, torch.Tensor], path: str) -> None:
    header, off = {}, 0
    items = []
    for k in sorted(tensors):
        t = tensors[k].detach().cpu().contiguous()
        nb = t.numel() * t.element_size()
        header[k] = {"dtype": _ST_NAMES[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + nb]}
        off += nb
        items.append(t)
    hb = json.dumps(header, separators=(",", ":")).encode("utf-8")
    hb += b" " * ((8 - len(hb) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hb)))
        f.write(hb)
        for t in items:
            f.write(t.view(torch.uint8).numpy().tobytes() if t.numel() else b"")


def load_state_dict_file(path: str) -> Dict[str, torch.Tensor]:
    ext = os.path.splitext(path)[1]
    if ext == ".safetensors":
        return read_safetensors(path)
    if ext == ".pth":
        return torch.load(path, map_location="cpu", weights_only=True)
    raise ValueError("Only ('.pth', '.safetensors') models are supported.")


def convert(weight_path: str, output_dir: str) -> str:
    """``lumina_next convert`` (entry_point.py:115-156): .pth -> .safetensors, .safetensors -> .pth, same file stem."""
    file_path, ext = os.path.splitext(weight_path)
    if ext not in (".pth", ".safetensors"):
        raise ValueError("Only ('.pth', '.safetensors') models are supported for conversion.")
    os.makedirs(output_dir, exist_ok=True)
    name = os.path.basename(file_path)
    sd = load_state_dict_file(weight_path)
    if ext == ".pth":
        out = os.path.join(output_dir, name + ".safetensors")
        write_safetensors(sd, out)
    else:
        out = os.path.join(output_dir, name + ".pth")
        torch.save({k: v.clone() for k, v in sd.items()}, out)
    return out


def save_packed(model, path: str) -> None:
    """Write the engine's finalized, GEMM-ready weights (needs the model on a CUDA device)."""
    lib, h = model._engine(next(model.parameters()).device)
    _lib.check(lib.ndit_save_packed(h, os.fsencode(path)), h)


def load_packed(model, path: str, device="cuda") -> None:
    """Create the model's engine directly from a packed file: no state dict, no per-tensor re-packing launches."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("load_packed needs a CUDA device; there is no CPU path")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    lib = _lib.load()
    model._check_supported()
    model._destroy()
    cfg = model._ndit_config()
    h = C.c_void_p()
    with torch.cuda.device(device):
        _lib.check(lib.ndit_create(C.byref(cfg), C.byref(h)), None)
        rc = lib.ndit_load_packed(h, os.fsencode(path))
        if rc != 0:
            msg = lib.ndit_last_error(h)
            lib.ndit_destroy(h)
            raise RuntimeError(f"ndit error {rc}: {msg.decode() if msg else '?'}")
    model._handle, model._dirty, model._cap_key = h, False, None
    if hasattr(model, "_label_key"):
        model._label_key = None
    model._packed_path = path
