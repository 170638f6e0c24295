"""Reference-side mirror of the caption encoder: ``transformers.AutoModel.from_pretrained("google/gemma-2b")`` as
lumina_next_t2i/sample.py uses it (sample.py:46-50, :111),

    prompt_embeds = text_encoder(input_ids=ids.cuda(), attention_mask=mask.cuda(), output_hidden_states=True).hidden_states[-2]

behind the C ABI of include/ndit_text.h (ntxt_*).  Same constructor surface as a ``GemmaModel`` built from a ``GemmaConfig``-like
object, same state-dict keys (``embed_tokens.weight``, ``layers.<i>.self_attn.{q,k,v,o}_proj.weight``,
``layers.<i>.mlp.{gate,up,down}_proj.weight``, ``layers.<i>.{input,post_attention}_layernorm.weight``, ``norm.weight``), so
``load_state_dict(hf_model.state_dict(), strict=True)`` works.  Only what the sampling path consumes is computed:
``hidden_states[-2]`` (the output of the second-to-last decoder layer); the other entries of the returned tuple are ``None``.
No PyTorch fallback: without the CUDA library / an sm_100 device the call raises."""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import _lib


class _Linear(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(o, i))
        nn.init.normal_(self.weight, std=i ** -0.5)


class _Norm(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(d))


class _Attn(nn.Module):
    def __init__(self, D, H, Hkv, hd):
        super().__init__()
        self.q_proj, self.k_proj, self.v_proj, self.o_proj = _Linear(D, H * hd), _Linear(D, Hkv * hd), _Linear(D, Hkv * hd), _Linear(H * hd, D)


class _MLP(nn.Module):
    def __init__(self, D, F):
        super().__init__()
        self.gate_proj, self.up_proj, self.down_proj = _Linear(D, F), _Linear(D, F), _Linear(F, D)


class _Layer(nn.Module):
    def __init__(self, D, H, Hkv, hd, F):
        super().__init__()
        self.self_attn, self.mlp = _Attn(D, H, Hkv, hd), _MLP(D, F)
        self.input_layernorm, self.post_attention_layernorm = _Norm(D), _Norm(D)


class GemmaTextEncoder(nn.Module):
    """``GemmaModel`` restricted to what sample.py reads from it.  ``config``: anything with GemmaConfig's attribute names."""

    def __init__(self, config=None, max_tokens: int = 1024, **kw):
        super().__init__()
        get = lambda n, d: kw.get(n, getattr(config, n, d) if config is not None else d)   # noqa: E731
        self.vocab_size, self.hidden_size = int(get("vocab_size", 256000)), int(get("hidden_size", 2048))
        self.num_hidden_layers = int(get("num_hidden_layers", 18))
        self.num_attention_heads, self.num_key_value_heads = int(get("num_attention_heads", 8)), int(get("num_key_value_heads", 1))
        self.head_dim, self.intermediate_size = int(get("head_dim", 256)), int(get("intermediate_size", 16384))
        self.rms_norm_eps = float(get("rms_norm_eps", 1e-6))
        rp = get("rope_parameters", None)
        self.rope_theta = float(rp["rope_theta"]) if isinstance(rp, dict) and "rope_theta" in rp else float(get("rope_theta", 10000.0))
        act = get("hidden_act", "gelu_pytorch_tanh")
        if act not in ("gelu_pytorch_tanh", "gelu_tanh"):
            raise NotImplementedError(f"hidden_act={act!r}: the engine implements Gemma's gelu_pytorch_tanh gating")
        self.config = SimpleNamespace(hidden_size=self.hidden_size, num_hidden_layers=self.num_hidden_layers, vocab_size=self.vocab_size)
        self.embed_tokens = nn.Embedding(self.vocab_size, self.hidden_size)
        self.layers = nn.ModuleList([_Layer(self.hidden_size, self.num_attention_heads, self.num_key_value_heads, self.head_dim,
                                            self.intermediate_size) for _ in range(self.num_hidden_layers)])
        self.norm = _Norm(self.hidden_size)
        self._max_tokens, self._handle, self._dirty = int(max_tokens), None, True

    # ------------------------------------------------------------------ engine plumbing
    def load_state_dict(self, *a, **k):
        self._dirty = True
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._dirty = True
        return super()._apply(fn, *a, **k)

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().ntxt_destroy(self._handle)
        except Exception:
            pass

    def _engine(self, device):
        lib = _lib.load()
        if self._handle is not None and not self._dirty:
            return lib, self._handle
        if self._handle is not None:
            lib.ntxt_destroy(self._handle)
            self._handle = None
        cfg = _lib.NtxtConfig(self.vocab_size, self.hidden_size, self.num_hidden_layers, self.num_attention_heads, self.num_key_value_heads,
                              self.head_dim, self.intermediate_size, self.rms_norm_eps, self.rope_theta, self._max_tokens)
        h = C.c_void_p()
        with torch.cuda.device(device):
            rc = lib.ntxt_create(C.byref(cfg), C.byref(h))
            if rc != 0:
                raise RuntimeError(f"ntxt_create failed ({rc}): {(lib.ntxt_last_error(None) or b'?').decode()}")
            stream = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
            for key, p in self.state_dict().items():
                t = p.detach()
                if t.device != device:
                    raise RuntimeError(f"parameter {key} is on {t.device}, expected {device}")
                if t.dtype == torch.bfloat16:
                    dt = _lib.NDIT_BF16
                else:
                    t, dt = t.float(), _lib.NDIT_F32
                t = t.contiguous()
                shape = (C.c_int64 * t.dim())(*t.shape)
                self._check(lib, h, lib.ntxt_set_weight(h, key.encode(), C.c_void_p(t.data_ptr()), shape, t.dim(), dt, stream))
                del t
            torch.cuda.current_stream(device).synchronize()
            self._check(lib, h, lib.ntxt_finalize_weights(h, stream))
        self._handle, self._dirty = h, False
        return lib, h

    @staticmethod
    def _check(lib, h, rc):
        if rc != 0:
            raise RuntimeError(f"ntxt error {rc}: {(lib.ntxt_last_error(h) or b'?').decode()}")

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, output_hidden_states: bool = True, **unused):
        if input_ids is None or not input_ids.is_cuda:
            raise RuntimeError("GemmaTextEncoder (B200 engine) needs CUDA input_ids; there is no CPU path")
        if not output_hidden_states:
            raise NotImplementedError("the engine computes hidden_states[-2] (what sample.py reads); call with output_hidden_states=True")
        dev = input_ids.device
        lib, h = self._engine(dev)
        ids = input_ids.detach().to(torch.int64).contiguous()
        B, T = ids.shape
        if B * T > self._max_tokens:
            raise ValueError(f"batch * sequence length {B * T} > max_tokens {self._max_tokens}")
        mask = None if attention_mask is None else attention_mask.detach().to(device=dev, dtype=torch.int64).contiguous()
        out = torch.empty(B, T, self.hidden_size, dtype=torch.bfloat16, device=dev)
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            self._check(lib, h, lib.ntxt_encode(h, C.c_void_p(ids.data_ptr()), C.c_void_p(mask.data_ptr()) if mask is not None else None, B, T,
                                                C.c_void_p(out.data_ptr()), stream))
        pd = next(self.parameters()).dtype
        hs = [None] * (self.num_hidden_layers + 1)
        hs[-2] = out.to(pd) if pd in (torch.float32, torch.float16) else out
        return SimpleNamespace(hidden_states=tuple(hs), last_hidden_state=None)
