"""ctypes binding of libndit_b200.so (C ABI: include/ndit.h).  Fails loudly: there is no
CPU or PyTorch fallback for the product path."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libndit_b200.so")

NDIT_BF16, NDIT_F32 = 0, 1
NDIT_EULER, NDIT_MIDPOINT, NDIT_RK4 = 0, 1, 2


class NditConfig(C.Structure):
    _fields_ = [("dim", C.c_int32), ("n_layers", C.c_int32), ("n_heads", C.c_int32), ("n_kv_heads", C.c_int32),
                ("cap_feat_dim", C.c_int32), ("in_channels", C.c_int32), ("patch_size", C.c_int32),
                ("multiple_of", C.c_int32), ("learn_sigma", C.c_int32), ("norm_eps", C.c_float),
                ("max_tokens", C.c_int32), ("max_cap_len", C.c_int32), ("max_batch", C.c_int32),
                ("num_classes", C.c_int32), ("flag_dit", C.c_int32),
                ("moe_time_experts", C.c_int32), ("moe_space_experts", C.c_int32),
                ("ffn_dim", C.c_int32), ("no_qk_norm", C.c_int32)]


class NditStepParams(C.Structure):
    _fields_ = [("cfg_scale", C.c_float), ("scale_factor", C.c_float), ("scale_watershed", C.c_float),
                ("proportional_attn", C.c_int32), ("base_seqlen", C.c_int32), ("ntk_factor", C.c_float)]


class NditSdePoint(C.Structure):
    _fields_ = [("t", C.c_float), ("ratio", C.c_float), ("var", C.c_float), ("diffusion", C.c_float), ("sqrt_2diffusion", C.c_float)]


_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); every symbol declared in include/ndit.h
SIGNATURES = {
    "ndit_abi_version": (C.c_int, []),
    "ndit_create": (C.c_int, [C.POINTER(NditConfig), C.POINTER(_vp)]),
    "ndit_destroy": (C.c_int, [_vp]),
    "ndit_last_error": (C.c_char_p, [_vp]),
    "ndit_set_weight": (C.c_int, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i32, _i32, _vp]),
    "ndit_finalize_weights": (C.c_int, [_vp, _vp]),
    "ndit_parameter_count": (_i64, [_vp]),
    "ndit_reserve": (C.c_int, [_vp, _i32, _i32, _i32]),
    "ndit_save_packed": (C.c_int, [_vp, C.c_char_p]),
    "ndit_load_packed": (C.c_int, [_vp, C.c_char_p]),
    "ndit_set_caption": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp]),
    "ndit_set_caption_regions": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _vp]),
    "ndit_set_labels": (C.c_int, [_vp, _vp, _i32, _vp]),
    "ndit_forward_cfg": (C.c_int, [_vp, _vp, _f32, _i32, _i32, _i32, C.POINTER(NditStepParams), _vp, _vp]),
    "ndit_forward": (C.c_int, [_vp, _vp, C.POINTER(_f32), _i32, _i32, _i32, C.POINTER(NditStepParams), _vp, _vp]),
    "ndit_debug_read_residual": (C.c_int, [_vp, _vp, _i64, _vp]),
    "ndit_op_moe_gate": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "ndit_forward_list": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_f32), _i32,
                                    C.POINTER(NditStepParams), C.POINTER(_vp), _vp]),
    "ndit_sample": (C.c_int, [_vp, _vp, _i32, _i32, _i32, C.POINTER(_f32), _i32, _i32, C.POINTER(NditStepParams),
                              _vp, _vp, _vp]),
    "ndit_sample_sde": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, C.POINTER(NditSdePoint), _f32, _f32, _f32, _vp,
                                  C.POINTER(NditStepParams), _vp, _vp]),
    "ndit_sample_host": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, C.POINTER(_f32), _i32, _i32,
                                   C.POINTER(NditStepParams), _vp, _vp]),
    "ndit_launch_count": (_i64, [_vp]),
    "ndit_graph_replay_count": (_i64, [_vp]),
    "ndit_set_option": (C.c_int, [_vp, C.c_char_p, _i32]),
    "ndit_profile_read": (C.c_int, [_vp, C.POINTER(_f32), C.POINTER(_i64), _i32]),
    "ndit_op_gemm": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "ndit_op_gemm_bench": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, C.POINTER(_f32), _vp]),
    "ndit_op_ln_rope": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _i32, _vp]),
    "ndit_op_attention": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _i32, _vp]),
    "ndit_op_attention_hd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _i32, _vp]),
    "ndit_op_attention_bench": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _i32,
                                          C.POINTER(_f32), _vp]),
    "ndit_op_resid_rms_mod": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _vp]),
}

class NtxtConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("vocab_size", "hidden_size", "num_hidden_layers", "num_attention_heads", "num_key_value_heads",
                                         "head_dim", "intermediate_size")] + [("rms_norm_eps", C.c_float), ("rope_theta", C.c_float),
                                                                              ("max_tokens", C.c_int32)]


# include/ndit_text.h (caption-encoder end)
TEXT_SIGNATURES = {
    "ntxt_create": (C.c_int, [C.POINTER(NtxtConfig), C.POINTER(_vp)]),
    "ntxt_destroy": (C.c_int, [_vp]),
    "ntxt_last_error": (C.c_char_p, [_vp]),
    "ntxt_set_weight": (C.c_int, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i32, _i32, _vp]),
    "ntxt_finalize_weights": (C.c_int, [_vp, _vp]),
    "ntxt_encode": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp]),
}


class NvaeConfig(C.Structure):
    _fields_ = [("latent_channels", C.c_int32), ("out_channels", C.c_int32), ("block_out_channels", C.c_int32 * 4),
                ("layers_per_block", C.c_int32), ("norm_num_groups", C.c_int32)]


# include/ndit_vae.h (VAE-decode end)
VAE_SIGNATURES = {
    "nvae_create": (C.c_int, [C.POINTER(NvaeConfig), C.POINTER(_vp)]),
    "nvae_destroy": (C.c_int, [_vp]),
    "nvae_last_error": (C.c_char_p, [_vp]),
    "nvae_set_weight": (C.c_int, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i32, _i32, _vp]),
    "nvae_finalize_weights": (C.c_int, [_vp, _vp]),
    "nvae_decode": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library.  Raises if it has not been built (``__graft_entry__.build()`` or
    ``python lumina_t2x_b200/build.py``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the B200 engine has no fallback path. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in list(SIGNATURES.items()) + list(TEXT_SIGNATURES.items()) + list(VAE_SIGNATURES.items()):
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.ndit_abi_version() != 5:
        raise RuntimeError("libndit_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, handle=None) -> None:
    if rc != 0:
        msg = load().ndit_last_error(handle)
        raise RuntimeError(f"ndit error {rc}: {msg.decode() if msg else '?'}")
