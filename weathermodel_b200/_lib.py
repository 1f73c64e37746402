"""ctypes binding of libwm_b200.so (the C ABI declared in include/wm_b200.h).

The product path has NO fallback: if the shared library is missing this module raises at import of
`lib()`; if a call returns non-zero, `check()` raises RuntimeError with the library's own message.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WM_B200_LIB") or os.path.join(_HERE, "libwm_b200.so")  # override: diagnostic builds (tools/)

_vp, _i, _i64, _u64, _f, _sz, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_size_t, C.c_double


class GemmEpilogue(C.Structure):
    _fields_ = [
        ("bias", _vp), ("relu", _i), ("dropout_p", _f), ("seed", _u64), ("stream_id", _u64),
        ("gate_bf16", _vp), ("ld_gate", _i), ("gate_scale", _f), ("residual_bf16", _vp), ("ld_res", _i),
        ("sign_bits_out", _vp), ("gate_bits", _vp),
    ]


class EncoderConfig(C.Structure):
    _fields_ = [
        ("B", _i), ("S", _i), ("F", _i), ("D", _i), ("H", _i), ("L", _i), ("FF", _i), ("out_dim", _i),
        ("dropout_p", _f), ("ln_eps", _f), ("eval_only", _i),
    ]

    def __init__(self, B=0, S=0, F=0, D=0, H=0, L=0, FF=0, out_dim=0, dropout_p=0.0, ln_eps=1e-5, eval_only=0):
        super().__init__(B, S, F, D, H, L, FF, out_dim, dropout_p, ln_eps, eval_only)


# name -> (restype, argtypes); kept in the order of include/wm_b200.h
SIGNATURES = {
    "wm_abi_version": (_i, []),
    "wm_strerror": (C.c_char_p, [_i]),
    "wm_device_error": (_i, []),
    "wm_launch_count": (C.c_longlong, []),
    "wm_debug_ticks": (_i, [C.POINTER(C.c_longlong), _i]),
    "wm_set_option": (_i, [C.c_char_p, _i]),
    "wm_rand_grid_x": (_i, [_i64]),
    "wm_mask_bert": (_i, [_u64, _u64, _i, _f, _i64, _vp, _vp, _vp]),
    "wm_mask_former": (_i, [_u64, _u64, _i, _i, _i64, _i, _vp, _vp]),
    "wm_embed_fwd": (_i, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "wm_gemm_tn": (_i, [_vp, _i, _vp, _i, _i, _i, _i, C.POINTER(GemmEpilogue), _vp, _i, _i, _i, _vp]),
    "wm_gemm_set_variant": (_i, [_i, _i, _i, C.POINTER(GemmEpilogue), _i, _i, _i, _i]),
    "wm_gemm_sign_bits_bytes": (_sz, [_i, _i]),
    "wm_gemm_wgrad_workspace_bytes": (_sz, [_i, _i, _i]),
    "wm_gemm_wgrad": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "wm_umma_probe": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "wm_attn_dropout_words_bytes": (_sz, [_i, _i, _i]),
    "wm_attn_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "wm_attn_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _u64, _u64, _vp]),
    "wm_attn_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "wm_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _vp]),
    "wm_layernorm_bwd_workspace_bytes": (_sz, [_i, _i]),
    "wm_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _f, _u64, _u64, _vp, _vp]),
    "wm_colsum_workspace_bytes": (_sz, [_i, _i]),
    "wm_colsum": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "wm_loss_bert": (_i, [_vp, _i, _vp, _vp, _i64, _i, _vp, _vp, _vp, _i, _vp]),
    "wm_loss_former": (_i, [_vp, _i, _vp, _vp, _i64, _i64, _i, _i, _i, _f, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "wm_loss_bert_grad": (_i, [_vp, _i, _vp, _vp, _i64, _i, _vp, _vp, _vp, _i, _vp]),
    "wm_loss_former_grad": (_i, [_vp, _i, _vp, _vp, _i64, _i64, _i, _i, _i, _f, _vp, _vp, _vp, _i, _vp]),
    "wm_yield_head_param_count": (_i, [_i, _i, _i]),
    "wm_yield_head_fwd": (_i, [_vp, _i, _i, _vp, _vp, _i64, _i64, _vp, _vp, _i] + [_vp] * 8 + [_vp, _vp, _i, _i, _i, _i, _vp]),
    "wm_yield_head_bwd": (_i, [_vp, _vp, _i, _i, _vp, _i64, _i64, _vp, _vp, _vp, _i] + [_vp] * 8 +
                          [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "wm_adam_fused": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _f, _d, _d, _f, _f, _i, _f, _vp]),
    "wm_adam_fused_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _d, _d, _f, _f, _f, _vp]),
    "wm_step_params_apply": (_i, [_vp, _vp]),
    "wm_encoder_param_count": (_i64, [C.POINTER(EncoderConfig)]),
    "wm_encoder_param_layout": (_i, [C.POINTER(EncoderConfig), C.POINTER(_i64), _i]),
    "wm_encoder_workspace_bytes": (_sz, [C.POINTER(EncoderConfig)]),
    "wm_encoder_create": (_i, [C.POINTER(EncoderConfig), _vp, _sz, C.POINTER(_vp)]),
    "wm_encoder_destroy": (_i, [_vp]),
    "wm_encoder_refresh_weights": (_i, [_vp, _vp, _vp]),
    "wm_encoder_shadow": (_vp, [_vp]),
    "wm_encoder_refresh_transposes": (_i, [_vp, _vp, _vp]),
    "wm_encoder_forward": (_i, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i, _i, _u64, _u64, _vp]),
    "wm_encoder_backward_head": (_i, [_vp, _vp, _vp, _vp]),
    "wm_encoder_backward_layers": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "wm_encoder_backward_embed": (_i, [_vp, _vp, _vp]),
    "wm_encoder_activation": (_vp, [_vp, _i, _i]),
}

_lib = None


def lib():
    """Load (once) and return the shared library; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m weathermodel_b200.build` "
                "(there is no CPU or PyTorch fallback for the hot path)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.wm_abi_version() != 3:
            raise RuntimeError("libwm_b200.so ABI version mismatch")
        # WM_OPTIONS="name=value,..." applies wm_set_option tuning switches at load time (A/B measurements)
        for item in filter(None, os.environ.get("WM_OPTIONS", "").split(",")):
            name, _, value = item.partition("=")
            if handle.wm_set_option(name.strip().encode(), int(value or 1)) != 0:
                raise RuntimeError(f"WM_OPTIONS: unknown option {name!r}")
        _lib = handle
    return _lib


def check(code: int, what: str = "libwm_b200"):
    if code != 0:
        msg = lib().wm_strerror(code).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {code})")
