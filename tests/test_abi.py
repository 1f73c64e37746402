"""C-ABI checks that need no GPU: the library loads, exports every symbol include/wm_b200.h declares, the
ctypes table covers exactly those symbols, and the host-only entry points (layout, sizing, errors) behave."""
import ctypes as C
import os
import re

import pytest

from weathermodel_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "wm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wm_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/wm_b200.h but not exported by libwm_b200.so"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES and the header disagree"
    assert lib.wm_abi_version() == 3
    assert b"no fallback" in lib.wm_strerror(6)


@pytest.mark.parametrize("size,kind,expected", [("mini", 31, 59743), ("mini", 62, 61262), ("small", 62, 1949862),
                                                ("medium", 31, 8176927), ("large", 62, 31966334), ("large", 31, 31948447)])
def test_param_layout_matches_reference_parameter_counts(size, kind, expected):
    """parameter counts of the reference models; the flat layout holds exactly those tensors."""
    heads, layers, factor = {"mini": (4, 2, 12), "small": (10, 4, 20), "medium": (12, 6, 28), "large": (16, 8, 36)}[size]
    D = heads * factor
    cfg = _lib.EncoderConfig(8, 365, 31, D, heads, layers, 4 * D, kind, 0.1, 1e-5)
    lib = _lib.lib()
    total = lib.wm_encoder_param_count(C.byref(cfg))
    n = lib.wm_encoder_param_layout(C.byref(cfg), None, 0)
    assert n == 4 + 12 * layers
    offs = (C.c_int64 * n)()
    assert lib.wm_encoder_param_layout(C.byref(cfg), offs, n) == n
    offs = list(offs)
    assert offs[0] == 0 and all(o % 64 == 0 for o in offs) and offs == sorted(offs) and total >= offs[-1] + kind
    sizes = [D * 34, D] + [3 * D * D, 3 * D, D * D, D, 4 * D * D, 4 * D, 4 * D * D, D, D, D, D, D] * layers + [kind * D, kind]
    assert sum(sizes) == expected
    for o, s, nxt in zip(offs, sizes, offs[1:] + [total]):
        assert o + s <= nxt, "tensors overlap in the flat layout"


def test_workspace_sizing_and_rejections():
    lib = _lib.lib()
    ok = _lib.EncoderConfig(512, 365, 31, 576, 16, 8, 2304, 62, 0.1, 1e-5)
    ws = lib.wm_encoder_workspace_bytes(C.byref(ok))
    assert 15 * 2 ** 30 < ws < 40 * 2 ** 30  # ~22 GB of saved activations for config 4 on one B200
    for bad in [
        _lib.EncoderConfig(8, 400, 31, 576, 16, 8, 2304, 62, 0.1, 1e-5),   # seq_len > 384
        _lib.EncoderConfig(8, 365, 31, 580, 16, 8, 2304, 62, 0.1, 1e-5),   # D not divisible by heads
        _lib.EncoderConfig(8, 365, 31, 512, 8, 8, 2048, 62, 0.1, 1e-5),    # head_dim 64 > 48
        _lib.EncoderConfig(8, 365, 31, 576, 16, 8, 2304, 70, 0.1, 1e-5),   # head too wide
        _lib.EncoderConfig(8, 365, 31, 576, 16, 8, 2304, 62, 1.0, 1e-5),   # dropout 1.0
    ]:
        assert lib.wm_encoder_workspace_bytes(C.byref(bad)) == 0
        assert lib.wm_encoder_param_count(C.byref(bad)) == -1
    out = C.c_void_p()
    assert lib.wm_encoder_create(C.byref(ok), None, 0, C.byref(out)) != 0  # NULL workspace refused, no crash
    assert lib.wm_gemm_wgrad_workspace_bytes(186880, 1728, 576) > 0
    assert lib.wm_layernorm_bwd_workspace_bytes(186880, 576) == 592 * 3 * 576 * 4
