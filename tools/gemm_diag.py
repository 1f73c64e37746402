"""Bisect of the gemm_tn epilogue cost with a DIAGNOSTIC build of the library (-DWM_DIAG, tools/_diag/; never the
product .so): which part of the epilogue sets the pace of the K = 576 GEMMs?
    nvcc ... -DWM_DIAG -o tools/_diag/libwm_b200_diag.so ; WM_B200_LIB=tools/_diag/libwm_b200_diag.so python tools/gemm_diag.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weathermodel_b200 import ops
from weathermodel_b200._lib import lib

def t(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

M = 186880
for N, K in [(2304, 576), (576, 576), (1728, 576), (576, 2304)]:
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    for ew, stg in ((8, 0), (16, 0), (16, 1)):
        lib().wm_set_option(b"gemm_two_cta", 0); lib().wm_set_option(b"gemm_epi_warps", ew); lib().wm_set_option(b"gemm_staged", stg)
        line = []
        for diag, label in ((0, "full"), (8, "compact stores"), (1, "no stores"), (2, "no math/stores"), (6, "no tmem ld")):
            assert lib().wm_set_option(b"gemm_diag", diag) == 0
            ms = t(lambda: ops.gemm_tn(a, w, bias=bias, relu=True))
            line.append(f"{label}: {ms:.4f} ms {2.0*M*N*K/ms/1e9:5.0f} TF")
        lib().wm_set_option(b"gemm_diag", 0)
        print(f"N={N} K={K} epi_warps={ew:2d} staged={stg}  " + " | ".join(line), flush=True)
print("device_error", ops.device_error())
