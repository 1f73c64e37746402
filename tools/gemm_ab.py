"""A/B: single-CTA gemm_tn vs CTA-pair gemm_tn2 on the large-config shapes."""
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
for N, K in [(1728, 576), (576, 576), (2304, 576), (576, 2304), (576, 1728)]:
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    ref = None
    for mode in (0, 1):
        lib().wm_set_option(b"gemm_two_cta", mode)
        out = ops.gemm_tn(a, w, bias=bias)
        if ref is None: ref = out
        else: assert torch.equal(ref, out), "2-CTA result differs from 1-CTA"
        ms = t(lambda: ops.gemm_tn(a, w, bias=bias))
        print(f"N={N:5d} K={K:5d} two_cta={mode}: {ms:.4f} ms  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
print("device_error", ops.device_error())
