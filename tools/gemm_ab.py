"""A/B of the gemm_tn variants on the large-config shapes: single-CTA vs CTA-pair tiles (gemm_two_cta) x 8 vs 16
epilogue warps (gemm_epi_warps). Every variant must be bit-identical."""
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

M = int(os.environ.get("M", 186880))
for N, K in [(1728, 576), (576, 576), (2304, 576), (576, 2304), (576, 1728)]:
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    res = (torch.randn(M, N, device="cuda") * 0.5).to(torch.bfloat16)
    for label, kw in (("bias", {}), ("bias+relu+drop", dict(relu=True, dropout_p=0.1, seed=3, stream_id=9)),
                      ("bias+res", dict(residual=res))):
        ref = None
        for two, tma in ((0, 8), (0, 16), (1, 8), (1, 16)):
            lib().wm_set_option(b"gemm_two_cta", two)
            lib().wm_set_option(b"gemm_epi_warps", tma)
            out = ops.gemm_tn(a, w, bias=bias, **kw)
            if ref is None: ref = out
            else: assert torch.equal(ref, out), f"variant two_cta={two} epi_warps={tma:2d} differs"
            ms = t(lambda: ops.gemm_tn(a, w, bias=bias, **kw))
            print(f"N={N:5d} K={K:5d} {label:15s} two_cta={two} epi_warps={tma:2d}: {ms:.4f} ms  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
print("device_error", ops.device_error())
