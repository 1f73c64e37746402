"""A/B of the wgrad tile width (wm_set_option("wgrad_bn", v)) on the four weight-gradient shapes of one encoder layer."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weathermodel_b200 import ops
M, D, FF = 186880, 576, 2304
bf = lambda *s: (torch.randn(*s, device="cuda") * 0.5).to(torch.bfloat16)
def t(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, n, k in (("qkv", 3 * D, D), ("out", D, D), ("lin1", FF, D), ("lin2", D, FF)):
    a, b = bf(M, n), bf(M, k)
    ref = None
    line = f"wgrad {name:5s} [{n}x{k}]"
    for bn in (0, 128, 192, 256):
        ops.lib().wm_set_option(b"wgrad_bn", bn)
        dw, db = ops.gemm_wgrad(a, b, want_bias_grad=True)
        if ref is None: ref = (dw.clone(), db.clone())
        err = ((dw - ref[0]).abs().max() / ref[0].abs().max()).item()
        ms = t(lambda: ops.gemm_wgrad(a, b, want_bias_grad=True))
        line += f"  bn={bn}: {ms:.4f} ms ({2.0 * M * n * k / ms / 1e9:.0f} TF, err {err:.1e})"
    ops.lib().wm_set_option(b"wgrad_bn", 0)
    print(line)
print("device_error", ops.device_error())
