"""A/B of the A^T tiles per CTA (wm_set_option("wgrad_mh", 1 | 2)) on the four weight-gradient shapes of one encoder layer."""
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
tot = {1: 0.0, 2: 0.0}
for name, n, k in (("qkv", 3 * D, D), ("out", D, D), ("lin1", FF, D), ("lin2", D, FF)):
    a, b = bf(M, n), bf(M, k)
    ref = (a.float().t() @ b.float(), a.float().sum(0))
    line = f"wgrad {name:5s} [{n}x{k}]"
    for mh in (1, 2, 1, 2):
        ops.lib().wm_set_option(b"wgrad_mh", mh)
        dw, db = ops.gemm_wgrad(a, b, want_bias_grad=True)
        err = max(((dw - ref[0]).abs().max() / ref[0].abs().max()).item(), ((db - ref[1]).abs().max() / ref[1].abs().max()).item())
        ms = t(lambda: ops.gemm_wgrad(a, b, want_bias_grad=True))
        tot[mh] += ms / 2
        line += f"  mh={mh}: {ms:.4f} ms ({2.0 * M * n * k / ms / 1e9:.0f} TF, err {err:.1e})"
    ops.lib().wm_set_option(b"wgrad_mh", 0)
    print(line)
print("sum per layer:", {k: round(v, 4) for k, v in tot.items()}, "device_error", ops.device_error())
