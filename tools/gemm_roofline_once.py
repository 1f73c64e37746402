"""The roofline kernel of bench.py alone: linear1 shape (186,880 x 2304 x 576, bias + ReLU epilogue) under one kernel
variant (default: CTA pairs, 16 epilogue warps, TMA-store epilogue -- what the tuner picks), for `ncu -k regex:gemm_tn`.
    python tools/gemm_roofline_once.py [two_cta epi_warps staged]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weathermodel_b200 import ops  # noqa: E402

two, ew, stg = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (1, 16, 2)
for name, v in (("gemm_two_cta", two), ("gemm_epi_warps", ew), ("gemm_staged", stg)):
    ops.lib().wm_set_option(name.encode(), v)
M, N, K = 186880, 2304, 576
a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
bias = torch.zeros(N, device="cuda")
for _ in range(4):
    ops.gemm_tn(a, w, bias=bias, relu=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.gemm_tn(a, w, bias=bias, relu=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"variant ({two},{ew},{stg}): {ms:.4f} ms = {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")
