"""A few gemm_tn launches at one WeatherFormer-large shape (for `ncu -k regex:gemm_tn`).
    python tools/gemm_once.py N K [res|relu|plain]      (M = 186880; WM_OPTIONS selects the kernel variant)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weathermodel_b200 import ops  # noqa: E402

M = 186880
N, K = int(sys.argv[1]), int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "plain"
a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
bias = torch.zeros(N, device="cuda")
res = (torch.randn(M, N, device="cuda") * 0.5).to(torch.bfloat16) if mode == "res" else None
for _ in range(3):
    ops.gemm_tn(a, w, bias=bias, relu=(mode == "relu"), residual=res)
torch.cuda.synchronize()
