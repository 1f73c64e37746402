"""One weight-gradient product of the encoder (linear1: dW[2304, 576] over 186,880 tokens, bias gradient fused), for ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weathermodel_b200 import ops
M, n, k = 186880, 2304, 576
if len(sys.argv) > 2:
    n, k = int(sys.argv[1]), int(sys.argv[2])
a = (torch.randn(M, n, device="cuda") * 0.5).to(torch.bfloat16)
b = (torch.randn(M, k, device="cuda") * 0.5).to(torch.bfloat16)
for _ in range(4):
    ops.gemm_wgrad(a, b, want_bias_grad=True)
torch.cuda.synchronize()
