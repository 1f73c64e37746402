"""LayerNorm backward in the encoder's form (no bias gradient, dropout mask re-applied) at the large shape, for ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weathermodel_b200 import ops
M, D = 186880, 576
x = (torch.randn(M, D, device="cuda") * 0.5).to(torch.bfloat16)
dy = (torch.randn(M, D, device="cuda") * 0.5).to(torch.bfloat16)
gamma, beta = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
y, mean, rstd = ops.layernorm_fwd(x, gamma, beta)
for _ in range(4):
    ops.layernorm_bwd(dy, x, gamma, mean, rstd, dropout_p=0.1, seed=1, stream_id=1, want_bias_grad=False)
torch.cuda.synchronize()
