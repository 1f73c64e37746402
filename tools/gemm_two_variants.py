"""Two gemm_tn variants on the linear1 shape with the bias + ReLU + dropout + sign-bit epilogue of the encoder (for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weathermodel_b200 import ops
M, N, K = 186880, 2304, 576
a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
bias = torch.zeros(N, device="cuda")
bits = ops.gemm_sign_bits(M, N, "cuda")
for two, ew, stg in ((1, 8, 0), (1, 16, 2)):
    for name, v in (("gemm_two_cta", two), ("gemm_epi_warps", ew), ("gemm_staged", stg)):
        ops.lib().wm_set_option(name.encode(), v)
    for _ in range(3):
        ops.gemm_tn(a, w, bias=bias, relu=True, dropout_p=0.1, seed=1, stream_id=2, sign_bits_out=bits)
torch.cuda.synchronize()
