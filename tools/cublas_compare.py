"""Information only: cuBLAS (torch.matmul, bf16) on the same GEMM shapes as gemm_tn, to see how much headroom the
hand-written kernel has left on each shape. Not part of the product path."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weathermodel_b200 import ops

def t(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

M = 186880
for N, K in [(1728, 576), (576, 576), (2304, 576), (576, 2304), (576, 1728), (8192, 8192)]:
    m = M if N != 8192 else 8192
    a = (torch.randn(m, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    ms_c = t(lambda: torch.matmul(a, w.t()))
    ms_o = t(lambda: ops.gemm_tn(a, w))
    fl = 2.0 * m * N * K / 1e9
    print(f"[{m}x{N}x{K}] cuBLAS {ms_c:.4f} ms {fl/ms_c:7.1f} TF | gemm_tn {ms_o:.4f} ms {fl/ms_o:7.1f} TF", flush=True)
# wgrad shapes: dW[N,K] = A[M,N]^T B[M,K]
for N, K in [(1728, 576), (576, 576), (2304, 576), (576, 2304)]:
    a = (torch.randn(M, N, device="cuda") * 0.5).to(torch.bfloat16)
    b = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    ms_c = t(lambda: torch.matmul(a.t(), b))
    ms_o = t(lambda: ops.gemm_wgrad(a, b))
    fl = 2.0 * M * N * K / 1e9
    print(f"wgrad [{N}x{K} over {M}] cuBLAS {ms_c:.4f} ms {fl/ms_c:7.1f} TF | gemm_wgrad {ms_o:.4f} ms {fl/ms_o:7.1f} TF", flush=True)
