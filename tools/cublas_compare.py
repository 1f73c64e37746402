"""Information only: cuBLAS (torch.matmul / F.linear, bf16) on the GEMM shapes of one encoder layer against gemm_tn,
(1) plain product against plain product, per kernel variant, and (2) every call site with its fused epilogue against
what a library user runs for the same result (F.linear with bias, then eager ReLU / dropout / residual ops).
Not part of the product path.   python tools/cublas_compare.py [large|medium|small]"""
import os, sys
import torch
import torch.nn.functional as Fn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weathermodel_b200 import ops
from weathermodel_b200._lib import lib

SIZES = {"mini": (4, 2, 12, 64), "small": (10, 4, 20, 128), "medium": (12, 6, 28, 256), "large": (16, 8, 36, 512)}
H, L, f, B = SIZES[sys.argv[1] if len(sys.argv) > 1 else "large"]
D, FF, M = H * f, 4 * H * f, B * 365


def t(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def set_variant(two, ew, stg):
    lib().wm_set_option(b"gemm_two_cta", two); lib().wm_set_option(b"gemm_epi_warps", ew); lib().wm_set_option(b"gemm_staged", stg)


def best_variant(fn):
    times = {}
    for var in ops.GEMM_VARIANTS:
        set_variant(*var)
        times[var] = t(fn)
    set_variant(-1, 0, -1)
    v = min(times, key=times.get)
    return v, times[v]


bf = lambda *s: (torch.randn(*s, device="cuda") * 0.5).to(torch.bfloat16)
print(f"# plain products, M = {M}")
for N, K in [(3 * D, D), (D, D), (FF, D), (D, FF), (D, 3 * D)]:
    a, w = bf(M, K), (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    ms_c = t(lambda: torch.matmul(a, w.t()))
    var, ms_o = best_variant(lambda: ops.gemm_tn(a, w))
    fl = 2.0 * M * N * K / 1e9
    print(f"[{M}x{N}x{K}] cuBLAS {ms_c:.4f} ms {fl/ms_c:7.1f} TF | gemm_tn {ms_o:.4f} ms {fl/ms_o:7.1f} TF (variant {var}) | ratio {ms_c/ms_o:.3f}", flush=True)

print("# call sites with their epilogues: gemm_tn (fused) vs F.linear + eager elementwise ops")
x, x2, h, qkv = bf(M, D), bf(M, D), bf(M, FF), bf(M, 3 * D)
bits = torch.zeros(lib().wm_gemm_sign_bits_bytes(M, FF), dtype=torch.uint8, device="cuda")
drop = dict(dropout_p=0.1, seed=1, stream_id=2)
sites = [
    ("F1 qkv (bias)", x, 3 * D, D, dict(bias=True), lambda a, w, b: Fn.linear(a, w, b)),
    ("F2 out-proj (bias+drop+res)", x, D, D, dict(bias=True, residual=x2, **drop), lambda a, w, b: x2 + Fn.dropout(Fn.linear(a, w, b), 0.1, True)),
    ("F3 linear1 (bias+relu+drop)", x, FF, D, dict(bias=True, relu=True, sign_bits_out=bits, **drop), lambda a, w, b: Fn.dropout(torch.relu(Fn.linear(a, w, b)), 0.1, True)),
    ("F4 linear2 (bias+drop+res)", h, D, FF, dict(bias=True, residual=x2, **drop), lambda a, w, b: x2 + Fn.dropout(Fn.linear(a, w, b), 0.1, True)),
    ("B1 linear2 dgrad (relu gate)", x, FF, D, dict(gate_bits=bits, gate_scale=1.0 / 0.9), lambda a, w, b: torch.where(h > 0, Fn.linear(a, w) * (1.0 / 0.9), 0.0)),
    ("B2 linear1 dgrad (+res)", h, D, FF, dict(residual=x2), lambda a, w, b: x2 + Fn.linear(a, w)),
    ("B3 out-proj dgrad (plain)", x, D, D, dict(), lambda a, w, b: Fn.linear(a, w)),
    ("B4 qkv dgrad (+res)", qkv, D, 3 * D, dict(residual=x2), lambda a, w, b: x2 + Fn.linear(a, w)),
]
tot_o = tot_c = tot_p = 0.0
for name, A, N, K, kw, ref_fn in sites:
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    kw = dict(kw)
    bias32 = torch.zeros(N, device="cuda") if kw.pop("bias", False) else None
    if bias32 is not None: kw["bias"] = bias32
    bias16 = bias32.to(torch.bfloat16) if bias32 is not None else None
    var, ms_o = best_variant(lambda: ops.gemm_tn(A, w, **kw))
    ms_c = t(lambda: ref_fn(A, w, bias16))
    ms_p = t(lambda: torch.matmul(A, w.t()))
    tot_o += ms_o; tot_c += ms_c; tot_p += ms_p
    fl = 2.0 * M * N * K / 1e9
    print(f"{name:32s} [{M}x{N}x{K}] gemm_tn {ms_o:.4f} ms {fl/ms_o:7.1f} TF {var} | cuBLAS product alone {ms_p:.4f} | cuBLAS + eager epilogue {ms_c:.4f}", flush=True)
print(f"per layer: gemm_tn {tot_o:.3f} ms | cuBLAS products alone {tot_p:.3f} ms | cuBLAS + eager epilogues {tot_c:.3f} ms")

print("# wgrad shapes: dW[N,K] = A[M,N]^T B[M,K]")
for N, K in [(3 * D, D), (D, D), (FF, D), (D, FF)]:
    a, b = bf(M, N), bf(M, K)
    ms_c = t(lambda: torch.matmul(a.t(), b))
    ms_o = t(lambda: ops.gemm_wgrad(a, b, want_bias_grad=True))
    fl = 2.0 * M * N * K / 1e9
    print(f"wgrad [{N}x{K} over {M}] cuBLAS {ms_c:.4f} ms {fl/ms_c:7.1f} TF (bf16 out, no bias grad) | gemm_wgrad {ms_o:.4f} ms {fl/ms_o:7.1f} TF (fp32 out + bias grad)", flush=True)
print("device_error", ops.device_error())
