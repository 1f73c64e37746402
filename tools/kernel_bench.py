"""Per-kernel timing of the hot path at a BASELINE workload's shapes (CUDA events, kernel timed alone, operands
larger than L2). Used to find where a training step's time goes and as the `ncu --set full` target.

    python tools/kernel_bench.py [--workload large] [--only gemm|wgrad|attn|mem] [--reps 10]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weathermodel_b200 import ops  # noqa: E402

SIZES = {"mini": (4, 2, 12, 64), "small": (10, 4, 20, 128), "medium": (12, 6, 28, 256), "large": (16, 8, 36, 512)}


def timeit(fn, reps):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="large")
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=0)
    a = ap.parse_args()
    H, L, f, B = SIZES[a.workload]
    B = a.batch or B
    D, FF, S = H * f, 4 * H * f, 365
    M = B * S
    dev = "cuda"
    bf = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(torch.bfloat16)  # noqa: E731
    rows = []

    def rec(name, ms, flops=0.0, bytes_=0.0, per_layer=1):
        rows.append({"kernel": name, "ms": round(ms, 4), "TFLOP/s": round(flops / ms / 1e9, 1) if flops else None,
                     "GB/s": round(bytes_ / ms / 1e6, 1) if bytes_ else None, "per_step": per_layer})
        print(rows[-1], flush=True)

    x, h = bf(M, D), bf(M, FF)
    qkv = bf(M, 3 * D)
    if a.only in ("", "gemm"):
        for name, A, N, K, kw in [("qkv fwd", x, 3 * D, D, {}), ("out-proj fwd (+res)", x, D, D, {"residual": x}),
                                  ("linear1 fwd (relu)", x, FF, D, {"relu": True}), ("linear2 fwd (+res)", h, D, FF, {"residual": x}),
                                  ("qkv dgrad (+res)", qkv, D, 3 * D, {"residual": x}), ("linear1 dgrad (+res)", h, D, FF, {"residual": x}),
                                  ("linear2 dgrad (gate)", x, FF, D, {"gate": h})]:
            w = (torch.randn(N, K, device=dev) * K ** -0.5).to(torch.bfloat16)
            bias = torch.zeros(N, device=dev)
            ms = timeit(lambda: ops.gemm_tn(A, w, bias=bias, **kw), a.reps)
            rec(f"gemm_tn {name} [{M}x{N}x{K}]", ms, 2.0 * M * N * K, per_layer=L)
            ms = timeit(lambda: ops.gemm_tn(A, w, bias=bias, dropout_p=0.1, seed=1, stream_id=2, **kw), a.reps)
            rec(f"gemm_tn {name} +dropout", ms, 2.0 * M * N * K, per_layer=L)
    if a.only in ("", "wgrad"):
        for name, A, Bm in [("dW_qkv", qkv, x), ("dW_o", x, x), ("dW1", h, x), ("dW2", x, h)]:
            ms = timeit(lambda: ops.gemm_wgrad(A, Bm), a.reps)
            rec(f"gemm_wgrad {name} [{A.shape[1]}x{Bm.shape[1]} over {M}]", ms, 2.0 * M * A.shape[1] * Bm.shape[1], per_layer=L)
    if a.only in ("", "attn"):
        for p in (0.0, 0.1):
            ms = timeit(lambda: ops.attn_fwd(qkv, B, S, H, f, dropout_p=p, seed=1, stream_id=1), a.reps)
            rec(f"attn_fwd p={p}", ms, 4.0 * S * S * D * B, per_layer=L)
            ctx, lse = ops.attn_fwd(qkv, B, S, H, f, dropout_p=p, seed=1, stream_id=1)
            ms = timeit(lambda: ops.attn_bwd(qkv, ctx, x, lse, B, S, H, f, dropout_p=p), a.reps)
            rec(f"attn_bwd p={p}", ms, 8.0 * S * S * D * B, per_layer=L)
    if a.only in ("", "mem"):
        gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
        ms = timeit(lambda: ops.layernorm_fwd(x, gamma, beta), a.reps)
        rec("layernorm_fwd", ms, bytes_=4.0 * M * D, per_layer=2 * L)
        y, mean, rstd = ops.layernorm_fwd(x, gamma, beta)
        ms = timeit(lambda: ops.layernorm_bwd(x, x, gamma, mean, rstd), a.reps)
        rec("layernorm_bwd", ms, bytes_=6.0 * M * D, per_layer=0)
        ms = timeit(lambda: ops.layernorm_bwd(x, x, gamma, mean, rstd, dropout_p=0.1, seed=1, stream_id=1), a.reps)
        rec("layernorm_bwd +dropout (with bias gradient, 8 row warps)", ms, bytes_=8.0 * M * D, per_layer=0)
        ms = timeit(lambda: ops.layernorm_bwd(x, x, gamma, mean, rstd, dropout_p=0.1, seed=1, stream_id=1, want_bias_grad=False), a.reps)
        rec("layernorm_bwd +dropout (encoder form, 15 row warps)", ms, bytes_=8.0 * M * D, per_layer=2 * L)
        ms = timeit(lambda: ops.colsum(h), a.reps)
        rec("colsum [M,FF]", ms, bytes_=2.0 * M * FF, per_layer=L)
        ms = timeit(lambda: ops.colsum(qkv), a.reps)
        rec("colsum [M,3D]", ms, bytes_=2.0 * M * 3 * D, per_layer=L)
        w = torch.randn(B, S, 31, device=dev)
        mask = torch.rand(B, S, 31, device=dev) < 0.15
        yr = torch.rand(B, S, device=dev) + 1990
        co = torch.rand(B, 2, device=dev)
        w_in, b_in, pe = torch.randn(D, 34, device=dev), torch.zeros(D, device=dev), torch.randn(365, D, device=dev)
        ms = timeit(lambda: ops.embed_fwd(w, mask, yr, co, w_in, b_in, pe, want_xin=True), a.reps)
        rec("embed_fwd", ms, bytes_=M * (31 * 5 + 4 + 2 * D + 128.0))
        n = 32_000_000
        p_, g_, m_, v_ = (torch.zeros(n, device=dev) for _ in range(4))
        ms = timeit(lambda: ops.adam_fused(p_, g_, m_, v_, 1, 1e-3), a.reps)
        rec("adam 32M", ms, bytes_=28.0 * n)
    est = sum(r["ms"] * r["per_step"] for r in rows)
    print(json.dumps({"workload": a.workload, "sum_ms_weighted": est}))


if __name__ == "__main__":
    main()
