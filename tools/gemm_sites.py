"""The eight gemm_tn call sites of one encoder layer (forward + dgrad) with their real epilogues, timed under every
kernel variant (gemm_two_cta, gemm_epi_warps, gemm_staged). Decides the per-site variant choice in launch_gemm_tn_impl.
    python tools/gemm_sites.py [workload]"""
import os, sys
import torch
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

bf = lambda *s: (torch.randn(*s, device="cuda") * 0.5).to(torch.bfloat16)
x, x2, h, qkv = bf(M, D), bf(M, D), bf(M, FF), bf(M, 3 * D)
bits = torch.zeros(lib().wm_gemm_sign_bits_bytes(M, FF), dtype=torch.uint8, device="cuda")
drop = dict(dropout_p=0.1, seed=1, stream_id=2)
sites = [  # name, A, N, K, kwargs
    ("F1 qkv (bias)", x, 3 * D, D, dict(bias=True)),
    ("F2 out-proj (bias+drop+res)", x, D, D, dict(bias=True, residual=x2, **drop)),
    ("F3 linear1 (bias+relu+drop+bits)", x, FF, D, dict(bias=True, relu=True, sign_bits_out=bits, **drop)),
    ("F4 linear2 (bias+drop+res)", h, D, FF, dict(bias=True, residual=x2, **drop)),
    ("B1 linear2 dgrad (gate bits)", x, FF, D, dict(gate_bits=bits, gate_scale=1.0 / 0.9)),
    ("B2 linear1 dgrad (+res)", h, D, FF, dict(residual=x2)),
    ("B3 out-proj dgrad (plain)", x, D, D, dict()),
    ("B4 qkv dgrad (+res)", qkv, D, 3 * D, dict(residual=x2)),
]
TILE_N = int(os.environ.get("TILE_N", 0))  # force the tile width (sites whose N it does not divide keep the default)
total = {}
for name, A, N, K, kw in sites:
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    kw = dict(kw)
    if kw.pop("bias", False): kw["bias"] = torch.zeros(N, device="cuda")
    if TILE_N and N % TILE_N == 0: kw["tile_n"] = TILE_N
    ref, line = None, []
    for two, ew, stg in ops.GEMM_VARIANTS:
        lib().wm_set_option(b"gemm_two_cta", two); lib().wm_set_option(b"gemm_epi_warps", ew); lib().wm_set_option(b"gemm_staged", stg)
        out = ops.gemm_tn(A, w, **kw)
        if ref is None: ref = out
        else: assert torch.equal(ref, out), f"{name}: variant ({two},{ew},{stg}) differs"
        ms = t(lambda: ops.gemm_tn(A, w, **kw))
        total[(two, ew, stg)] = total.get((two, ew, stg), 0.0) + ms
        line.append(f"({two},{ew:2d},{stg}) {ms:.4f}")
    print(f"{name:34s} [{M}x{N}x{K}]  " + "  ".join(line), flush=True)
print("sum per layer:", {k: round(v, 4) for k, v in total.items()})
print("device_error", ops.device_error())
