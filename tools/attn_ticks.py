"""Phase timeline (clock64 deltas, CTA 0) of attn_fwd at the large config: where do a head's cycles go?"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weathermodel_b200 import ops
from weathermodel_b200._lib import lib

B, S, H, dh = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 365, 16, 36
qkv = (torch.randn(B * S, 3 * H * dh, device="cuda") * 0.5).to(torch.bfloat16)
for p in (0.0, 0.1):
    for _ in range(3):
        ops.attn_fwd(qkv, B, S, H, dh, dropout_p=p, seed=1, stream_id=1)
    buf = (C.c_longlong * 64)()
    lib().wm_debug_ticks(buf, 64)
    t = list(buf)
    names = {}
    for it in range(3):
        for c in range(4):
            names[36 + (it * 4 + c) * 2] = f"   MMA: t{it} P chunk{c} seen"
            names[37 + (it * 4 + c) * 2] = f"   MMA: t{it} chunk{c} PV + next S issued"
    for it in range(3):
        for k, nm in enumerate(["tile top", "next Q fixed", "scores ready", "pass1 done", "max exchanged", "chunk0 -> P",
                                "chunk1 -> P", "chunk2 -> P", "chunk3 -> P", "masks + sum exchanged", "O ready",
                                "epilogue done"]):
            names[it * 12 + k] = f"t{it} {nm}"
    prev = t[0]
    print(f"--- attn_fwd v5 p={p}: second head of CTA 0, softmax warp 0 + MMA thread: total {t[35] - t[0]} cycles")
    for i in sorted(names, key=lambda k: t[k]):
        print(f"{names[i]:40s} +{t[i] - prev:7d}  (@{t[i] - t[0]})")
        prev = t[i]


dctx = (torch.randn(B * S, H * dh, device="cuda") * 0.5).to(torch.bfloat16)
for p in (0.0, 0.1):
    ctx, lse = ops.attn_fwd(qkv, B, S, H, dh, dropout_p=p, seed=1, stream_id=1)
    for _ in range(3):
        ops.attn_bwd(qkv, ctx, dctx, lse, B, S, H, dh, dropout_p=p)
    buf = (C.c_longlong * 64)()
    lib().wm_debug_ticks(buf, 64)
    t = list(buf)
    names = {}
    for g in range(0, 18, 2):  # elementwise group 0 handles the even half-tiles of the head
        names[g] = f"EW group 0: half-tile {g:2d} scores seen"
        names[g + 1] = f"EW group 0: half-tile {g:2d} P^T / dS^T written"
    prev = t[0]
    print(f"--- attn_bwd v6 p={p}: second head of CTA 0, elementwise warp 0: {t[17] - t[0]} cycles for 9 of its 18 half-tiles")
    for i in sorted(names):
        print(f"{names[i]:46s} +{t[i] - prev:7d}  (@{t[i] - t[0]})")
        prev = t[i]
