"""One forward + backward attention launch at the WeatherFormer-large shape (for `ncu -k regex:attn_`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weathermodel_b200 import ops  # noqa: E402

B, S, H, dh = 512, 365, 16, 36
p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
qkv = (torch.randn(B * S, 3 * H * dh, device="cuda") * 0.5).to(torch.bfloat16)
dctx = (torch.randn(B * S, H * dh, device="cuda") * 0.5).to(torch.bfloat16)
for _ in range(2):
    ctx, lse = ops.attn_fwd(qkv, B, S, H, dh, dropout_p=p, seed=1, stream_id=1)
    ops.attn_bwd(qkv, ctx, dctx, lse, B, S, H, dh, dropout_p=p)
torch.cuda.synchronize()
