import sys, torch
sys.path.insert(0, "/root/repo")
from weathermodel_b200 import ops
B, S, D = 512, 365, 576
dev = "cuda"
w = torch.randn(B, S, 31, device=dev); mask = torch.rand(B, S, 31, device=dev) < 0.15
yr = torch.rand(B, S, device=dev) + 1990; co = torch.rand(B, 2, device=dev)
w_in, b_in, pe = torch.randn(D, 34, device=dev), torch.zeros(D, device=dev), torch.randn(365, D, device=dev)
for _ in range(2):
    ops.embed_fwd(w, mask, yr, co, w_in, b_in, pe, want_xin=True)
torch.cuda.synchronize()
