"""One small launch of every kernel of libwm_b200.so, for `compute-sanitizer --tool memcheck|racecheck|synccheck`
(the reference has no sanitizer story; the build runs its kernels' unit shapes under one).
    compute-sanitizer --tool memcheck python tools/sanitize_once.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from weathermodel_b200 import engine, ops  # noqa: E402
from weathermodel_b200.optim import FusedAdam  # noqa: E402
from weathermodel_b200.pretraining.models.weatherformer import WeatherFormer  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
bf = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(torch.bfloat16)  # noqa: E731
done = []


def mark(name):
    torch.cuda.synchronize()
    done.append(name)
    print("ok", name, flush=True)


# masks + embedding
m1 = ops.mask_bert(365, 31, 3, 0.15, device=dev)
m2 = ops.mask_former(365, 31, 3, 10, device=dev)
mark("mask_bert / mask_former")
w = torch.randn(3, 365, 31, device=dev)
out, xin = ops.embed_fwd(w, m1, torch.rand(3, 365, device=dev) + 1990, torch.rand(3, 2, device=dev), torch.randn(200, 34, device=dev),
                         torch.zeros(200, device=dev), torch.randn(365, 200, device=dev), want_xin=True)
mark("embed_fwd")
# GEMMs: single-CTA (direct / staged / TMA-store epilogues), CTA pairs, fp32 head, wgrad
a, b = bf(1100, 576), bf(576, 576)
res, bias = bf(1100, 576), torch.randn(576, device=dev)
for two, ew, stg in ((0, 8, 0), (0, 16, 1), (0, 16, 2), (1, 8, 0), (1, 16, 1), (1, 16, 2)):
    for name, v in (("gemm_two_cta", two), ("gemm_epi_warps", ew), ("gemm_staged", stg)):
        ops.lib().wm_set_option(name.encode(), v)
    bits = ops.gemm_sign_bits(1100, 576, dev)
    y = ops.gemm_tn(a, b, bias=bias, relu=True, dropout_p=0.1, seed=1, stream_id=2, sign_bits_out=bits)
    ops.gemm_tn(a, b, gate_bits=bits, gate_scale=1.1)
    ops.gemm_tn(a, b, bias=bias, residual=res, dropout_p=0.1, seed=1, stream_id=3)
    mark(f"gemm_tn variant {(two, ew, stg)}")
for name, v in (("gemm_two_cta", -1), ("gemm_epi_warps", 0), ("gemm_staged", -1)):
    ops.lib().wm_set_option(name.encode(), v)
ops.gemm_tn(a, bf(64, 576), bias=torch.zeros(64, device=dev), out_fp32=True)
mark("gemm_tn fp32 head")
ops.gemm_wgrad(bf(1100, 576), bf(1100, 192), want_bias_grad=True)
ops.gemm_wgrad(bf(1100, 192), bf(1100, 2304), want_bias_grad=True)
mark("gemm_wgrad")
# attention with dropout, forward + backward
B, S, H, dh = 2, 365, 4, 36
qkv, dctx = bf(B * S, 3 * H * dh), bf(B * S, H * dh)
ctx, lse = ops.attn_fwd(qkv, B, S, H, dh, dropout_p=0.1, seed=1, stream_id=1)
ops.attn_bwd(qkv, ctx, dctx, lse, B, S, H, dh, dropout_p=0.1)
ctx, lse = ops.attn_fwd(bf(B * S, 3 * 4 * 12), B, S, 4, 12)
mark("attn_fwd / attn_bwd")
# LayerNorm, column sums, losses, Adam
x = bf(777, 200)
g, be = torch.ones(200, device=dev), torch.zeros(200, device=dev)
yln, mean, rstd = ops.layernorm_fwd(x, g, be)
ops.layernorm_bwd(bf(777, 200), x, g, mean, rstd, dropout_p=0.1, seed=1, stream_id=1, want_bias_grad=False)
ops.layernorm_bwd(bf(777, 200), x, g, mean, rstd)
ops.colsum(x)
mark("layernorm / colsum")
yh = torch.randn(3 * 365, 64, device=dev)
ops.loss_bert(yh[:, :32].contiguous(), w.view(-1, 31), m1.reshape(-1, 31))
ops.loss_former(yh, w, m2, 0.5)
p_, g_, mm, vv = (torch.zeros(5000, device=dev) for _ in range(4))
ops.adam_fused(p_, g_ + 1, mm, vv, 1, 1e-3)
mark("losses / adam")
# whole step through the model classes + the fused yield head
model = WeatherFormer(31, 31, torch.device(dev), num_heads=4, num_layers=2, hidden_dim_factor=12).to(dev).train()
opt = FusedAdam(model.parameters(), lr=1e-3, runtime=model.runtime)
opt.zero_grad()
co, yr, iv = torch.rand(3, 2, device=dev), torch.rand(3, 365, device=dev) + 1990, torch.full((3, 1), 7.0, device=dev)
engine.former_elbo(model.forward_raw(w, co, yr, iv, m2), w, m2, 0.5)["total_loss"].backward()
opt.step()
mark("encoder step")
from weathermodel_b200.crop_yield.models.weatherformer_yield_model import WeatherFormerYieldModel  # noqa: E402

ym = WeatherFormerYieldModel("y", torch.device(dev), 31, 6, num_heads=4, num_layers=2, hidden_dim_factor=12).to(dev).train()
pred, z, mu, var = ym(w[:, :364], co, yr[:, :364], iv, m2[:, :364], torch.randn(3, 7, device=dev))
(pred.sum() + 1e-3 * (mu ** 2 + var).mean()).backward()
mark("yield head")
print("device_error", ops.device_error(), "sections", len(done))
