"""Diagnostic: where does a recorded step differ from the eager step? (Adam kernel variants; gradients of one step)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import wm_oracle as O
from weathermodel_b200 import engine, ops
from weathermodel_b200.graph_step import CapturedTrainStep
from weathermodel_b200.optim import FusedAdam
from weathermodel_b200.pretraining.models.weatherformer import WeatherFormer
dev = "cuda"
torch.manual_seed(0)
p = torch.randn(100000, device=dev); g = torch.randn(100000, device=dev) * 1e-2
m = torch.randn(100000, device=dev) * 1e-3; v = torch.rand(100000, device=dev) * 1e-5
for k in (1, 3, 100):
    p1, m1, v1 = p.clone(), m.clone(), v.clone()
    ops.adam_fused(p1, g, m1, v1, k, 1e-3)
    p2, m2, v2 = p.clone(), m.clone(), v.clone()
    hyper = torch.tensor([1e-3, 1.0 - 0.9 ** k, (1.0 - 0.999 ** k) ** 0.5], device=dev)
    ops.adam_fused_dev(p2, g, m2, v2, hyper)
    d = (p1 - p).abs().max().item()
    print(f"adam step {k}: max |dp eager - dp dev| {(p1 - p2).abs().max().item():.3e} (update size {d:.3e}), m {(m1 - m2).abs().max().item():.1e} v {(v1 - v2).abs().max().item():.1e}")
torch.manual_seed(31)
model = WeatherFormer(31, 31, torch.device(dev), **O.get_model_params("mini")).to(dev).train()
for mod in model.modules():
    if isinstance(mod, torch.nn.Dropout): mod.p = 0.0
    if isinstance(mod, torch.nn.MultiheadAttention): mod.dropout = 0.0
opt = FusedAdam(model.parameters(), lr=1e-3, runtime=model.runtime)
def batch(i):
    w, c, y, iv = (torch.from_numpy(a).to(dev) for a in O.synthetic_batch(8, 365, seed=40 + i))
    g_ = torch.Generator(device=dev).manual_seed(i)
    return w, c, y, iv, (torch.rand(8, 31, device=dev, generator=g_) < 0.3).unsqueeze(1).expand(-1, 365, -1).contiguous()
def loss_fn(w, c, y, iv, mask):
    return engine.former_elbo(model.forward_raw(w, c, y, iv, mask), w, mask, 0.5)
def eager(b):
    opt.zero_grad(); l = loss_fn(*b); l["total_loss"].backward(); opt.step(); return l["total_loss"].item()
rt = model.runtime
for i in range(2): eager(batch(i))
snap = (rt.flat_params.clone(), opt._flat_m.clone(), opt._flat_v.clone(), opt._step)
l_e = eager(batch(2)); g_e = rt.flat_grads.clone(); p_e = rt.flat_params.clone(); m_e = opt._flat_m.clone(); v_e = opt._flat_v.clone()
rt.flat_params.copy_(snap[0]); opt._flat_m.copy_(snap[1]); opt._flat_v.copy_(snap[2]); opt._step = snap[3]; rt.mark_weights_dirty()
step = CapturedTrainStep(opt, loss_fn, batch(2))
l_g = step.first_losses["total_loss"].item()
rel = lambda a, b: ((a - b).norm() / b.norm()).item()
print(f"loss eager {l_e!r} recorded {l_g!r}")
print(f"grads rel diff {rel(rt.flat_grads, g_e):.3e}  params rel diff of update {rel(rt.flat_params - snap[0], p_e - snap[0]):.3e}  m {rel(opt._flat_m, m_e):.3e}  v {rel(opt._flat_v, v_e):.3e}")
print("hyper on device", step._dev_hyper.tolist(), "expected", [1e-3, 1 - 0.9 ** 3, (1 - 0.999 ** 3) ** 0.5])
