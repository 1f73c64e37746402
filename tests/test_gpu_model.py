"""Whole-path parity on a real B200, through the drop-in classes (which bind the C ABI via ctypes):
WeatherBERT / WeatherFormer forward + fused loss + backward + FusedAdam against
  (1) the committed golden vectors produced by the unmodified reference (tests/golden, oracle/make_golden.py)
  (2) the numpy oracle (oracle/wm_oracle.py) on freshly seeded inputs at a second size.

Tolerances (north_star: losses and gradients within 1e-3 relative with fp32 accumulation; the product path
keeps activations and GEMM operands in bf16, 2^-9 per rounding):
  loss            |got - ref| / |ref| <= 1e-3 for the total loss; its two ELBO components (reconstruction, KL)
                  within 1e-3 of the TOTAL's scale and 3e-3 of their own (the KL term is ~10% of the total and
                  sits at 1.2e-3 of itself with bf16 weights)
  gradient norms  | ||g|| - ||g_ref|| | / ||g_ref|| <= 5e-3 per parameter tensor, 3e-3 for the global norm (measured: <= 2e-3)
  gradient field  ||g - g_ref||_F / ||g_ref||_F <= 3e-2 per tensor (bf16 operand rounding noise, unbiased)
  outputs         ||y - y_ref||_F / ||y_ref||_F <= 1e-2
Second oracle (oracle/wm_oracle.py EncoderOracleBf16: the same algorithm with a bf16 rounding wherever the kernels
STORE bf16, everything else exact): what is left is fp32 accumulation order, ex2.approx and values sitting on a bf16
rounding boundary (once two computations differ by 1e-4 somewhere, a few per cent of the later roundings fall the
other way, so the two decorrelate with depth) --
  2-layer golden cases   ||g - g_bf16||_F / ||g_bf16||_F <= 6e-3 per tensor, <= 1.5e-3 global; loss within 2e-4
                         (measured 2.1e-3 / 5.0e-4 for WeatherBERT, 4.7e-3 / 1.1e-3 for WeatherFormer)
  4-8 layer models       <= 1.5e-2 per tensor, <= 7e-3 global (measured 0.95-1.2e-2 / 2.4-5.2e-3), always below the
                         distance to the fp32 oracle
i.e. the kernels implement the reference's math to ~1e-3 where that can be observed; the remaining gap to the fp32
reference is the rounding of the bf16 operands north_star prescribes, and is the same size as
torch.autocast(bfloat16)'s on the same GPU (profiles/r02_parity.txt prints all three per tensor).
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import wm_oracle as O  # noqa: E402

from src.pretraining.models.weatherbert import WeatherBERT  # noqa: E402
from src.pretraining.models.weatherformer import WeatherFormer  # noqa: E402
from weathermodel_b200 import engine  # noqa: E402
from weathermodel_b200.optim import FusedAdam  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DEV = "cuda"


def _neutralise_dropout(model):
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def _check_losses(got, ref):
    total = abs(ref["total_loss"])
    assert abs(got["total_loss"] - ref["total_loss"]) <= 1e-3 * total, (got, ref)
    for k, r in ref.items():
        assert abs(got[k] - r) <= min(1e-3 * total, 3e-3 * abs(r)), (k, got[k], r)


def _record(line):
    """Every measured parity figure goes to stdout (pytest -s) and, on the GPU box, into gpurun_out/ so that the
    numbers behind the asserts are kept (the errors were never printed into a kept log)."""
    print(line)
    keep = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(keep):
        with open(os.path.join(keep, "r02_test_parity.log"), "a") as fh:
            fh.write(line + "\n")


def _check_grads(model, ref_grads, tag):
    tot_g, tot_r, tot_d = 0.0, 0.0, 0.0
    worst, worst_norm = (0.0, ""), (0.0, "")
    for name, p in model.named_parameters():
        assert p.grad is not None, f"{tag}: {name} has no gradient"
        g = p.grad.detach().float().cpu().numpy().astype(np.float64)
        r = np.asarray(ref_grads[name], dtype=np.float64)
        assert np.isfinite(g).all(), f"{tag}: {name} gradient not finite"
        ng, nr = np.linalg.norm(g), np.linalg.norm(r)
        tot_g += ng ** 2
        tot_r += nr ** 2
        tot_d += np.linalg.norm(g - r) ** 2
        worst_norm = max(worst_norm, (abs(ng - nr) / (nr + 1e-30), name))
        assert abs(ng - nr) <= 5e-3 * nr + 1e-9, f"{tag}: ||grad {name}|| {ng:.6g} vs {nr:.6g}"
        rel = _rel(g, r)
        worst = max(worst, (rel, name))
        assert rel <= 3e-2, f"{tag}: grad {name} rel fro err {rel:.4g}"
    gn = abs(np.sqrt(tot_g) - np.sqrt(tot_r)) / np.sqrt(tot_r)
    gf = np.sqrt(tot_d / tot_r)
    _record(f"{tag}: gradient rel fro worst {worst[0]:.3e} ({worst[1]}), norm err worst {worst_norm[0]:.3e} ({worst_norm[1]}), "
            f"global rel fro {gf:.3e}, global norm err {gn:.3e}")
    assert gn <= 3e-3, f"{tag}: global grad norm"
    assert gf <= 2e-2, f"{tag}: global gradient field"
    return worst


def _check_grads_bf16_oracle(model, ref_grads, tag, per_tensor=6e-3, global_tol=1.5e-3):
    """Against the bf16-storage oracle: only accumulation order / ex2.approx / rounding-boundary flips remain."""
    tot_r, tot_d, worst = 0.0, 0.0, (0.0, "")
    for name, p in model.named_parameters():
        g = p.grad.detach().float().cpu().numpy().astype(np.float64)
        r = np.asarray(ref_grads[name], dtype=np.float64)
        tot_r += np.linalg.norm(r) ** 2
        tot_d += np.linalg.norm(g - r) ** 2
        worst = max(worst, (_rel(g, r), name))
    gf = float(np.sqrt(tot_d / tot_r))
    _record(f"{tag} vs bf16-storage oracle: gradient rel fro worst {worst[0]:.3e} ({worst[1]}), global rel fro {gf:.3e}")
    assert worst[0] <= per_tensor, f"{tag}: {worst}"
    assert gf <= global_tol, f"{tag}: global {gf}"


def _load_golden(fname, cls):
    g = dict(np.load(os.path.join(GOLD, fname)))
    torch.manual_seed(1234)
    # built on the CPU generator like the golden run (with device=cuda the reference, too, would draw the
    # encoder-layer weights from the CUDA generator instead)
    model = cls(weather_dim=31, output_dim=31, device=torch.device("cpu"), **O.get_model_params("mini"))
    state = {k[len("param/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param/")}
    # our constructor consumes the RNG exactly like the reference: same initial weights bit for bit
    for k, v in model.state_dict().items():
        assert torch.equal(v.cpu(), state[k]), f"initial weight {k} differs from the reference's for seed 1234"
    model = model.to(DEV)
    model.train()
    _neutralise_dropout(model)
    t = lambda k, dt=torch.float32: torch.from_numpy(g[k]).to(DEV).to(dt)  # noqa: E731
    batch = (t("weather"), t("coords"), t("year"), t("interval"), torch.from_numpy(g["mask"]).to(DEV))
    grads = {k[len("grad/"):]: v for k, v in g.items() if k.startswith("grad/")}
    return g, model, batch, grads


def test_weatherbert_matches_reference_golden():
    g, model, (w, c, yr, iv, mask), ref_grads = _load_golden("weatherbert_mini_b8.npz", WeatherBERT)
    y_pad = model.forward_raw(w, c, yr, iv, mask)
    loss = engine.bert_masked_mse(y_pad, w, mask)
    loss.backward()
    assert _rel(y_pad[..., :31].detach().cpu().numpy(), g["y"]) <= 1e-2
    assert abs(loss.item() - g["loss"][0]) <= 1e-3 * abs(g["loss"][0]), (loss.item(), g["loss"][0])
    worst = _check_grads(model, ref_grads, "bert")
    print("bert mini: loss", loss.item(), "ref", g["loss"][0], "worst grad rel", worst)
    state = {k[len("param/"):]: v for k, v in g.items() if k.startswith("param/")}
    lq, _, gq = O.train_step_grads(state, 4, "weatherbert", g["weather"], g["coords"], g["year"], g["interval"], g["mask"],
                                   storage="bf16")
    assert abs(loss.item() - lq["total_loss"]) <= 2e-4 * abs(lq["total_loss"]), (loss.item(), lq)
    _check_grads_bf16_oracle(model, gq, "bert")
    # the public forward (sliced view) gives the same numbers and supports autograd through torch ops
    model.zero_grad()
    out = model(w, c, yr, iv, weather_feature_mask=mask)
    assert out.shape == (8, 365, 31)
    loss2 = nn.functional.mse_loss(w[mask], out[mask])
    loss2.backward()
    assert abs(loss2.item() - loss.item()) <= 1e-5 * abs(loss.item())
    _check_grads(model, ref_grads, "bert/torch-loss")


def test_weatherformer_matches_reference_golden():
    g, model, (w, c, yr, iv, mask), ref_grads = _load_golden("weatherformer_mini_b8.npz", WeatherFormer)
    mask = mask[:, :1, :].expand(-1, 365, -1)  # stride-0 feature mask, as the reference loader produces
    mu, var = model(w, c, yr, iv, weather_feature_mask=mask)
    assert _rel(mu.detach().cpu().numpy(), g["mu"]) <= 1e-2
    assert _rel(var.detach().cpu().numpy(), g["var"]) <= 1e-2
    losses = engine.former_elbo(mu._wm_raw, w, mask, float(g["beta"][0]))
    losses["total_loss"].backward()
    _check_losses({k: v.item() for k, v in losses.items()}, dict(zip(("total_loss", "reconstruction", "kl_term"), g["loss"])))
    worst = _check_grads(model, ref_grads, "former")
    state = {k[len("param/"):]: v for k, v in g.items() if k.startswith("param/")}
    lq, _, gq = O.train_step_grads(state, 4, "weatherformer", g["weather"], g["coords"], g["year"], g["interval"],
                                   g["mask"], beta=float(g["beta"][0]), storage="bf16")
    assert abs(losses["total_loss"].item() - lq["total_loss"]) <= 2e-4 * abs(lq["total_loss"]), (losses, lq)
    _check_grads_bf16_oracle(model, gq, "former")
    print("former mini: losses", {k: v.item() for k, v in losses.items()}, "ref", g["loss"], "worst", worst)


@pytest.mark.parametrize("kind,size,B,S", [("weatherformer", "small", 3, 365), ("weatherbert", "medium", 2, 364),
                                           ("weatherformer", "large", 2, 365)])
def test_model_matches_numpy_oracle(kind, size, B, S):
    torch.manual_seed(7)
    cls = WeatherBERT if kind == "weatherbert" else WeatherFormer
    hp = O.get_model_params(size)
    model = cls(weather_dim=31, output_dim=31, device=torch.device(DEV), **hp).to(DEV).train()
    _neutralise_dropout(model)
    with torch.no_grad():  # de-generate the all-equal layers and zero biases a little
        for p in model.parameters():
            p.add_(torch.randn_like(p) * 0.02)
    state = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    weather, coords, year, interval = O.synthetic_batch(B, S, seed=3)
    rs = np.random.RandomState(5)
    if kind == "weatherbert":
        mask = rs.rand(B, S, 31) < 0.3
    else:
        mask = np.ascontiguousarray(O.weatherformer_mask(rs.rand(B, 31).astype(np.float32), 10, S))
    losses_ref, y_ref, grads_ref = O.train_step_grads(state, hp["num_heads"], kind, weather, coords, year, interval,
                                                      mask, beta=0.5)
    tw, tc, ty, ti = (torch.from_numpy(a).to(DEV) for a in (weather, coords, year, interval))
    tm = torch.from_numpy(mask).to(DEV)
    y_pad = model.forward_raw(tw, tc, ty, ti, tm)
    if kind == "weatherbert":
        losses = {"total_loss": engine.bert_masked_mse(y_pad, tw, tm)}
    else:
        losses = engine.former_elbo(y_pad, tw, tm, 0.5)
    losses["total_loss"].backward()
    assert _rel(y_pad[..., : y_ref.shape[-1]].detach().cpu().numpy(), y_ref) <= 1e-2
    _check_losses({k: v.item() for k, v in losses.items()}, losses_ref)
    worst = _check_grads(model, grads_ref, f"{kind}-{size}")
    print(kind, size, "loss", losses["total_loss"].item(), "oracle", losses_ref["total_loss"], "worst grad", worst)
    lq, _, gq = O.train_step_grads(state, hp["num_heads"], kind, weather, coords, year, interval, mask, beta=0.5,
                                   storage="bf16")
    assert abs(losses["total_loss"].item() - lq["total_loss"]) <= 2e-4 * abs(lq["total_loss"]), (losses, lq)
    _check_grads_bf16_oracle(model, gq, f"{kind}-{size}", per_tensor=1.5e-2, global_tol=7e-3)


def test_fused_adam_training_reduces_loss_and_matches_torch_adam():
    torch.manual_seed(11)
    hp = O.get_model_params("mini")
    model = WeatherFormer(31, 31, torch.device(DEV), **hp).to(DEV).train()
    _neutralise_dropout(model)
    weather, coords, year, interval = (torch.from_numpy(a).to(DEV) for a in O.synthetic_batch(16, 365, seed=1))
    mask = (torch.rand(16, 31, device=DEV) < 0.3).unsqueeze(1).expand(-1, 365, -1)
    opt = FusedAdam(model.parameters(), lr=1e-3, runtime=model.runtime)
    # shadow: torch.optim.Adam on a copy of the parameters fed the SAME gradients
    shadow = [p.detach().clone().requires_grad_(True) for p in model.parameters()]
    ref_opt = torch.optim.Adam(shadow, lr=1e-3)
    first = last = None
    for step in range(6):
        opt.zero_grad()
        mu, var = model(weather, coords, year, interval, weather_feature_mask=mask)
        loss = engine.former_elbo(mu._wm_raw, weather, mask, 0.5)["total_loss"]
        loss.backward()
        for s, p in zip(shadow, model.parameters()):
            s.grad = p.grad.detach().clone()
        opt.step()
        ref_opt.step()
        first = loss.item() if first is None else first
        last = loss.item()
    assert np.isfinite(last) and last < first, (first, last)
    for s, (n, p) in zip(shadow, model.named_parameters()):
        assert torch.allclose(p.detach(), s.detach(), rtol=1e-5, atol=1e-7), f"FusedAdam diverged from torch Adam on {n}"
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 6.0


def test_fused_adam_restarts_like_torch_adam_after_loading_a_state_dict():
    """load_state_dict must reset the flat step counter. Step 5 times, load (a) the empty initial state and
    (b) the state after 2 steps, and compare the next update with torch.optim.Adam doing the same."""
    torch.manual_seed(13)
    hp = O.get_model_params("mini")
    model = WeatherBERT(31, 31, torch.device(DEV), **hp).to(DEV).train()
    _neutralise_dropout(model)
    weather, coords, year, interval = (torch.from_numpy(a).to(DEV) for a in O.synthetic_batch(4, 365, seed=6))
    mask = torch.rand(4, 365, 31, device=DEV) < 0.15
    opt = FusedAdam(model.parameters(), lr=1e-3, runtime=model.runtime)
    shadow = [p.detach().clone().requires_grad_(True) for p in model.parameters()]
    ref = torch.optim.Adam(shadow, lr=1e-3)
    import copy
    states = {0: (copy.deepcopy(opt.state_dict()), copy.deepcopy(ref.state_dict()))}

    def one_step():
        opt.zero_grad()
        engine.bert_masked_mse(model.forward_raw(weather, coords, year, interval, mask), weather, mask).backward()
        for s_, p in zip(shadow, model.parameters()):
            s_.grad = p.grad.detach().clone()
        opt.step()
        ref.step()

    for k in range(5):
        one_step()
        if k == 1:
            states[2] = (copy.deepcopy(opt.state_dict()), copy.deepcopy(ref.state_dict()))
    for at in (0, 2):
        opt.load_state_dict(states[at][0])
        ref.load_state_dict(states[at][1])
        one_step()
        for s_, (n, p) in zip(shadow, model.named_parameters()):
            assert torch.allclose(p.detach(), s_.detach(), rtol=1e-5, atol=1e-7), f"after reload at step {at}: {n}"
        assert float(opt.state_dict()["state"][0]["step"]) == at + 1


def test_dropout_training_mode_runs_and_is_replayable():
    torch.manual_seed(3)
    model = WeatherBERT(31, 31, torch.device(DEV), **O.get_model_params("small")).to(DEV).train()
    weather, coords, year, interval = (torch.from_numpy(a).to(DEV) for a in O.synthetic_batch(4, 365, seed=2))
    mask = torch.rand(4, 365, 31, device=DEV) < 0.15
    rt = model.runtime
    assert abs(rt.dropout_p - 0.1) < 1e-9
    y1 = model.forward_raw(weather, coords, year, interval, mask)
    l1 = engine.bert_masked_mse(y1, weather, mask)
    l1.backward()
    g1 = model.in_proj.weight.grad.clone()
    rt.step_counter -= 1  # replay the same dropout stream
    model.zero_grad()
    y2 = model.forward_raw(weather, coords, year, interval, mask)
    engine.bert_masked_mse(y2, weather, mask).backward()
    assert torch.equal(y1, y2) and torch.equal(g1, model.in_proj.weight.grad)
    y3 = model.forward_raw(weather, coords, year, interval, mask)  # next step: different masks
    assert not torch.equal(y1, y3)
    model.eval()
    with torch.no_grad():
        e1 = model(weather, coords, year, interval, weather_feature_mask=mask)
        e2 = model(weather, coords, year, interval, weather_feature_mask=mask)
    assert torch.equal(e1, e2) and torch.isfinite(e1).all()


def test_checkpoint_roundtrip_and_load_pretrained(tmp_path):
    torch.manual_seed(5)
    hp = O.get_model_params("mini")
    bert = WeatherBERT(31, 31, torch.device(DEV), **hp).to(DEV)
    weather, coords, year, interval = (torch.from_numpy(a).to(DEV) for a in O.synthetic_batch(2, 365, seed=4))
    mask = torch.rand(2, 365, 31, device=DEV) < 0.15
    bert.eval()
    with torch.no_grad():
        ref = bert(weather, coords, year, interval, weather_feature_mask=mask)
    path = tmp_path / "weatherbert_59.7k_latest.pth"
    torch.save(bert, path)  # whole pickled module, as BaseTrainer.save_checkpoint does
    again = torch.load(path, weights_only=False).to(DEV).eval()
    assert type(again).__module__.endswith("pretraining.models.weatherbert")
    with torch.no_grad():
        assert torch.equal(again(weather, coords, year, interval, weather_feature_mask=mask), ref)
    former = WeatherFormer(31, 31, torch.device(DEV), **hp).to(DEV).eval()
    former.load_pretrained(again)  # BERT -> Former keeps Former's own wider head
    assert former.out_proj.out_features == 62
    assert torch.equal(former.in_proj.weight, again.in_proj.weight)
    with torch.no_grad():
        mu, var = former(weather, coords, year, interval, weather_feature_mask=mask)
    assert mu.shape == (2, 365, 31) and (var > 0).all() and (var <= 1).all()
    with pytest.raises(ValueError):
        WeatherBERT(30, 30, torch.device(DEV), **hp).to(DEV).load_pretrained(again)


def test_eval_forward_is_lean_and_matches_the_saving_forward():
    """Validation path (reference: model.eval() + torch.no_grad(), base_trainer.py:262-285): the no-save schedule must
    give bit-identical outputs to the saving one, from a handle whose workspace holds one layer of activations."""
    import ctypes as C

    from weathermodel_b200 import _lib

    torch.manual_seed(21)
    hp = O.get_model_params("small")
    model = WeatherFormer(31, 31, torch.device(DEV), **hp).to(DEV).train()
    _neutralise_dropout(model)
    weather, coords, year, interval = (torch.from_numpy(a).to(DEV) for a in O.synthetic_batch(5, 365, seed=8))
    mask = (torch.rand(5, 31, device=DEV) < 0.3).unsqueeze(1).expand(-1, 365, -1)
    rt = model.runtime
    with torch.no_grad():  # first use of this shape under no_grad: an eval-only handle
        y_eval = model.forward_raw(weather, coords, year, interval, mask)
    assert (5, 365, True) in rt._handles and (5, 365, False) not in rt._handles
    ws_eval = rt.workspace.numel()
    y_train = model.forward_raw(weather, coords, year, interval, mask)  # grad enabled: saving schedule, bigger workspace
    assert y_train.requires_grad and (5, 365, False) in rt._handles
    assert torch.equal(y_eval, y_train.detach())
    cfg_t, cfg_e = rt._config(5, 365, False), rt._config(5, 365, True)
    need_t = _lib.lib().wm_encoder_workspace_bytes(C.byref(cfg_t))
    need_e = _lib.lib().wm_encoder_workspace_bytes(C.byref(cfg_e))
    assert ws_eval >= need_e and need_e < 0.45 * need_t, (need_e, need_t)  # 4 layers: one set of activations instead of four + temporaries
    model.eval()
    with torch.no_grad():  # the training handle exists now: reused with the lean schedule
        mu, var = model(weather, coords, year, interval, weather_feature_mask=mask)
    assert torch.equal(mu, y_train.detach()[..., :31])
    # backward after a lean forward on the same handle must be refused, not silently wrong
    model.train()
    y2 = model.forward_raw(weather, coords, year, interval, mask)
    with torch.no_grad():
        model.forward_raw(weather, coords, year, interval, mask)
    with pytest.raises(RuntimeError):
        engine.former_elbo(y2, weather, mask, 0.5)["total_loss"].backward()


def test_backward_accumulates_without_zero_grad_and_mixes_with_torch_ops():
    """ a second backward without zero_grad() adds to .grad like torch; and a loss that uses the head output
    through the fused kernel AND through torch ops gets both gradient contributions."""
    torch.manual_seed(22)
    hp = O.get_model_params("mini")
    model = WeatherBERT(31, 31, torch.device(DEV), **hp).to(DEV).train()
    _neutralise_dropout(model)
    weather, coords, year, interval = (torch.from_numpy(a).to(DEV) for a in O.synthetic_batch(4, 365, seed=9))
    mask = torch.rand(4, 365, 31, device=DEV) < 0.15

    def fused():
        return engine.bert_masked_mse(model.forward_raw(weather, coords, year, interval, mask), weather, mask)

    fused().backward()
    g1 = {n: p.grad.clone() for n, p in model.named_parameters()}
    fused().backward()  # no zero_grad in between
    for n, p in model.named_parameters():
        assert torch.allclose(p.grad, 2 * g1[n], rtol=1e-6, atol=1e-12), n
    model.zero_grad()
    y = model.forward_raw(weather, coords, year, interval, mask)
    extra = 1e-3 * (y[..., :31] ** 2).mean()
    (engine.bert_masked_mse(y, weather, mask) + extra).backward()
    g_both = {n: p.grad.clone() for n, p in model.named_parameters()}
    model.zero_grad()
    y = model.forward_raw(weather, coords, year, interval, mask)
    (1e-3 * (y[..., :31] ** 2).mean()).backward()
    for n, p in model.named_parameters():
        ref = g1[n] + p.grad
        err = (g_both[n] - ref).norm() / (ref.norm() + 1e-30)
        assert err < 2e-2, (n, err.item())  # two bf16 roundings of dy instead of one
    # a scaled loss: the upstream gradient reaches the kernels as a device scalar
    model.zero_grad()
    (3.0 * fused()).backward()
    for n, p in model.named_parameters():
        err = (p.grad - 3 * g1[n]).norm() / (3 * g1[n]).norm()
        assert err < 1e-2, (n, err.item())  # 3 * dy is rounded to bf16 at different points than dy
