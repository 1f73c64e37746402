"""Recorded training steps (weathermodel_b200/graph_step.py): the body of BaseTrainer's step
(reference src/base_trainer/base_trainer.py:239-252) replayed as one CUDA graph must train exactly like the eager step."""
import copy
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
from weathermodel_b200 import engine, ops  # noqa: E402
from weathermodel_b200.graph_step import CapturedTrainStep  # noqa: E402
from weathermodel_b200.optim import FusedAdam  # noqa: E402

DEV = "cuda"


def _neutralise(model):
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0


def _batches(n, B, seed):
    out = []
    for i in range(n):
        w, c, y, iv = (torch.from_numpy(a).to(DEV) for a in O.synthetic_batch(B, 365, seed=seed + i))
        # dense, like the recorded step's static input buffer (the loss kernels sum a stride-0 feature mask in another order)
        mask = (torch.rand(B, 31, device=DEV) < 0.3).unsqueeze(1).expand(-1, 365, -1).contiguous()
        out.append((w, c, y, iv, mask))
    return out


def test_recorded_step_trains_exactly_like_the_eager_step():
    torch.manual_seed(31)
    model = WeatherFormer(31, 31, torch.device(DEV), **O.get_model_params("mini")).to(DEV).train()
    _neutralise(model)
    opt = FusedAdam(model.parameters(), lr=1e-3, runtime=model.runtime)
    batches = _batches(6, 8, seed=40)

    def loss_fn(w, c, y, iv, mask):
        return engine.former_elbo(model.forward_raw(w, c, y, iv, mask), w, mask, 0.5)

    def eager(batch):
        opt.zero_grad()
        losses = loss_fn(*batch)
        losses["total_loss"].backward()
        opt.step()
        return {k: v.item() for k, v in losses.items()}

    for b in batches[:2]:
        eager(b)
    snap_p = model.runtime.flat_params.clone()
    snap_m, snap_v, snap_step = opt._flat_m.clone(), opt._flat_v.clone(), opt._step
    ref_losses = [eager(b) for b in batches[2:]]
    ref_params = model.runtime.flat_params.clone()
    # rewind and do the same four steps as one recording + three replays
    model.runtime.flat_params.copy_(snap_p)
    opt._flat_m.copy_(snap_m)
    opt._flat_v.copy_(snap_v)
    opt._step = snap_step
    model.runtime.mark_weights_dirty()
    launches0 = ops.lib().wm_launch_count()
    step = CapturedTrainStep(opt, loss_fn, batches[2])
    got = [{k: v.item() for k, v in step.first_losses.items()}]
    recorded = ops.lib().wm_launch_count() - launches0
    for b in batches[3:]:
        got.append({k: v.item() for k, v in step(*b).items()})
    assert ops.lib().wm_launch_count() - launches0 == recorded  # replays launch nothing through the library
    assert opt._step == snap_step + 4 and float(opt.state_dict()["state"][0]["step"]) == snap_step + 4
    print("eager   :", [r["total_loss"] for r in ref_losses])
    print("recorded:", [g["total_loss"] for g in got])
    for g, r in zip(got, ref_losses):
        for k in r:
            assert abs(g[k] - r[k]) <= 1e-6 * max(1.0, abs(r[k])), (k, got, ref_losses)
    assert torch.allclose(model.runtime.flat_params, ref_params, rtol=1e-6, atol=1e-9)
    # an eager forward after replays sees the updated weights (shadows are marked dirty)
    model.eval()
    with torch.no_grad():
        mu, _ = model(*batches[0][:4], weather_feature_mask=batches[0][4])
    assert torch.isfinite(mu).all()
    assert ops.device_error() == 0


def test_recorded_step_draws_fresh_dropout_masks_and_follows_the_lr():
    torch.manual_seed(32)
    model = WeatherBERT(31, 31, torch.device(DEV), **O.get_model_params("small")).to(DEV).train()
    opt = FusedAdam(model.parameters(), lr=0.0, runtime=model.runtime)  # lr 0: parameters stay put, only masks change
    w, c, y, iv = (torch.from_numpy(a).to(DEV) for a in O.synthetic_batch(4, 365, seed=50))
    mask = torch.rand(4, 365, 31, device=DEV) < 0.15

    def loss_fn(w, c, y, iv, mask):
        return {"total_loss": engine.bert_masked_mse(model.forward_raw(w, c, y, iv, mask), w, mask)}

    for _ in range(2):
        opt.zero_grad()
        loss_fn(w, c, y, iv, mask)["total_loss"].backward()
        opt.step()
    p0 = model.runtime.flat_params.clone()
    step = CapturedTrainStep(opt, loss_fn, (w, c, y, iv, mask))
    losses = [step.first_losses["total_loss"].item()] + [step(w, c, y, iv, mask)["total_loss"].item() for _ in range(5)]
    assert torch.equal(model.runtime.flat_params, p0)                      # lr = 0 was honoured by the replayed Adam
    assert len({round(v, 7) for v in losses}) == len(losses), losses      # six replays, six different dropout draws
    assert np.std(losses) < 0.05 * np.mean(losses)                         # ... of the same model on the same batch
    opt.param_groups[0]["lr"] = 1e-3                                       # the scheduler changes lr between epochs
    before = step(w, c, y, iv, mask)["total_loss"].item()
    assert not torch.equal(model.runtime.flat_params, p0)
    for _ in range(30):
        last = step(w, c, y, iv, mask)["total_loss"].item()
    assert last < before, (before, last)
    assert ops.device_error() == 0


def test_trainer_uses_recorded_steps_for_small_shapes(tmp_path, monkeypatch):
    """BaseTrainer._captured_step: two eager batches, then a recording, then replays -- same result as an eager trainer."""
    from src.pretraining.trainers.weatherbert_trainer import WeatherBertTrainer

    monkeypatch.chdir(tmp_path)
    batches = [(b[0], b[1], b[2], b[3], torch.rand(8, 365, 31, device=DEV) < 0.15) for b in _batches(7, 8, seed=60)]

    def run(mode):
        monkeypatch.setenv("WM_CUDA_GRAPH", mode)
        torch.manual_seed(33)
        model = WeatherBERT(31, 31, torch.device(DEV), **O.get_model_params("mini")).to(DEV)
        _neutralise(model)
        tr = WeatherBertTrainer(model, masking_prob=0.15, n_masked_features=1, batch_size=8, num_epochs=3, init_lr=1e-3,
                                num_warmup_epochs=0, decay_factor=0.99)
        tr.current_epoch = 0
        tr._run_epoch(batches, "train", training=True)
        return tr, model.runtime.flat_params.clone()

    tr_e, p_e = run("0")
    assert not hasattr(tr_e, "_graph_steps")
    tr_g, p_g = run("1")
    assert len(tr_g._graph_steps) == 1 and not getattr(tr_g, "_graph_disabled", False)
    assert torch.allclose(p_g, p_e, rtol=1e-6, atol=1e-9)
    assert ops.device_error() == 0


def test_a_failed_recording_leaves_the_process_usable(monkeypatch):
    """If recording a step fails, training must go on eagerly and the default CUDA generator must still work (torch
    leaves it flagged as capturing when capture_end throws; graph_step.repair_default_generator swaps the state)."""
    from weathermodel_b200.graph_step import repair_default_generator

    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    x = torch.zeros(8, device=DEV)
    torch.cuda.synchronize()
    with pytest.raises(Exception):
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            torch.rand(4, device=DEV)          # registers the generator with the capture
            ev = torch.cuda.Event()
            with torch.cuda.stream(s):
                x.add_(1)                       # uncaptured work on another stream ...
                ev.record(s)
            torch.cuda.current_stream().wait_event(ev)  # ... that the captured stream now depends on: invalidates the capture
            x.mul_(2)
    torch.manual_seed(5)
    want_offset = torch.cuda.default_generators[0].get_offset()
    repair_default_generator(torch.device(DEV))
    assert torch.cuda.default_generators[0].get_offset() == want_offset
    a = torch.rand(16, device=DEV)              # would raise "Offset increment outside graph capture ..." without the repair
    torch.manual_seed(5)
    assert torch.equal(a, torch.rand(16, device=DEV))
    assert torch.isfinite(torch.randperm(10, device=DEV).float()).all()
