"""Checkpoint interchange with files SAVED BY THE REFERENCE.

tests/golden/ref_weatherbert_mini_latest.pth            whole pickled reference module (class path
                                                         src.pretraining.models.weatherbert.WeatherBERT)
tests/golden/ref_weatherbert_mini_latest_checkpoint.pth  the resume dictionary of BaseTrainer.save_checkpoint
                                                         (src/base_trainer/base_trainer.py:121-170) after 3 Adam steps
tests/golden/ref_weatherbert_mini_eval.npz               eval-mode output of that module on fixed inputs
All three were written by oracle/make_golden.py from the unmodified reference.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from src.pretraining.models.weatherbert import WeatherBERT  # noqa: E402  (the alias package -> weathermodel_b200)
from src.utils.utils import get_model_params, get_scheduler  # noqa: E402
from weathermodel_b200.optim import FusedAdam  # noqa: E402


def _ckpt():
    return torch.load(os.path.join(GOLD, "ref_weatherbert_mini_latest_checkpoint.pth"), map_location="cpu",
                      weights_only=False)


def test_reference_pickle_becomes_our_class():
    m = torch.load(os.path.join(GOLD, "ref_weatherbert_mini_latest.pth"), map_location="cpu", weights_only=False)
    assert isinstance(m, WeatherBERT) and type(m).__module__.startswith("weathermodel_b200")
    want = _ckpt()["model_state_dict"]
    got = m.state_dict()
    assert list(got.keys()) == list(want.keys())
    for k in want:
        assert torch.equal(got[k], want[k]), k
    assert m.total_params() == _ckpt()["output_json"]["model_config"]["total_params"]
    # the unpickled object carries no runtime yet and builds one on demand (flat buffers alias the parameters)
    assert "_wm_runtime" not in m.__dict__
    m.runtime.ensure_flat(torch.device("cpu"))
    for k in want:
        assert torch.equal(m.state_dict()[k], want[k]), k
    # and it is a valid source for load_pretrained (the fine-tuning entry point, weatherbert.py:58-82)
    fresh = WeatherBERT(31, 31, torch.device("cpu"), **get_model_params("mini"))
    fresh.load_pretrained(m)
    assert torch.equal(fresh.in_proj.weight, m.in_proj.weight)


def test_reference_resume_dictionary_loads_and_round_trips():
    ck = _ckpt()
    model = WeatherBERT(31, 31, torch.device("cpu"), **get_model_params("mini"))
    model.load_state_dict(ck["model_state_dict"])
    opt = FusedAdam(model.parameters(), lr=5e-4, runtime=model.runtime, allow_host_params=True)
    sch = get_scheduler(opt, 2, 10, 0.99)  # same order as BaseTrainer: build both, then load_checkpoint (:390-409)
    opt.load_state_dict(ck["optimizer_state_dict"])
    sch.load_state_dict(ck["scheduler_state_dict"])
    g = np.load(os.path.join(GOLD, "ref_weatherbert_mini_eval.npz"))
    assert abs(opt.param_groups[0]["lr"] - float(g["lr"][0])) < 1e-12
    assert sch.last_epoch == ck["scheduler_state_dict"]["last_epoch"] == 3
    # what we would save back is what the reference's torch.optim.Adam wrote: same structure, same numbers
    back = opt.state_dict()
    ref = ck["optimizer_state_dict"]
    assert back["param_groups"][0]["params"] == ref["param_groups"][0]["params"]
    for key in ("lr", "betas", "eps", "weight_decay"):
        assert back["param_groups"][0][key] == ref["param_groups"][0][key], key
    assert back["state"].keys() == ref["state"].keys()
    for i, st in ref["state"].items():
        assert float(back["state"][i]["step"]) == float(st["step"]) == 3.0
        assert torch.equal(back["state"][i]["exp_avg"], st["exp_avg"])
        assert torch.equal(back["state"][i]["exp_avg_sq"], st["exp_avg_sq"])


@pytest.mark.gpu
def test_reference_checkpoint_runs_on_the_gpu_and_continues_like_torch_adam():
    g = np.load(os.path.join(GOLD, "ref_weatherbert_mini_eval.npz"))
    m = torch.load(os.path.join(GOLD, "ref_weatherbert_mini_latest.pth"), map_location="cpu", weights_only=False)
    m = m.to("cuda").eval()
    t = lambda k: torch.from_numpy(g[k]).to("cuda")  # noqa: E731
    with torch.no_grad():
        y = m(t("weather"), t("coords"), t("year"), t("interval"), weather_feature_mask=t("mask"))
    ref = g["y"].astype(np.float64)
    rel = np.linalg.norm(y.float().cpu().numpy().astype(np.float64) - ref) / np.linalg.norm(ref)
    assert rel <= 1e-2, rel  # bf16 activations against the fp32 reference output
    # resume: one more optimiser step from the loaded moments == torch.optim.Adam from the same state
    ck = _ckpt()
    m.train()
    opt = FusedAdam(m.parameters(), lr=5e-4, runtime=m.runtime)
    opt.load_state_dict(ck["optimizer_state_dict"])
    twin = WeatherBERT(31, 31, torch.device("cpu"), **get_model_params("mini"))
    twin.load_state_dict(ck["model_state_dict"])
    topt = torch.optim.Adam(twin.parameters(), lr=5e-4)
    topt.load_state_dict(ck["optimizer_state_dict"])
    gen = torch.Generator().manual_seed(3)
    m.runtime.ensure_flat(torch.device("cuda"))
    for (_, p), (_, q) in zip(m.named_parameters(), twin.named_parameters()):
        grad = torch.randn(q.shape, generator=gen) * 1e-2
        q.grad = grad.clone()
        p.grad.copy_(grad) if p.grad is not None else None
    if any(p.grad is None for p in m.parameters()):  # gradients must be the flat-buffer views
        m.runtime.publish_grads()
        gen = torch.Generator().manual_seed(3)
        for (_, p), (_, q) in zip(m.named_parameters(), twin.named_parameters()):
            p.grad.copy_(torch.randn(q.shape, generator=gen) * 1e-2)
    opt.step()
    topt.step()
    for (n, p), (_, q) in zip(m.named_parameters(), twin.named_parameters()):
        assert torch.allclose(p.detach().cpu(), q.detach(), rtol=1e-5, atol=1e-7), n
        assert float(opt.state[p]["step"]) == 4.0
