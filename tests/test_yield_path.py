"""Yield fine-tune path (BASELINE config 5): dataset semantics on CPU, end-to-end CLI on a B200."""
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOIL = ["bdod", "cec", "cfvo", "clay", "nitrogen", "ocd", "ocs", "phh2o", "sand", "silt", "soc"]
DEPTHS = ["0-5cm", "5-15cm", "15-30cm", "30-60cm", "60-100cm", "100-200cm"]


def synthetic_yield_csv(path, n_counties=12, years=range(1995, 2019), seed=0):
    """CSV in the khaki_multi_crop_yield.csv layout: one row per (county, year)."""
    rs = np.random.RandomState(seed)
    rows = []
    for loc in range(n_counties):
        lat, lng = rs.uniform(30, 48), rs.uniform(-100, -80)
        for y in years:
            r = {"loc_ID": loc, "year": y, "State": "S", "County": f"C{loc}", "lat": lat, "lng": lng,
                 "soybean_yield": 45 + 8 * rs.randn()}
            w = rs.randn(6, 52) * 3 + 10
            for i in range(6):
                for j in range(52):
                    r[f"W_{i + 1}_{j + 1}"] = w[i, j]
            for i in range(14):
                r[f"P_{i + 1}"] = rs.rand()
            for m in SOIL:
                for d in DEPTHS:
                    r[f"{m}_mean_{d}"] = rs.rand()
            rows.append(r)
    df = pd.DataFrame(rows)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    df.to_csv(path, index=False)
    return df


def test_crop_dataset_sample_semantics(tmp_path):
    from src.crop_yield.dataloader.yield_dataloader import CropDataset, WEATHER_INDICES, get_train_test_loaders

    df = synthetic_yield_csv(str(tmp_path / "d.csv"), n_counties=4, years=range(2000, 2012))
    df = df.sort_values(["loc_ID", "year"])
    n_past = 3
    ds = CropDataset(df.copy(), start_year=2005, test_year=2010, test_dataset=False, n_past_years=n_past)
    # literal restatement of the reference sample construction for every candidate
    want = df[(df["year"] >= 2005) & (df["year"] < 2010)]
    assert len(ds) == len(want) == 4 * 5
    for k, (_, row) in enumerate(want.iterrows()):
        q = df[(df["year"] <= row["year"]) & (df["loc_ID"] == row["loc_ID"])].tail(n_past + 1)
        wcols = [f"W_{i}_{j}" for i in range(1, 7) for j in range(1, 53)]
        weather = q[wcols].values.astype("float32").reshape(-1, 6, 52).transpose(0, 2, 1).reshape(-1, 6)
        padded, coord, year, interval, mask, practices, soil, y_past, y = ds[k]
        assert padded.shape == (4 * 52, 31) and mask.shape == (4 * 52, 31)
        assert np.array_equal(padded[:, WEATHER_INDICES].numpy(), weather)
        rest = [c for c in range(31) if c not in WEATHER_INDICES]
        assert (padded[:, rest] == 0).all() and mask[:, rest].all() and not mask[:, WEATHER_INDICES].any()
        exp_year = (torch.tensor(q["year"].values.astype("float32")).unsqueeze(1)
                    + torch.arange(1, 53, dtype=torch.float32).unsqueeze(0) / 52).reshape(-1)
        assert torch.equal(year, exp_year) and interval.item() == 7.0
        yy = q["soybean_yield"].values.astype("float32")
        assert y.item() == yy[-1] and y_past[-1] == yy[-2] and np.array_equal(y_past[:-1], yy[:-1])
        assert np.allclose(coord.numpy(), q[["lat", "lng"]].values.astype("float32")[0])
    with pytest.raises(ValueError):
        CropDataset(df.copy(), 2005, 2010, n_past_years=7)  # 8 * 52 > 365
    tr, te = get_train_test_loaders(df, n_train_years=4, test_year=2010, n_past_years=3, batch_size=5, shuffle=False,
                                    num_workers=0, crop_type="soybean", country="usa")
    b = next(iter(tr))
    assert len(b) == 9 and b[0].shape == (5, 208, 31) and b[4].dtype == torch.bool and b[8].shape == (5, 1)
    assert len(te.dataset) == 4


def test_yield_models_build_and_share_the_encoder_classes():
    from src.crop_yield.models.weatherbert_yield_model import WeatherBERTYieldModel
    from src.crop_yield.models.weatherformer_yield_model import WeatherFormerYieldModel
    from src.pretraining.models.weatherformer import WeatherFormer

    hp = dict(num_heads=4, num_layers=2, hidden_dim_factor=12)
    m = WeatherFormerYieldModel("weatherformer_soybean_yield", torch.device("cpu"), 31, 6, **hp)
    assert isinstance(m.weather_model, WeatherFormer) and m.yield_mlp[0].in_features == 31 + 6 + 1
    heads = 31 * 16 + 16 + 16 + 1 + 38 * 120 + 120 + 120 + 1
    assert m.total_params() == 61262 + heads
    b = WeatherBERTYieldModel("weatherbert_soybean_yield", torch.device("cpu"), 31, 6, **hp)
    b.freeze_weather_model()
    assert not any(p.requires_grad for p in b.weather_model.parameters()) and b.yield_mlp[0].weight.requires_grad
    b.unfreeze_weather_model()
    assert all(p.requires_grad for p in b.weather_model.parameters())
    with pytest.raises(ValueError):
        b.load_pretrained(torch.nn.Linear(2, 2))


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["weatherformer", "weatherbert"])
def test_yield_cli_single_fold_on_gpu(tmp_path, model):
    synthetic_yield_csv(str(tmp_path / "data" / "khaki_soybeans" / "khaki_multi_crop_yield.csv"))
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "src.crop_yield.yield_main", "--model", model, "--model-size", "mini",
           "--batch-size", "16", "--n-past-years", "6", "--n-train-years", "8", "--n-epochs", "4",
           "--n-warmup-epochs", "1", "--test-year", "2016", "--init-lr", "0.002", "--beta", "0.0001"]
    res = subprocess.run(cmd, cwd=tmp_path, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-4000:]
    assert "Final average best RMSE for soybean" in res.stderr + res.stdout
    out_dir = tmp_path / "data" / "trained_models" / "crop_yield"
    files = sorted(os.listdir(out_dir))
    assert any(f.endswith("_best.pth") for f in files) and any(f.endswith("_output.json") for f in files), files


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["weatherbert", "weatherformer"])
def test_yield_model_matches_reference_golden(kind, monkeypatch):
    """Prediction, loss and every gradient of the reference's yield models (tests/golden/*_yield_mini_b6.npz,
    written by oracle/make_golden.py from the unmodified reference) on the loader's input shape: 312 weekly
    steps, 25 of 31 features masked and imputed. Tolerance: 1e-3 relative on the loss, 3e-2 per-tensor
    Frobenius on gradients (bf16 activations), 5e-3 on gradient norms."""
    from src.crop_yield.models.weatherbert_yield_model import WeatherBERTYieldModel
    from src.crop_yield.models.weatherformer_yield_model import WeatherFormerYieldModel
    from src.utils.losses import compute_gaussian_kl_divergence
    from src.utils.utils import get_model_params

    g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"{kind}_yield_mini_b6.npz")))
    cls = WeatherBERTYieldModel if kind == "weatherbert" else WeatherFormerYieldModel
    torch.manual_seed(1234)
    model = cls(name="y", device=torch.device("cpu"), weather_dim=31, n_past_years=5, **get_model_params("mini"))
    state = {k[len("param/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param/")}
    for k, v in model.state_dict().items():
        assert torch.equal(v.cpu(), state[k]), f"initial weight {k} differs from the reference's for seed 1234"
    model = model.to("cuda").train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    t = lambda k: torch.from_numpy(g[k]).to("cuda")  # noqa: E731
    w, c, yr, iv, mask, y_past, target = (t(k) for k in ("weather", "coords", "year", "interval", "mask", "y_past", "target"))
    if kind == "weatherformer":
        eps = t("epsilon")
        monkeypatch.setattr(torch, "randn_like", lambda x, *a, **k: eps)
    out = model(w, c, yr, iv, mask, y_past)
    if kind == "weatherbert":
        pred = out
        loss = torch.nn.functional.mse_loss(pred, target)
    else:
        pred, z, mu, var = out
        assert _rel(mu.detach().cpu().numpy(), g["mu"]) <= 1e-2
        assert _rel(var.detach().cpu().numpy(), g["var"]) <= 1e-2
        kl = compute_gaussian_kl_divergence(mask, mu, var, torch.zeros_like(mu), torch.ones_like(var)).mean()
        assert abs(kl.item() - g["kl"][0]) <= 2e-3 * abs(g["kl"][0]), (kl.item(), g["kl"][0])
        loss = torch.nn.functional.mse_loss(pred, target) + 1e-4 * kl
    loss.backward()
    assert _rel(pred.detach().cpu().numpy(), g["pred"]) <= 5e-3, (pred.flatten(), g["pred"].flatten())
    assert abs(loss.item() - g["loss"][0]) <= 2e-3 * abs(g["loss"][0]), (loss.item(), g["loss"][0])
    worst = (0.0, "")
    gnorm = np.sqrt(sum(float(np.linalg.norm(v.astype(np.float64))) ** 2 for k, v in g.items() if k.startswith("grad/")))
    for name, p in model.named_parameters():
        r = g["grad/" + name].astype(np.float64)
        assert p.grad is not None, name
        got = p.grad.detach().float().cpu().numpy().astype(np.float64)
        nr = np.linalg.norm(r)
        if nr < 1e-6 * gnorm:  # mathematically zero (the bias in front of the pooling softmax): only rounding noise
            assert np.linalg.norm(got) < 1e-5 * gnorm, name
            continue
        assert abs(np.linalg.norm(got) - nr) <= 1e-2 * nr, name
        rel = _rel(got, r)
        worst = max(worst, (rel, name))
        assert rel <= 3e-2, f"{kind} yield: grad {name} rel err {rel:.4g}"
    print(kind, "yield: loss", loss.item(), "ref", g["loss"][0], "worst grad", worst)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["weatherbert", "weatherformer"])
def test_fused_yield_head_matches_the_torch_op_head(kind, monkeypatch):
    """wm_yield_head_fwd / _bwd (one kernel each) against the reference's own torch-op head
    (weatherbert_yield_model.py:40-67, weatherformer_yield_model.py:58-60) on the same encoder output: prediction,
    z, every head-parameter gradient and the gradient handed back to the encoder (seen through in_proj.weight.grad),
    fp32 on both sides -> 1e-5; twice the same result bit for bit (deterministic reductions)."""
    from src.crop_yield.models.weatherbert_yield_model import WeatherBERTYieldModel
    from src.crop_yield.models.weatherformer_yield_model import WeatherFormerYieldModel
    from src.utils.utils import get_model_params

    cls = WeatherBERTYieldModel if kind == "weatherbert" else WeatherFormerYieldModel
    torch.manual_seed(5)
    model = cls(name="y", device=torch.device("cuda"), weather_dim=31, n_past_years=6, **get_model_params("mini")).to("cuda").train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    with torch.no_grad():
        for p in model.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    B, S = 7, 364
    g = torch.Generator(device="cuda").manual_seed(1)
    w = torch.randn(B, S, 31, device="cuda", generator=g)
    c = torch.rand(B, 2, device="cuda", generator=g) * 60
    yr = 1990 + torch.arange(S, device="cuda").float()[None].repeat(B, 1) / 52
    iv = torch.full((B, 1), 7.0, device="cuda")
    mask = torch.ones(B, S, 31, dtype=torch.bool, device="cuda")
    mask[:, :, [7, 8, 11, 1, 2, 29]] = False  # the yield loader's six observed features
    y_past = torch.randn(B, 7, device="cuda", generator=g)
    target = torch.randn(B, 1, device="cuda", generator=g)
    eps = torch.randn(B, S, 31, device="cuda", generator=g)
    monkeypatch.setattr(torch, "randn_like", lambda x, *a, **k: eps)

    def run(fused):
        monkeypatch.setattr(cls, "_fused_head_ok", (lambda self, *_: True) if fused else (lambda self, *_: False))
        model.zero_grad()
        out = model(w, c, yr, iv, mask, y_past)
        pred = out if kind == "weatherbert" else out[0]
        loss = torch.nn.functional.mse_loss(pred, target)
        if kind == "weatherformer":
            loss = loss + 1e-3 * (out[2] ** 2 + out[3]).mean()  # touches mu and var through torch ops as the KL term does
        loss.backward()
        return (pred.detach().clone(), None if kind == "weatherbert" else out[1].detach().clone(),
                {n: p.grad.detach().clone() for n, p in model.named_parameters()})

    pred_t, z_t, g_t = run(False)
    pred_f, z_f, g_f = run(True)
    pred_f2, _, g_f2 = run(True)
    assert torch.equal(pred_f, pred_f2) and all(torch.equal(g_f[n], g_f2[n]) for n in g_f)
    assert torch.allclose(pred_f, pred_t, rtol=1e-5, atol=1e-6), (pred_f.flatten(), pred_t.flatten())
    if z_t is not None:
        assert torch.allclose(z_f, z_t, rtol=1e-5, atol=1e-6)
    for n in g_t:
        ref = g_t[n]
        err = (g_f[n] - ref).norm() / (ref.norm() + 1e-20)
        if "weather_model" in n:
            assert err < 2e-2, (n, err.item())  # encoder gradients pass through bf16 kernels on both paths
        elif ref.norm() > 1e-7:
            assert err < 2e-5, (n, err.item())
        else:
            assert g_f[n].norm() < 1e-6, n  # the bias in front of the softmax: mathematically zero
