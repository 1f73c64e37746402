"""Encoder siblings: WeatherAutoencoder, SimMTM, WeatherFormerSinusoid, WeatherFormerMixture.
They inherit the B200 encoder; what is new is checked here: the SimMTM segment mask (bit-exact against the
reference's CPU stream), the learned priors (same parameter draws for a seed, loss terms and gradients against
tests/golden/weatherformer_{sinusoid,mixture}_mini_b4.npz written by the unmodified reference), and the CLI."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from src.pretraining.dataloader.pretraining_dataloader import StreamingDataset  # noqa: E402
from src.pretraining.models.simmtm import SimMTM  # noqa: E402
from src.pretraining.models.weatherautoencoder import WeatherAutoencoder  # noqa: E402
from src.pretraining.models.weatherbert import WeatherBERT  # noqa: E402
from src.pretraining.models.weatherformer import WeatherFormer  # noqa: E402
from src.pretraining.models.weatherformer_mixture import WeatherFormerMixture  # noqa: E402
from src.pretraining.models.weatherformer_sinusoid import WeatherFormerSinusoid  # noqa: E402
from src.utils.losses import (compute_gaussian_kl_divergence, compute_mixture_kl_divergence,  # noqa: E402
                              gaussian_log_likelihood)
from src.utils.utils import get_model_params  # noqa: E402


def _build(kind, device="cpu"):
    torch.manual_seed(1234)
    hp = get_model_params("mini")
    if kind == "sinusoid":
        return WeatherFormerSinusoid(weather_dim=31, output_dim=31, k=4, device=torch.device(device), **hp)
    return WeatherFormerMixture(weather_dim=31, output_dim=31, k=7, device=torch.device(device), **hp)


@pytest.mark.parametrize("name,p", [("simmtm_p15", 0.15), ("simmtm_p30", 0.30)])
def test_simmtm_mask_matches_reference_stream(name, p):
    g = np.load(os.path.join(GOLD, "masks_cpu.npz"))
    ds = StreamingDataset([], masking_function="simmtm", masking_prob=p, n_masked_features=1)
    ds.device = "cpu"
    torch.manual_seed(1234)
    m = ds.masking_function(365, 31, 64)
    assert m.shape == (64, 365, 31) and m.dtype == torch.bool
    m = m.contiguous().numpy()
    assert (m == m[:, :, :1]).all()  # the same time steps for every feature
    assert int(m.sum()) == int(g[name + "_sum"][0])
    assert np.array_equal(np.packbits(m[:, :, 0]), g[name + "_packed"])
    assert m[:, :, 0].sum(1).max() <= int(365 * p)
    assert torch.equal(StreamingDataset([], masking_function="simmtm", masking_prob=0.0).masking_function(365, 31, 2),
                       torch.zeros(2, 365, 31, dtype=torch.bool))


def test_sibling_classes_keep_the_reference_hierarchy_and_draws():
    hp = get_model_params("mini")
    ae = WeatherAutoencoder(31, 31, torch.device("cpu"), **hp)
    sm = SimMTM(31, 31, torch.device("cpu"), **hp)
    assert isinstance(ae, WeatherBERT) and ae.name == "weatherautoencoder"
    assert isinstance(sm, WeatherBERT) and sm.name == "simmtm"
    for kind in ("sinusoid", "mixture"):
        g = np.load(os.path.join(GOLD, f"weatherformer_{kind}_mini_b4.npz"))
        m = _build(kind)
        assert isinstance(m, WeatherFormer) and m.name == f"weatherformer_{kind}"
        names = [k for k, _ in m.named_parameters()]
        assert sorted("psum/" + n for n in names) == sorted(k for k in g.files if k.startswith("psum/"))
        for n, prm in m.named_parameters():  # same seed, same draws as the reference, prior parameters included
            want = g["psum/" + n]
            got = [prm.detach().double().sum().item(), prm.detach().double().abs().sum().item()]
            assert np.allclose(got, want, rtol=1e-12, atol=1e-12), n
    with pytest.raises(ValueError):
        _build("mixture").load_pretrained(WeatherFormerMixture(31, 31, torch.device("cpu"), k=3, **hp))


def test_mixture_kl_formula():
    g = torch.Generator().manual_seed(0)
    B, k, S, F = 3, 4, 5, 6
    z, mu = torch.randn(B, S, F, generator=g), torch.randn(B, S, F, generator=g)
    var = torch.rand(B, S, F, generator=g) + 0.1
    mu_k, var_k = torch.randn(B, k, S, F, generator=g), torch.rand(B, k, S, F, generator=g) + 0.1
    log_w = torch.log_softmax(torch.randn(B, k, generator=g), dim=1)
    mask = torch.rand(B, S, F, generator=g) < 0.5
    got = compute_mixture_kl_divergence(z, mask, mu, var, mu_k, var_k, log_w)
    for b in range(B):
        lq = sum(-0.5 * np.log(2 * np.pi * var[b, s, f].item()) - 0.5 * (z[b, s, f] - mu[b, s, f]).item() ** 2 / var[b, s, f].item()
                 for s in range(S) for f in range(F) if mask[b, s, f])
        comps = []
        for i in range(k):
            lp = sum(-0.5 * np.log(2 * np.pi * var_k[b, i, s, f].item())
                     - 0.5 * (z[b, s, f] - mu_k[b, i, s, f]).item() ** 2 / var_k[b, i, s, f].item()
                     for s in range(S) for f in range(F) if mask[b, s, f])
            comps.append(log_w[b, i].item() + lp)
        mx = max(comps)
        want = lq - (mx + np.log(sum(np.exp(c - mx) for c in comps)))
        assert abs(got[b].item() - want) < 1e-3 * max(1.0, abs(want))


def test_cli_rejects_models_outside_the_encoder_family(tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT)
    res = subprocess.run([sys.executable, "-m", "src.pretraining.pretraining_main", "--model", "mlp"], cwd=tmp_path, env=env,
                         capture_output=True, text=True, timeout=300)
    assert res.returncode != 0 and "outside the B200 hot path" in (res.stderr + res.stdout)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["sinusoid", "mixture"])
def test_learned_prior_models_match_reference_golden(kind, monkeypatch):
    g = dict(np.load(os.path.join(GOLD, f"weatherformer_{kind}_mini_b4.npz")))
    model = _build(kind).to("cuda").train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    t = lambda k: torch.from_numpy(g[k]).to("cuda")  # noqa: E731
    w, c, yr, iv = t("weather"), t("coords"), t("year"), t("interval")
    mask = t("mask").unsqueeze(1).expand(-1, 365, -1)
    beta = float(g["beta"][0])
    out = model(w, c, yr, iv, weather_feature_mask=mask)
    mu, var = out[0], out[1]
    n_bar = mask.sum(dim=(1, 2)).float().mean()
    recon = (-gaussian_log_likelihood(w, mu, var, mask) / n_bar).mean()
    if kind == "sinusoid":
        kl_raw = compute_gaussian_kl_divergence(mask, mu, var, out[2], out[3])
    else:
        z = mu + torch.sqrt(var) * t("epsilon")
        kl_raw = compute_mixture_kl_divergence(z=z, feature_mask=mask, mu_x=mu, var_x=var, mu_k=out[2], var_k=out[3],
                                               log_w_k=out[4])
    kl = beta * kl_raw.mean() / n_bar
    loss = recon + kl
    loss.backward()
    ref = g["loss"]
    total = abs(ref[0])
    assert abs(loss.item() - ref[0]) <= 1e-3 * total, (loss.item(), ref[0])
    assert abs(recon.item() - ref[1]) <= 1e-3 * total and abs(kl.item() - ref[2]) <= 1e-3 * total
    gtot = np.sqrt(sum(float(v[0]) ** 2 for k, v in g.items() if k.startswith("gnorm/")))
    for n, prm in model.named_parameters():
        assert prm.grad is not None, n
        got = prm.grad.detach().float().cpu().numpy().astype(np.float64)
        want_norm = float(g["gnorm/" + n][0])
        if want_norm < 1e-6 * gtot:
            assert np.linalg.norm(got) < 1e-5 * gtot, n
            continue
        assert abs(np.linalg.norm(got) - want_norm) <= 1e-2 * want_norm, (n, np.linalg.norm(got), want_norm)
        sample = got.reshape(-1)[::max(1, got.size // 2048)]
        want = g["gsample/" + n].astype(np.float64)
        rel = np.linalg.norm(sample - want) / (np.linalg.norm(want) + 1e-30)
        assert rel <= 5e-2, (n, rel)


def _write_chunks(base, ids, n, seed=0):
    gen = torch.Generator().manual_seed(seed)
    os.makedirs(base, exist_ok=True)
    for cid in ids:
        w = torch.randn(n, 365, 31, generator=gen)
        coords = torch.stack([torch.rand(n, generator=gen) * 120 - 60, torch.rand(n, generator=gen) * 360 - 180], 1)
        index = torch.stack([torch.randint(0, 2, (n,), generator=gen).float(), torch.full((n,), 7.0)], 1)
        torch.save(torch.utils.data.TensorDataset(w, coords, index), os.path.join(base, f"weather_dataset_weekly_{cid}.pt"))


@pytest.mark.gpu
@pytest.mark.parametrize("model,extra,keys", [
    ("weatherautoencoder", ["--n-masked-features", "10"], {"total_loss"}),
    ("simmtm", ["--masking-prob", "0.15"], {"total_loss"}),
    ("weatherformersinusoid", ["--n-masked-features", "10", "--n-mixture-components", "4"], {"total_loss", "reconstruction", "kl_term"}),
    ("weatherformermixture", ["--n-masked-features", "10", "--n-mixture-components", "3"], {"total_loss", "reconstruction", "kl_term"}),
])
def test_sibling_cli_dry_run(tmp_path, model, extra, keys):
    base = tmp_path / "data" / "nasa_power" / "processed"
    _write_chunks(str(base), [1, 34, 53, 72, 81, 7, 30, 56, 59], n=96)
    env = dict(os.environ, DRY_RUN="1", PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "src.pretraining.pretraining_main", "--model", model, "--model-size", "mini",
           "--batch-size", "64", "--n-epochs", "3", "--n-warmup-epochs", "1", "--init-lr", "0.001"] + extra
    res = subprocess.run(cmd, cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    out_dir = tmp_path / "data" / "trained_models" / "pretraining"
    outs = [f for f in os.listdir(out_dir) if f.endswith("_output.json")]
    assert len(outs) == 1, sorted(os.listdir(out_dir))
    js = json.load(open(out_dir / outs[0]))
    assert set(js["losses"]["train"]) == keys
    tr = js["losses"]["train"]["total_loss"]
    assert len(tr) == 3 and all(v == v for v in tr) and tr[2] < tr[0], tr
