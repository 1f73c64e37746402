"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference (torch CPU, fp32).

Run in the build container only (the GPU box has no /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py
Recipe follows SURVEY.md 8(c): model built right after torch.manual_seed(1234); inputs from
torch.Generator().manual_seed(0); mask from the reference's own masking function right after a fresh
torch.manual_seed(1234); dropout neutralised without leaving train() mode. The script asserts the
known-answer values recorded in SURVEY.md 8(c) before writing anything.
"""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
os.chdir("/tmp")

from src.pretraining.dataloader.pretraining_dataloader import StreamingDataset  # noqa: E402
from src.pretraining.models.weatherbert import WeatherBERT  # noqa: E402
from src.pretraining.models.weatherformer import WeatherFormer  # noqa: E402
from src.utils.losses import compute_gaussian_kl_divergence, gaussian_log_likelihood  # noqa: E402
from src.utils.utils import get_model_params, get_scheduler  # noqa: E402


def neutralise_dropout(model):
    for mod in model.modules():
        if isinstance(mod, nn.Dropout):
            mod.p = 0.0
        if isinstance(mod, nn.MultiheadAttention):
            mod.dropout = 0.0


def inputs(batch):
    g = torch.Generator().manual_seed(0)
    w = torch.randn(batch, 365, 31, generator=g)
    lat = torch.rand(batch, generator=g) * 120 - 60
    lon = torch.rand(batch, generator=g) * 360 - 180
    idx = torch.randint(0, 2, (batch,), generator=g)
    t = torch.arange(365, dtype=torch.float32)
    year = 1984.0 + ((idx[:, None].float() * 365 + t[None]) * 7.0) / 365
    interval = torch.full((batch, 1), 7.0)
    return w, torch.stack([lat, lon], 1), year, interval, idx


def dataset(kind, p=0.15, n=10):
    return StreamingDataset([], masking_function=kind, masking_prob=p, n_masked_features=n)


def sha16(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def model_case(kind, size, batch, beta=0.5):
    torch.manual_seed(1234)
    cls = WeatherBERT if kind == "weatherbert" else WeatherFormer
    model = cls(weather_dim=31, output_dim=31, device=torch.device("cpu"), **get_model_params(size))
    model.train()
    neutralise_dropout(model)
    w, coords, year, interval, idx = inputs(batch)
    torch.manual_seed(1234)
    if kind == "weatherbert":
        mask = dataset("weatherbert", p=0.15).masking_function(365, 31, batch)
    else:
        mask = dataset("weatherformer", n=10).masking_function(365, 31, batch)
    out = model(w, coords, year, interval, weather_feature_mask=mask)
    rec = {}
    if kind == "weatherbert":
        loss = nn.MSELoss(reduction="mean")(w[mask], out[mask])
        y = out
        rec["loss"] = np.array([loss.item()], dtype=np.float64)
    else:
        mu, var = out
        nm = mask.sum(dim=(1, 2)).float().mean()
        recon = (-gaussian_log_likelihood(w, mu, var, mask) / nm).mean()
        kl = beta * compute_gaussian_kl_divergence(mask, mu, var, torch.zeros_like(mu), torch.ones_like(var)).mean() / nm
        loss = recon + kl
        rec["loss"] = np.array([loss.item(), recon.item(), kl.item()], dtype=np.float64)
        rec["mu"] = mu.detach().numpy()
        rec["var"] = var.detach().numpy()
        y = None
    loss.backward()
    if y is not None:
        rec["y"] = y.detach().numpy()
    rec.update(weather=w.numpy(), coords=coords.numpy(), year=year.numpy(), interval=interval.numpy(),
               mask=mask.contiguous().numpy(), beta=np.array([beta]), num_heads=np.array([get_model_params(size)["num_heads"]]))
    for k, v in model.state_dict().items():
        rec["param/" + k] = v.detach().numpy().copy()
    for k, v in model.named_parameters():
        rec["grad/" + k] = v.grad.detach().numpy().copy()
    # one Adam step on these gradients (lr 5e-4) for the optimiser oracle
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)
    opt.step()
    rec["adam/in_proj.weight"] = model.in_proj.weight.detach().numpy()
    rec["adam/out_proj.bias"] = model.out_proj.bias.detach().numpy()
    return rec, model, mask, loss


def yield_case(kind, batch=6, n_past=5):
    """Reference yield models (crop_yield/models/weatherbert_yield_model.py, weatherformer_yield_model.py) on the
    yield loader's input shape: (n_past+1)*52 weekly steps, 6 observed features, the other 25 masked and imputed."""
    from src.crop_yield.models.weatherbert_yield_model import WeatherBERTYieldModel
    from src.crop_yield.models.weatherformer_yield_model import WeatherFormerYieldModel

    torch.manual_seed(1234)
    cls = WeatherBERTYieldModel if kind == "weatherbert" else WeatherFormerYieldModel
    model = cls(name=kind + "_yield", device=torch.device("cpu"), weather_dim=31, n_past_years=n_past,
                **get_model_params("mini"))
    model.train()
    neutralise_dropout(model)
    S = (n_past + 1) * 52
    g = torch.Generator().manual_seed(7)
    obs = [7, 8, 11, 1, 2, 29]
    w = torch.zeros(batch, S, 31)
    w[:, :, obs] = torch.randn(batch, S, 6, generator=g)
    coords = torch.stack([torch.rand(batch, generator=g) * 20 + 30, -(torch.rand(batch, generator=g) * 30 + 80)], 1)
    y0 = torch.randint(1990, 2015, (batch,), generator=g).float()
    week = torch.arange(1, 53, dtype=torch.float32) / 52
    year = (y0[:, None, None] + torch.arange(n_past + 1).float()[None, :, None] + week[None, None, :]).reshape(batch, S)
    interval = torch.full((batch, 1), 7.0)
    mask = torch.ones(batch, S, 31, dtype=torch.bool)
    mask[:, :, obs] = False
    y_past = torch.randn(batch, n_past + 1, generator=g)
    target = torch.randn(batch, 1, generator=g)
    rec = dict(weather=w.numpy(), coords=coords.numpy(), year=year.numpy(), interval=interval.numpy(),
               mask=mask.numpy(), y_past=y_past.numpy(), target=target.numpy())
    torch.manual_seed(99)  # epsilon of the reparameterisation (weatherformer_yield_model.py:58)
    out = model(w, coords, year, interval, mask, y_past)
    if kind == "weatherbert":
        pred = out
        loss = nn.MSELoss(reduction="mean")(pred, target)
    else:
        pred, z, mu, var = out
        torch.manual_seed(99)
        rec["epsilon"] = torch.randn_like(mu).numpy()
        kl = compute_gaussian_kl_divergence(mask, mu, var, torch.zeros_like(mu), torch.ones_like(var)).mean()
        loss = nn.MSELoss(reduction="mean")(pred, target) + 1e-4 * kl
        rec["kl"] = np.array([kl.item()])
        rec["mu"], rec["var"] = mu.detach().numpy(), var.detach().numpy()
    loss.backward()
    rec["pred"] = pred.detach().numpy()
    rec["loss"] = np.array([loss.item()])
    for k, v in model.state_dict().items():
        rec["param/" + k] = v.detach().numpy().copy()
    for k, v in model.named_parameters():
        rec["grad/" + k] = v.grad.detach().numpy().copy()
    return rec


def sibling_case(kind, batch=4, beta=0.5):
    """WeatherFormerSinusoid / WeatherFormerMixture: one trainer-formula loss and its
    gradients on the mini size. Parameters are not stored (same seed => same draws, checked through checksums);
    big gradient tensors are stored as norms plus a strided sample."""
    from src.pretraining.models.weatherformer_mixture import WeatherFormerMixture
    from src.pretraining.models.weatherformer_sinusoid import WeatherFormerSinusoid
    from src.utils.losses import compute_mixture_kl_divergence

    torch.manual_seed(1234)
    if kind == "sinusoid":
        model = WeatherFormerSinusoid(weather_dim=31, output_dim=31, k=4, device=torch.device("cpu"), **get_model_params("mini"))
    else:
        model = WeatherFormerMixture(weather_dim=31, output_dim=31, k=7, device=torch.device("cpu"), **get_model_params("mini"))
    model.train()
    neutralise_dropout(model)
    w, coords, year, interval, _ = inputs(batch)
    torch.manual_seed(1234)
    mask = dataset("weatherformer", n=10).masking_function(365, 31, batch)
    out = model(w, coords, year, interval, weather_feature_mask=mask)
    mu, var = out[0], out[1]
    nm = mask.sum(dim=(1, 2)).float().mean()
    recon = (-gaussian_log_likelihood(w, mu, var, mask) / nm).mean()
    rec = {}
    if kind == "sinusoid":
        kl_raw = compute_gaussian_kl_divergence(mask, mu, var, out[2], out[3])
    else:
        torch.manual_seed(99)
        eps = torch.randn_like(mu)
        rec["epsilon"] = eps.numpy()
        z = mu + torch.sqrt(var) * eps
        kl_raw = compute_mixture_kl_divergence(z=z, feature_mask=mask, mu_x=mu, var_x=var, mu_k=out[2], var_k=out[3], log_w_k=out[4])
    kl = beta * kl_raw.mean() / nm
    loss = recon + kl
    loss.backward()
    rec["loss"] = np.array([loss.item(), recon.item(), kl.item()], dtype=np.float64)
    rec.update(weather=w.numpy(), coords=coords.numpy(), year=year.numpy(), interval=interval.numpy(),
               mask=mask.contiguous().numpy()[:, 0, :], beta=np.array([beta]))
    for k, v in model.named_parameters():
        gnp = v.grad.detach().numpy()
        rec["gnorm/" + k] = np.array([np.linalg.norm(gnp.astype(np.float64))])
        rec["psum/" + k] = np.array([v.detach().double().sum().item(), v.detach().double().abs().sum().item()])
        rec["gsample/" + k] = gnp.reshape(-1)[::max(1, gnp.size // 2048)].copy()
    return rec


def main():
    os.makedirs(OUT, exist_ok=True)
    # ---- masks on the CPU generator (config 1), SURVEY.md 8(c) golden values
    masks = {}
    for name, kind, kw, want_sum, want_sha, want_idx in [
        ("bert_p15", "weatherbert", dict(p=0.15), 108421, "9b5a0f934533c16f", [0, 4, 6]),
        ("bert_p30", "weatherbert", dict(p=0.30), 217284, "25e3b2039b038155", [0, 2, 4, 6, 14, 16, 18, 22, 29]),
        ("former_n10", "weatherformer", dict(n=10), 233600, "8186b49a39436e1f", [0, 1, 2, 6, 10, 11, 12, 14, 21, 23]),
        ("former_n1", "weatherformer", dict(n=1), 23360, "996391d64bc927d0", [0]),
    ]:
        torch.manual_seed(1234)
        m = dataset(kind, **kw).masking_function(365, 31, 64).contiguous().numpy()
        assert int(m.sum()) == want_sum, (name, int(m.sum()))
        assert sha16(m) == want_sha, (name, sha16(m))
        assert np.nonzero(m[0, 0])[0].tolist() == want_idx, name
        masks[name + "_sum"] = np.array([m.sum()])
        masks[name + "_sha16"] = np.frombuffer(want_sha.encode(), dtype=np.uint8)
        masks[name + "_row00"] = m[0, 0]
        masks[name + "_packed"] = np.packbits(m if kind == "weatherbert" else m[:, 0, :])
    for name, pm in [("simmtm_p15", 0.15), ("simmtm_p30", 0.30)]:  # contiguous-segment masks (reference :86-184)
        torch.manual_seed(1234)
        m = dataset("simmtm", p=pm).masking_function(365, 31, 64).contiguous().numpy()
        assert (m == m[:, :, :1]).all() and int(m[:, :, 0].sum(1).max()) <= int(365 * pm)
        masks[name + "_sum"] = np.array([m.sum()])
        masks[name + "_packed"] = np.packbits(m[:, :, 0])
    torch.manual_seed(1234)
    r5 = torch.rand(5).numpy()
    assert np.allclose(r5, [0.028979241847991943, 0.4018985629081726, 0.25984418392181396, 0.3666413426399231,
                            0.05830073356628418], atol=0, rtol=0)
    masks["rand5_seed1234"] = r5
    np.savez_compressed(os.path.join(OUT, "masks_cpu.npz"), **masks)

    # ---- model cases
    rec, model, mask, loss = model_case("weatherbert", "mini", 8)
    assert int(mask.sum()) == 13435
    assert abs(loss.item() - 1.33503056) < 2e-6, loss.item()
    assert abs(model.in_proj.weight.grad.norm().item() - 0.08772561) < 1e-6
    assert abs(model.out_proj.weight.grad.norm().item() - 0.65765929) < 1e-6
    np.savez_compressed(os.path.join(OUT, "weatherbert_mini_b8.npz"), **rec)
    rec, model, mask, loss = model_case("weatherformer", "mini", 8)
    assert abs(rec["loss"][0] - 1.75079405) < 2e-6 and abs(rec["loss"][1] - 1.66437364) < 2e-6
    assert abs(rec["loss"][2] - 0.08642045) < 1e-6
    assert abs(model.in_proj.weight.grad.norm().item() - 0.07980558) < 1e-6
    assert abs(model.out_proj.weight.grad.norm().item() - 0.70168072) < 1e-6
    np.savez_compressed(os.path.join(OUT, "weatherformer_mini_b8.npz"), **rec)

    # ---- yield fine-tune models (BASELINE.json configs[5])
    for kind in ["weatherbert", "weatherformer"]:
        np.savez_compressed(os.path.join(OUT, f"{kind}_yield_mini_b6.npz"), **yield_case(kind))

    # ---- encoder siblings with learned priors
    for kind in ["sinusoid", "mixture"]:
        np.savez_compressed(os.path.join(OUT, f"weatherformer_{kind}_mini_b4.npz"), **sibling_case(kind))

    # ---- checkpoint interchange fixtures: what BaseTrainer.save_checkpoint writes
    # (src/base_trainer/base_trainer.py:121-170): the whole pickled module and the resume dictionary, after one
    # Adam step on the golden gradients of the mini WeatherBERT case
    torch.manual_seed(1234)
    ref = WeatherBERT(weather_dim=31, output_dim=31, device=torch.device("cpu"), **get_model_params("mini"))
    g = dict(np.load(os.path.join(OUT, "weatherbert_mini_b8.npz")))
    for k, prm in ref.named_parameters():
        prm.grad = torch.from_numpy(g["grad/" + k].copy())
    opt = torch.optim.Adam(ref.parameters(), lr=5e-4)
    sch = get_scheduler(opt, 2, 10, 0.99)
    for _ in range(3):  # epoch 0 has lr 0 under warm-up: step a few epochs so the moments AND the weights move
        opt.step()
        sch.step()
    ref.eval()
    w, coords, year, interval, _ = inputs(2)
    mask = torch.zeros(2, 365, 31, dtype=torch.bool)
    mask[:, :, ::3] = True
    with torch.no_grad():
        y = ref(w, coords, year, interval, weather_feature_mask=mask)
    torch.save(ref, os.path.join(OUT, "ref_weatherbert_mini_latest.pth"))
    torch.save({"epoch": 3, "model_state_dict": ref.state_dict(), "optimizer_state_dict": opt.state_dict(),
                "scheduler_state_dict": sch.state_dict(), "best_val_loss": 1.25,
                "output_json": {"model_config": {"total_params": ref.total_params()},
                                "losses": {"train": {"total_loss": [1.5, 1.4, 1.3]}, "val": {"total_loss": [1.6, 1.5, 1.25]}}}},
               os.path.join(OUT, "ref_weatherbert_mini_latest_checkpoint.pth"))
    np.savez_compressed(os.path.join(OUT, "ref_weatherbert_mini_eval.npz"), weather=w.numpy(), coords=coords.numpy(),
                        year=year.numpy(), interval=interval.numpy(), mask=mask.numpy(), y=y.numpy(),
                        lr=np.array([opt.param_groups[0]["lr"]]))

    # ---- scheduler + sizes
    sched = {}
    for nm, warm, total, decay in [("exp", 10.0, 100, 0.99), ("cos", 5, 50, None), ("nowarm", 0, 20, 0.9)]:
        p = nn.Parameter(torch.zeros(1))
        opt = torch.optim.Adam([p], lr=5e-4)
        s = get_scheduler(opt, warm, total, decay)
        lrs = []
        for _ in range(total):
            lrs.append(opt.param_groups[0]["lr"])
            opt.step()
            s.step()
        sched[nm] = np.array(lrs)
    for size in ["mini", "small", "medium", "large"]:
        for kind, cls in [("weatherbert", WeatherBERT), ("weatherformer", WeatherFormer)]:
            m = cls(weather_dim=31, output_dim=31, device=torch.device("cpu"), **get_model_params(size))
            sched[f"params_{kind}_{size}"] = np.array([m.total_params()])
            if size == "mini" and kind == "weatherbert":
                sched["pe_mini"] = m.positional_encoding.pos_encoding.numpy()
    np.savez_compressed(os.path.join(OUT, "schedules.npz"), **sched)
    print("golden vectors written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
