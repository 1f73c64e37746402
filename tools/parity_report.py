#!/usr/bin/env python
"""Gradient / loss parity report of the B200 path against the oracle, per parameter tensor.

    python tools/parity_report.py [--out profiles/r02_parity.txt] [--cases mini,small,medium,large]

Cases: the two golden mini vectors written by the UNMODIFIED reference (tests/golden/*_mini_b8.npz) and freshly seeded
small / medium / large models against the numpy fp64 oracle (oracle/wm_oracle.py) -- the same cases as
tests/test_gpu_model.py, with every measured error printed instead of only asserted. For calibration the same step is
also run through torch's own mixed-precision path on the same GPU (the torch port of the reference under
torch.autocast(bfloat16), fused SDPA, fp32 master weights): the error a stock PyTorch bf16 user sees against fp32.
Columns: relative Frobenius error ||g - g_ref|| / ||g_ref||, relative norm error | ||g|| - ||g_ref|| | / ||g_ref||.
Test infrastructure (imports oracle/); needs a B200.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch_port  # noqa: E402
import wm_oracle as O  # noqa: E402

from weathermodel_b200 import engine  # noqa: E402
from weathermodel_b200.pretraining.models.weatherbert import WeatherBERT  # noqa: E402
from weathermodel_b200.pretraining.models.weatherformer import WeatherFormer  # noqa: E402

DEV = "cuda"
GOLD = os.path.join(ROOT, "tests", "golden")


def neutralise(model):
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def grad_table(grads, ref):
    rows, sq_d, sq_r, sq_g = [], 0.0, 0.0, 0.0
    for name, r in ref.items():
        g = np.asarray(grads[name], np.float64)
        r = np.asarray(r, np.float64)
        ng, nr = np.linalg.norm(g), np.linalg.norm(r)
        rows.append((name, rel(g, r), abs(ng - nr) / (nr + 1e-300), nr))
        sq_d += np.linalg.norm(g - r) ** 2
        sq_r += nr ** 2
        sq_g += ng ** 2
    glob = {"fro": float(np.sqrt(sq_d / sq_r)), "norm": float(abs(np.sqrt(sq_g) - np.sqrt(sq_r)) / np.sqrt(sq_r))}
    return rows, glob


def run_ours(model, kind, batch, beta):
    w, c, yr, iv, mask = batch
    model.zero_grad()
    y_pad = model.forward_raw(w, c, yr, iv, mask)
    if kind == "weatherbert":
        losses = {"total_loss": engine.bert_masked_mse(y_pad, w, mask)}
    else:
        losses = engine.former_elbo(y_pad, w, mask, beta)
    losses["total_loss"].backward()
    grads = {n: p.grad.detach().float().cpu().numpy() for n, p in model.named_parameters()}
    return {k: v.item() for k, v in losses.items()}, grads, y_pad.detach().float().cpu().numpy()


def run_torch_autocast(kind, hp, state, batch, beta):
    """The reference's step through torch's own bf16 path on the same GPU (calibration only)."""
    port = torch_port.PortModel(kind, **hp).to(DEV).train()
    port.load_reference_state({k: torch.as_tensor(v) for k, v in state.items()})
    torch_port.neutralise_dropout(port)
    w, c, yr, iv, mask = batch
    mask = mask.contiguous()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        losses = torch_port.port_loss(port, w, c, yr, iv, mask, beta)
    losses["total_loss"].backward()
    grads = {}
    for n, p in port.named_parameters():
        grads[n] = p.grad.detach().float().cpu().numpy()
    return {k: v.item() for k, v in losses.items()}, grads


def report(title, kind, losses, ref_losses, grads, ref_grads, y=None, y_ref=None, extra=None, out=None, bf16_grads=None,
           bf16_losses=None):
    rows, glob = grad_table(grads, ref_grads)
    lines = [f"== {title}"]
    for k, r in ref_losses.items():
        lines.append(f"   loss {k:15s} got {losses[k]:.8f} ref {r:.8f} rel {abs(losses[k] - r) / abs(r):.2e}")
    if y is not None:
        lines.append(f"   head output rel fro {rel(y, y_ref):.2e}")
    worst = max(rows, key=lambda t: t[1])
    fro = np.array([t[1] for t in rows])
    nrm = np.array([t[2] for t in rows])
    lines.append(f"   gradients: {len(rows)} tensors; rel fro max {fro.max():.2e} ({worst[0]}) median {np.median(fro):.2e}; "
                 f"norm err max {nrm.max():.2e} median {np.median(nrm):.2e}; GLOBAL rel fro {glob['fro']:.2e} norm {glob['norm']:.2e}")
    if extra is not None:
        rows2, glob2 = grad_table(extra, ref_grads)
        fro2 = np.array([t[1] for t in rows2])
        lines.append(f"   torch autocast(bf16) on the same GPU: rel fro max {fro2.max():.2e} median {np.median(fro2):.2e}; "
                     f"GLOBAL rel fro {glob2['fro']:.2e} norm {glob2['norm']:.2e}")
        extra_map = {t[0]: t[1] for t in rows2}
    q_map = None
    if bf16_grads is not None:  # the same algorithm with a rounding wherever the kernels store bf16 (EncoderOracleBf16)
        rows3, glob3 = grad_table(grads, bf16_grads)
        fro3 = np.array([t[1] for t in rows3])
        w3 = max(rows3, key=lambda t: t[1])
        lines.append(f"   vs bf16-STORAGE oracle (same roundings as the kernels): rel fro max {fro3.max():.2e} ({w3[0]}) median "
                     f"{np.median(fro3):.2e}; GLOBAL rel fro {glob3['fro']:.2e} norm {glob3['norm']:.2e}; loss rel "
                     f"{abs(losses['total_loss'] - bf16_losses['total_loss']) / abs(bf16_losses['total_loss']):.2e}")
        q_map = {t[0]: t[1] for t in rows3}
    lines.append(f"   {'tensor':58s} {'rel fro':>9s} {'norm err':>9s} {'||ref||':>10s}" + ("  torch-bf16 fro" if extra is not None else "") + ("  vs bf16-oracle" if q_map else ""))
    for name, f, n, nr in rows:
        lines.append(f"   {name:58s} {f:9.2e} {n:9.2e} {nr:10.3e}" + (f"  {extra_map[name]:9.2e}" if extra is not None else "") +
                     (f"      {q_map[name]:9.2e}" if q_map else ""))
    text = "\n".join(lines)
    print(text, flush=True)
    if out:
        out.write(text + "\n")
    res = {"fro_max": float(fro.max()), "fro_median": float(np.median(fro)), "norm_max": float(nrm.max()),
           "global_fro": glob["fro"], "global_norm": glob["norm"]}
    if q_map:
        res.update(bf16_oracle_fro_max=float(fro3.max()), bf16_oracle_global_fro=glob3["fro"])
    return res


def golden_case(fname, cls, kind, out):
    g = dict(np.load(os.path.join(GOLD, fname)))
    torch.manual_seed(1234)
    hp = O.get_model_params("mini")
    model = cls(weather_dim=31, output_dim=31, device=torch.device("cpu"), **hp).to(DEV).train()
    neutralise(model)
    t = lambda k: torch.from_numpy(g[k]).to(DEV).float()  # noqa: E731
    mask = torch.from_numpy(g["mask"]).to(DEV)
    if kind == "weatherformer":
        mask = mask[:, :1, :].expand(-1, 365, -1)
    batch = (t("weather"), t("coords"), t("year"), t("interval"), mask)
    beta = float(g["beta"][0]) if "beta" in g else 0.5
    losses, grads, y = run_ours(model, kind, batch, beta)
    ref_grads = {k[len("grad/"):]: v for k, v in g.items() if k.startswith("grad/")}
    names = ("total_loss", "reconstruction", "kl_term") if kind == "weatherformer" else ("total_loss",)
    ref_losses = dict(zip(names, [float(v) for v in g["loss"]]))
    state = {k[len("param/"):]: v for k, v in g.items() if k.startswith("param/")}
    _, tgrads = run_torch_autocast(kind, hp, state, batch, beta)
    lq, _, gq = O.train_step_grads(state, hp["num_heads"], kind, g["weather"], g["coords"], g["year"], g["interval"], g["mask"],
                                   beta=beta, storage="bf16")
    return report(f"{kind} mini, B=8, golden vectors of the unmodified reference ({fname})", kind, losses, ref_losses, grads,
                  ref_grads, extra=tgrads, out=out, bf16_grads=gq, bf16_losses=lq)


def oracle_case(kind, size, B, S, out):
    torch.manual_seed(7)
    cls = WeatherBERT if kind == "weatherbert" else WeatherFormer
    hp = O.get_model_params(size)
    model = cls(weather_dim=31, output_dim=31, device=torch.device(DEV), **hp).to(DEV).train()
    neutralise(model)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(torch.randn_like(p) * 0.02)
    state = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    weather, coords, year, interval = O.synthetic_batch(B, S, seed=3)
    rs = np.random.RandomState(5)
    if kind == "weatherbert":
        mask = rs.rand(B, S, 31) < 0.3
    else:
        mask = np.ascontiguousarray(O.weatherformer_mask(rs.rand(B, 31).astype(np.float32), 10, S))
    ref_losses, y_ref, ref_grads = O.train_step_grads(state, hp["num_heads"], kind, weather, coords, year, interval, mask, beta=0.5)
    batch = tuple(torch.from_numpy(a).to(DEV) for a in (weather, coords, year, interval)) + (torch.from_numpy(mask).to(DEV),)
    losses, grads, y = run_ours(model, kind, batch, 0.5)
    _, tgrads = run_torch_autocast(kind, hp, state, batch, 0.5)
    lq, _, gq = O.train_step_grads(state, hp["num_heads"], kind, weather, coords, year, interval, mask, beta=0.5, storage="bf16")
    return report(f"{kind} {size}, B={B}, S={S}, numpy fp64 oracle", kind, losses, ref_losses, grads, ref_grads,
                  y[..., : y_ref.shape[-1]], y_ref, extra=tgrads, out=out, bf16_grads=gq, bf16_losses=lq)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_parity.txt"))
    ap.add_argument("--cases", default="mini,small,medium,large")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    cases = args.cases.split(",")
    summary = {}
    with open(args.out, "w") as out:
        out.write("# tools/parity_report.py on " + torch.cuda.get_device_name(0) + "; dropout neutralised on both sides; "
                  "north_star asks 1e-3 relative on losses and gradients with fp32 accumulation\n")
        if "mini" in cases:
            summary["bert-mini-golden"] = golden_case("weatherbert_mini_b8.npz", WeatherBERT, "weatherbert", out)
            summary["former-mini-golden"] = golden_case("weatherformer_mini_b8.npz", WeatherFormer, "weatherformer", out)
        if "small" in cases:
            summary["former-small"] = oracle_case("weatherformer", "small", 3, 365, out)
        if "medium" in cases:
            summary["bert-medium"] = oracle_case("weatherbert", "medium", 2, 364, out)
        if "large" in cases:
            summary["former-large"] = oracle_case("weatherformer", "large", 2, 365, out)
        import json
        line = "SUMMARY " + json.dumps(summary)
        print(line)
        out.write(line + "\n")


if __name__ == "__main__":
    main()
