"""Torch-CPU port of the reference training step -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

The reference (Neehan/WeatherModel) is pure Python on top of stock torch modules and cannot travel to
the GPU box (/root/reference does not exist there, and its sources may not be copied). Its arithmetic,
however, lives entirely in torch: nn.Linear, nn.TransformerEncoder (post-LN, ReLU, dropout 0.1),
F.mse_loss / elementwise ops, optim.Adam. This module restates the reference's step with exactly those
stock modules in fp32 on the host cores, so bench.py can time "the reference's CPU PyTorch path" next to
the GPU numbers (cpu_baseline.kind = "port") and `bench.py --impl reference` has something to run.

Restated from (paths under /root/reference):
  src/pretraining/models/weatherbert.py:14-56,101-121   model construction and forward
  src/pretraining/models/weatherformer.py:42,87-92       (mu, clamp(exp(logvar))) head
  src/utils/utils.py:63-74                               input normalisation
  src/base_models/vanilla_pos_encoding.py:23-37,57       sinusoidal table
  src/pretraining/trainers/weatherbert_trainer.py:55-60  masked MSE
  src/pretraining/trainers/weatherformer_trainer.py:91-105, src/utils/losses.py:25-27,41-46   ELBO
  src/pretraining/dataloader/pretraining_dataloader.py:56-84   mask functions
  src/base_trainer/base_trainer.py:241-252,337           zero_grad / backward / Adam step
tests/test_oracle.py::test_torch_port_matches_golden pins it against the golden vectors.
"""
import math

import torch
import torch.nn as nn


class PortModel(nn.Module):
    def __init__(self, kind: str, num_heads: int, num_layers: int, hidden_dim_factor: int, weather_dim: int = 31,
                 max_len: int = 365):
        super().__init__()
        assert kind in ("weatherbert", "weatherformer")
        self.kind = kind
        self.weather_dim = weather_dim
        d = num_heads * hidden_dim_factor
        self.in_proj = nn.Linear(weather_dim + 3, d)
        pe = torch.zeros(max_len, d)
        pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div = torch.exp(torch.arange(0, d, 2).float() * (-math.log(10000.0) / d))
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        self.register_buffer("pos_encoding", pe)
        layer = nn.TransformerEncoderLayer(batch_first=True, d_model=d, nhead=num_heads, dim_feedforward=4 * d)
        self.transformer_encoder = nn.TransformerEncoder(layer, num_layers=num_layers)
        self.out_proj = nn.Linear(d, weather_dim * (2 if kind == "weatherformer" else 1))

    def load_reference_state(self, state):
        """state: dict keyed by the reference's state_dict names (numpy arrays or tensors)."""
        own = self.state_dict()
        for k, v in state.items():
            kk = "pos_encoding" if k == "positional_encoding.pos_encoding" else k
            own[kk].copy_(torch.as_tensor(v))

    def forward(self, weather, coords, year, interval, mask):
        b, s, _ = weather.shape
        year = ((year - 1970) / 100.0).unsqueeze(2)
        c = coords.clone()
        c[:, 0] = c[:, 0] / 360
        c[:, 1] = c[:, 1] / 180
        x = torch.cat([weather * (~mask), year, c.unsqueeze(1).expand(b, s, 2)], dim=2)
        h = self.in_proj(x) + self.pos_encoding[:s].unsqueeze(0)
        y = self.out_proj(self.transformer_encoder(h))
        if self.kind == "weatherbert":
            return y
        f = self.weather_dim
        return y[..., :f], torch.clamp(torch.exp(y[..., f:]), min=1e-6, max=1)


def port_loss(model: PortModel, weather, coords, year, interval, mask, beta: float = 0.5):
    out = model(weather, coords, year, interval, mask)
    if model.kind == "weatherbert":
        return {"total_loss": nn.functional.mse_loss(weather[mask], out[mask])}
    mu, var = out
    n_bar = mask.sum(dim=(1, 2)).float().mean()
    ll = (-0.5 * torch.log(2 * torch.pi * var) - 0.5 * (weather - mu) ** 2 / var) * mask
    recon = (-ll.sum(dim=(1, 2)) / n_bar).mean()
    kl = 0.5 * (torch.log(1.0 / var) + var + mu ** 2 - 1.0) * mask
    kl = beta * kl.sum(dim=(1, 2)).mean() / n_bar
    return {"total_loss": recon + kl, "reconstruction": recon, "kl_term": kl}


def port_mask(kind: str, batch: int, seq_len: int, n_features: int, masking_prob: float, n_masked: int):
    if kind == "weatherbert":
        return torch.rand(batch, seq_len, n_features) < masking_prob
    order = torch.argsort(torch.rand(batch, n_features), dim=-1)
    return (order < n_masked).unsqueeze(1).expand(-1, seq_len, -1)


def neutralise_dropout(model: nn.Module):
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0


def time_port_steps(kind: str, size_params: dict, batch: int, steps: int, warmup: int, threads: int,
                    masking_prob: float = 0.15, n_masked: int = 10, beta: float = 0.5, seed: int = 0):
    """Full reference step (mask -> forward -> loss -> backward -> Adam), fp32, dropout ON as shipped.
    Returns (seconds per step list, last loss)."""
    import time

    torch.set_num_threads(threads)
    torch.manual_seed(1234)
    model = PortModel(kind, **size_params).train()
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)
    g = torch.Generator().manual_seed(seed)
    weather = torch.randn(batch, 365, 31, generator=g)
    coords = torch.stack([torch.rand(batch, generator=g) * 120 - 60, torch.rand(batch, generator=g) * 360 - 180], 1)
    idx = torch.randint(0, 2, (batch,), generator=g).float()
    year = 1984.0 + ((idx[:, None] * 365 + torch.arange(365, dtype=torch.float32)[None]) * 7.0) / 365
    interval = torch.full((batch, 1), 7.0)
    times, last = [], float("nan")
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        mask = port_mask(kind, batch, 365, 31, masking_prob, n_masked)
        opt.zero_grad()
        loss = port_loss(model, weather, coords, year, interval, mask, beta)["total_loss"]
        loss.backward()
        opt.step()
        last = loss.item()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return times, last
