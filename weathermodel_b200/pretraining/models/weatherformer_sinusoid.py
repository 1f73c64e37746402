"""WeatherFormerSinusoid: WeatherFormer whose prior is a learned sum of k sinusoids per (position, feature) with a
learned variance (reference src/pretraining/models/weatherformer_sinusoid.py:16-125). The encoder runs on the
sm_100a runtime; the prior is a handful of [k, S, F]-sized torch ops on the same device (its parameters live
outside the encoder's flat bucket and are stepped by FusedAdam's per-tensor launches)."""
import copy
from typing import Optional, Tuple

import torch
import torch.nn as nn

from ...utils.constants import DEVICE, MAX_CONTEXT_LENGTH
from .weatherformer import WeatherFormer


def _sinusoid_means(module, seq_len: int, interval: torch.Tensor) -> torch.Tensor:
    """amplitude * sin(frequency * 2 pi pos interval / 365 + phase), shape [B, k, S, F]."""
    pos = module.positions[:, :, :seq_len, :].to(interval.device)
    scaled = pos * 2 * torch.pi * interval.view(-1, 1, 1, 1) / 365.0
    return module.amplitude[:, :, :seq_len, :] * torch.sin(module.frequency[:, :, :seq_len, :] * scaled
                                                            + module.phase[:, :, :seq_len, :])


class WeatherFormerSinusoid(WeatherFormer):
    def __init__(self, weather_dim, output_dim, k=4, num_heads=20, num_layers=8, hidden_dim_factor=24,
                 max_len=MAX_CONTEXT_LENGTH, device=DEVICE):
        super().__init__(weather_dim=weather_dim, output_dim=output_dim, num_heads=num_heads, num_layers=num_layers,
                         hidden_dim_factor=hidden_dim_factor, max_len=max_len, device=device)
        self.name = "weatherformer_sinusoid"
        self.positions = torch.arange(max_len, dtype=torch.float, device=device).reshape(1, 1, max_len, 1)
        self.k = k
        # same draw order as the reference: frequency, phase, amplitude, log-variance
        self.frequency = nn.Parameter(torch.randn(1, k, max_len, weather_dim) * 0.1)
        self.phase = nn.Parameter(torch.randn(1, k, max_len, weather_dim) * 0.1)
        self.amplitude = nn.Parameter(torch.randn(1, k, max_len, weather_dim) * 0.1)
        self.log_var_prior = nn.Parameter(torch.randn(1, max_len, weather_dim) * 0.1 - 1)

    def load_pretrained(self, pretrained_model: "WeatherFormerSinusoid", load_out_proj=True):
        super().load_pretrained(pretrained_model, load_out_proj)
        if self.k != pretrained_model.k:
            raise ValueError(f"k mismatch: {self.k} != {pretrained_model.k}. Please set k to the same value.")
        for name in ("frequency", "phase", "amplitude", "log_var_prior"):
            setattr(self, name, copy.deepcopy(getattr(pretrained_model, name)))

    def forward(self, weather, coords, year, interval, weather_feature_mask,
                src_key_padding_mask: Optional[torch.Tensor] = None
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        mu_x, var_x = super().forward(weather=weather, coords=coords, year=year, interval=interval,
                                      weather_feature_mask=weather_feature_mask,
                                      src_key_padding_mask=src_key_padding_mask)
        batch, seq_len = weather.shape[0], weather.shape[1]
        mu_p = _sinusoid_means(self, seq_len, interval).sum(dim=1)
        var_p = torch.clamp(torch.exp(self.log_var_prior)[:, :seq_len, :].expand(batch, -1, -1), min=1e-6, max=1)
        return mu_x, var_x, mu_p, var_p
