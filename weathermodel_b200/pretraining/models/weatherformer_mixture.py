"""WeatherFormerMixture: WeatherFormer with a learned mixture-of-Gaussians prior -- k sinusoidal means, k
variances and k mixture weights (reference src/pretraining/models/weatherformer_mixture.py:14-147)."""
import copy
from typing import Optional, Tuple

import torch
import torch.nn as nn

from ...utils.constants import MAX_CONTEXT_LENGTH
from .weatherformer import WeatherFormer
from .weatherformer_sinusoid import _sinusoid_means


class WeatherFormerMixture(WeatherFormer):
    def __init__(self, weather_dim, output_dim, device, k=7, num_heads=20, num_layers=8, hidden_dim_factor=24,
                 max_len=MAX_CONTEXT_LENGTH):
        super().__init__(weather_dim=weather_dim, output_dim=output_dim, num_heads=num_heads, num_layers=num_layers,
                         hidden_dim_factor=hidden_dim_factor, max_len=max_len, device=device)
        self.name = "weatherformer_mixture"
        self.k = k
        self.positions = torch.arange(max_len, dtype=torch.float, device=device).view(1, 1, max_len, 1)
        # same draw order as the reference: frequency, phase, amplitude, log-variances
        self.frequency = nn.Parameter(torch.randn(1, k, max_len, output_dim) * 0.1)
        self.phase = nn.Parameter(torch.randn(1, k, max_len, output_dim) * 0.1)
        self.amplitude = nn.Parameter(torch.randn(1, k, max_len, output_dim) * 0.1)
        self.log_var_k = nn.Parameter(torch.randn(1, k, max_len, output_dim) * 0.1 - 1.0)
        uniform = -torch.log(torch.tensor(k, dtype=torch.float32)).item()  # log(1 / k)
        self.mixture_logits = nn.Parameter(torch.full((1, k), uniform))

    def load_pretrained(self, pretrained_model: "WeatherFormerMixture", load_out_proj=True):
        if self.k != pretrained_model.k:
            raise ValueError(f"k mismatch: {self.k} != {pretrained_model.k}. Please ensure the models are compatible.")
        super().load_pretrained(pretrained_model, load_out_proj)
        if load_out_proj:
            for name in ("frequency", "phase", "amplitude", "log_var_k", "mixture_logits"):
                setattr(self, name, copy.deepcopy(getattr(pretrained_model, name)))
        self.k = pretrained_model.k

    def forward(self, weather, coords, year, interval, weather_feature_mask,
                src_key_padding_mask: Optional[torch.Tensor] = None
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        mu_x, var_x = super().forward(weather=weather, coords=coords, year=year, interval=interval,
                                      weather_feature_mask=weather_feature_mask,
                                      src_key_padding_mask=src_key_padding_mask)
        batch, seq_len = weather.shape[0], weather.shape[1]
        mu_k = _sinusoid_means(self, seq_len, interval)
        var_k = torch.clamp(torch.exp(self.log_var_k[:, :, :seq_len, :]), min=1e-6, max=1).expand(batch, -1, -1, -1)
        log_w_k = torch.log_softmax(self.mixture_logits, dim=1).expand(batch, -1)
        return mu_x, var_x, mu_k, var_k, log_w_k
