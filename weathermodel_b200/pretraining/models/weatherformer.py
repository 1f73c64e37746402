"""WeatherFormer: WeatherBERT encoder with a (mu, sigma^2) head
(reference src/pretraining/models/weatherformer.py:17-94)."""
from typing import Optional, Tuple, Union

import torch
import torch.nn as nn

from ...utils.constants import MAX_CONTEXT_LENGTH
from .weatherbert import WeatherBERT


class WeatherFormer(WeatherBERT):
    def __init__(self, weather_dim, output_dim, device, num_heads=20, num_layers=8, hidden_dim_factor=24,
                 max_len=MAX_CONTEXT_LENGTH):
        super().__init__(weather_dim=weather_dim, output_dim=output_dim, num_heads=num_heads, num_layers=num_layers,
                         hidden_dim_factor=hidden_dim_factor, max_len=max_len, device=device)
        self.name = "weatherformer"
        # twice the width: [mu | log sigma^2]
        self.out_proj = nn.Linear(hidden_dim_factor * num_heads, 2 * output_dim)

    def load_pretrained(self, pretrained_model: Union["WeatherBERT", "WeatherFormer"], load_out_proj: bool = True):
        if isinstance(pretrained_model, WeatherFormer):
            super().load_pretrained(pretrained_model, load_out_proj=load_out_proj)
        elif isinstance(pretrained_model, WeatherBERT):
            super().load_pretrained(pretrained_model, load_out_proj=False)  # head widths differ
        else:
            raise ValueError("Expected pretrained model to be either WeatherBERT or WeatherFormer, "
                             f"but got {type(pretrained_model)}")

    def forward(self, weather: torch.Tensor, coords: torch.Tensor, year: torch.Tensor, interval: torch.Tensor,
                weather_feature_mask: torch.Tensor, src_key_padding_mask: Optional[torch.Tensor] = None
                ) -> Tuple[torch.Tensor, torch.Tensor]:
        y = self.forward_raw(weather, coords, year, interval, weather_feature_mask, src_key_padding_mask)
        mu_x = y[..., : self.output_dim]
        var_x = torch.clamp(torch.exp(y[..., self.output_dim: 2 * self.output_dim]), min=1e-6, max=1)
        mu_x._wm_raw = y  # lets the fused ELBO kernel start from the raw head output
        return mu_x, var_x
