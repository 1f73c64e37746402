"""WeatherBERTYieldModel: county-level yield regression on top of the B200 encoder
(reference src/crop_yield/models/weatherbert_yield_model.py:11-132).

The encoder call -- the only part that matters for time -- runs on the sm_100a kernels through
WeatherBERT.forward (autograd flows back into the fused backward). The head is the reference's own tiny torch
modules: masked-feature imputation, a 31->16->1 attention pooling over the sequence and a (31 + n_past_years + 1)
->120->1 MLP (a few hundred FLOPs per sequence, "next" row f2 for a fused kernel)."""
from typing import Union

import torch
import torch.nn as nn

from ...base_models.base_model import BaseModel
from ...pretraining.models.weatherbert import WeatherBERT


class WeatherBERTYieldModel(BaseModel):
    def __init__(self, name: str, device: torch.device, weather_dim: int, n_past_years: int, **model_size_params):
        super().__init__(name)
        self.weather_model = WeatherBERT(weather_dim=weather_dim, output_dim=weather_dim, device=device,
                                         **model_size_params)
        self.weather_attention = nn.Sequential(nn.Linear(weather_dim, 16), nn.GELU(), nn.Linear(16, 1))
        self.yield_mlp = nn.Sequential(nn.Linear(weather_dim + n_past_years + 1, 120), nn.GELU(), nn.Linear(120, 1))
        self.weather_model_frozen = False

    def yield_model(self, weather, coord, year, interval, weather_feature_mask, y_past):
        """softmax-pool the (imputed) weather over the sequence, append past yields, regress."""
        scores = torch.softmax(self.weather_attention(weather), dim=1)  # [B, S, 1]
        pooled = torch.sum(weather * scores, dim=1)                     # [B, weather_dim]
        return self.yield_mlp(torch.cat([pooled, y_past], dim=1))

    def _impute_weather(self, original_weather, imputed_weather, weather_feature_mask):
        """observed features where the mask is False, model output where it is True"""
        return original_weather * (~weather_feature_mask) + imputed_weather * weather_feature_mask

    def load_pretrained(self, pretrained_model: Union[WeatherBERT, "WeatherBERTYieldModel"]):
        self.logger.info(f"provided model class: {pretrained_model.__class__.__name__}")
        if isinstance(pretrained_model, WeatherBERTYieldModel):
            encoder = pretrained_model.weather_model
            self.weather_attention = pretrained_model.weather_attention
            self.yield_mlp = pretrained_model.yield_mlp
        elif isinstance(pretrained_model, WeatherBERT):
            encoder = pretrained_model
        else:
            raise ValueError(f"provided model class: {pretrained_model.__class__.__name__} is not supported")
        self.weather_model.load_pretrained(encoder, load_out_proj=True)

    def forward(self, weather, coord, year, interval, weather_feature_mask, y_past):
        predicted = self.weather_model(weather, coord, year, interval, weather_feature_mask=weather_feature_mask)
        filled = self._impute_weather(weather, predicted, weather_feature_mask)
        return self.yield_model(filled, coord, year, interval, weather_feature_mask=None, y_past=y_past)

    def _set_encoder_trainable(self, flag: bool):
        for p in self.weather_model.parameters():
            p.requires_grad = flag

    def freeze_weather_model(self):
        if not self.weather_model_frozen:
            self.logger.info("Freezing weather model")
            self._set_encoder_trainable(False)
            self.weather_model_frozen = True

    def unfreeze_weather_model(self):
        if self.weather_model_frozen:
            self.logger.info("Unfreezing weather model")
            self._set_encoder_trainable(True)
            self.weather_model_frozen = False
