"""WeatherFormerYieldModel: the yield head on a WeatherFormer encoder with the reparameterisation trick
(reference src/crop_yield/models/weatherformer_yield_model.py:7-71). forward returns (yield, z, mu, var)."""
import torch

from ...engine import yield_head
from ...pretraining.models.weatherformer import WeatherFormer
from .weatherbert_yield_model import WeatherBERTYieldModel


class WeatherFormerYieldModel(WeatherBERTYieldModel):
    def __init__(self, name: str, device: torch.device, weather_dim: int, n_past_years: int, **model_size_params):
        super().__init__(name, device, weather_dim, n_past_years, **model_size_params)
        self.weather_model = WeatherFormer(weather_dim=weather_dim, output_dim=weather_dim, device=device,
                                           **model_size_params)

    def forward(self, padded_weather, coord, year, interval, weather_feature_mask, y_past):
        mu_x, var_x = self.weather_model(padded_weather, coord, year=year, interval=interval,
                                         weather_feature_mask=weather_feature_mask)
        if self._fused_head_ok(padded_weather, y_past) and getattr(mu_x, "_wm_raw", None) is not None:
            # the same epsilon draw as the reference (torch generator, randn_like), then ONE kernel: reparameterise,
            # impute, score, softmax-pool, MLP. mu / var stay torch views of the raw output for the trainer's KL term.
            eps = torch.randn_like(mu_x)
            pred, z = yield_head(mu_x._wm_raw, padded_weather, weather_feature_mask, eps, y_past, self._head_params(),
                                 is_former=True)
            return pred, z, mu_x, var_x
        z = mu_x + torch.sqrt(var_x) * torch.randn_like(mu_x)  # z ~ N(mu, var)
        z = self._impute_weather(padded_weather, z, weather_feature_mask)
        pred = self.yield_model(z, coord, year, interval, weather_feature_mask=None, y_past=y_past)
        return pred, z, mu_x, var_x
