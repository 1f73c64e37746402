"""WeatherAutoencoder: the deterministic baseline -- a WeatherBERT encoder trained with MSE under WeatherFormer's
whole-feature masking (reference src/pretraining/models/weatherautoencoder.py:15-36)."""
from ...utils.constants import MAX_CONTEXT_LENGTH
from .weatherbert import WeatherBERT


class WeatherAutoencoder(WeatherBERT):
    def __init__(self, weather_dim, output_dim, device, num_heads=20, num_layers=8, hidden_dim_factor=24,
                 max_len=MAX_CONTEXT_LENGTH):
        super().__init__(weather_dim=weather_dim, output_dim=output_dim, num_heads=num_heads, num_layers=num_layers,
                         hidden_dim_factor=hidden_dim_factor, max_len=max_len, device=device)
        self.name = "weatherautoencoder"
