"""WeatherFormerYieldTrainer: yield MSE + beta * mean_b KL(q(z|x) || N(0, 1)) over the masked (imputed) features
(reference src/crop_yield/trainers/weatherformer_yield_trainer.py:72-131; note the KL here is NOT divided by the
masked-feature count, unlike pretraining, and the reconstruction term is the constant 0)."""
from typing import Dict

import torch
import torch.nn as nn

from ...utils.constants import DRY_RUN
from ...utils.losses import compute_gaussian_kl_divergence
from ..models.weatherformer_yield_model import WeatherFormerYieldModel
from .weatherbert_yield_trainer import (WeatherBERTYieldTrainer, _create_yield_training_setup,
                                        _run_yield_cross_validation)


class WeatherFormerYieldTrainer(WeatherBERTYieldTrainer):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.criterion = nn.MSELoss(reduction="mean")
        if self.rank == 0:
            self.output_json["losses"] = {
                "train": {"total_loss": [], "yield": [], "reconstruction": [], "kl_term": []},
                "val": {"total_loss": []},
            }

    def compute_kl_loss(self, weather_feature_mask, z, mu_x, var_x, *args) -> torch.Tensor:
        return compute_gaussian_kl_divergence(feature_mask=weather_feature_mask, mu_x=mu_x, var_x=var_x,
                                              mu_p=torch.zeros_like(mu_x), var_p=torch.ones_like(var_x))

    def compute_elbo_loss(self, weather, weather_feature_mask, target_yield, yield_pred, z, mu_x, var_x, *args,
                          log_losses: bool = False) -> Dict[str, torch.Tensor]:
        yield_loss = self.criterion(yield_pred.squeeze(), target_yield.squeeze())
        reconstruction = torch.zeros((), device=yield_loss.device)
        kl = self._current_beta() * self.compute_kl_loss(weather_feature_mask, z, mu_x, var_x, *args).mean()
        if log_losses or DRY_RUN:
            self.logger.info(f"Yield Loss: {yield_loss.item():.6f}")
            self.logger.info(f"KL Term: {kl.item():.6f}")
        return {"total_loss": yield_loss + reconstruction + kl, "yield": yield_loss, "reconstruction": reconstruction,
                "kl_term": kl}

    def compute_train_loss(self, padded_weather, coord_processed, year_expanded, interval, weather_feature_mask,
                           practices, soil, y_past, target_yield) -> Dict[str, torch.Tensor]:
        outputs = self.model(padded_weather, coord_processed, year_expanded, interval, weather_feature_mask, y_past)
        return self.compute_elbo_loss(padded_weather, weather_feature_mask, target_yield, *outputs)

    def compute_validation_loss(self, padded_weather, coord_processed, year_expanded, interval, weather_feature_mask,
                                practices, soil, y_past, target_yield) -> Dict[str, torch.Tensor]:
        with torch.no_grad():
            outputs = self.model(padded_weather, coord_processed, year_expanded, interval, weather_feature_mask, y_past)
        parts = self.compute_elbo_loss(padded_weather, weather_feature_mask, target_yield, *outputs)
        return {"total_loss": parts["yield"] ** 0.5}


def weatherformer_yield_training_loop(args_dict, use_cropnet: bool):
    return _run_yield_cross_validation(
        setup_params=_create_yield_training_setup(args_dict, use_cropnet), model_class=WeatherFormerYieldModel,
        trainer_class=WeatherFormerYieldTrainer, model_name=f"weatherformer_{args_dict['crop_type']}_yield",
        args_dict=args_dict, extra_trainer_kwargs={"beta": args_dict["beta"]})
