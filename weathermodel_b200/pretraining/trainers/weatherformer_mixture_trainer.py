"""WeatherFormerMixtureTrainer: ELBO whose KL is a single-sample estimate against the learned mixture prior
(reference src/pretraining/trainers/weatherformer_mixture_trainer.py:14-102)."""
import torch

from ...utils.losses import compute_mixture_kl_divergence
from ..models.weatherformer import WeatherFormer
from ..models.weatherformer_mixture import WeatherFormerMixture
from .weatherformer_sinusoid_trainer import _former_family_loop
from .weatherformer_trainer import WeatherFormerTrainer


class WeatherFormerMixtureTrainer(WeatherFormerTrainer):
    def __init__(self, model: WeatherFormer, masking_prob: float, n_masked_features: int, beta: float, **kwargs):
        super().__init__(model=model, masking_prob=masking_prob, n_masked_features=n_masked_features, beta=beta,
                         **kwargs)
        self.output_json["model_config"]["n_mixture_components"] = model.k

    def compute_kl_loss(self, weather, weather_feature_mask, mu_x, var_x, mu_k, var_k, log_w_k) -> torch.Tensor:
        z = mu_x + torch.sqrt(var_x) * torch.randn_like(mu_x)  # z ~ q(z | x)
        return compute_mixture_kl_divergence(z=z, feature_mask=weather_feature_mask, mu_x=mu_x, var_x=var_x,
                                             mu_k=mu_k, var_k=var_k, log_w_k=log_w_k)


def weatherformer_mixture_training_loop(args_dict):
    return _former_family_loop(WeatherFormerMixture, WeatherFormerMixtureTrainer, args_dict)
