"""WeatherFormerSinusoidTrainer: ELBO whose KL is taken against the learned sinusoidal prior
(reference src/pretraining/trainers/weatherformer_sinusoid_trainer.py:11-91)."""
import logging

import torch

from ...utils.constants import TOTAL_WEATHER_VARS
from ...utils.losses import compute_gaussian_kl_divergence
from ..models.weatherformer_sinusoid import WeatherFormerSinusoid
from .weatherformer_trainer import WeatherFormerTrainer


class WeatherFormerSinusoidTrainer(WeatherFormerTrainer):
    def __init__(self, model: WeatherFormerSinusoid, masking_prob: float, n_masked_features: int, beta: float,
                 **kwargs):
        super().__init__(model=model, masking_prob=masking_prob, n_masked_features=n_masked_features, beta=beta,
                         **kwargs)
        self.output_json["model_config"]["n_mixture_components"] = model.k

    def compute_kl_loss(self, weather, weather_feature_mask, mu_x, var_x, mu_p, var_p) -> torch.Tensor:
        return compute_gaussian_kl_divergence(weather_feature_mask, mu_x, var_x, mu_p, var_p)


def _former_family_loop(model_cls, trainer_cls, args_dict):
    rank, world_size, local_rank = (args_dict.get(k, d) for k, d in (("rank", 0), ("world_size", 1), ("local_rank", 0)))
    device = torch.device(f"cuda:{local_rank}" if torch.cuda.is_available() else "cpu")
    model = model_cls(weather_dim=TOTAL_WEATHER_VARS, output_dim=TOTAL_WEATHER_VARS,
                      k=args_dict["n_mixture_components"], device=device, **args_dict["model_size_params"]).to(device)
    if rank == 0:
        logging.info(str(model))
    trainer = trainer_cls(
        model=model, batch_size=args_dict["batch_size"], num_epochs=args_dict["n_epochs"],
        init_lr=args_dict["init_lr"], num_warmup_epochs=args_dict["n_warmup_epochs"],
        decay_factor=args_dict["decay_factor"], pretrained_model_path=args_dict["pretrained_model_path"],
        masking_prob=args_dict["masking_prob"], n_masked_features=args_dict["n_masked_features"],
        beta=args_dict["beta"], resume_from_checkpoint=args_dict.get("resume_from_checkpoint"), rank=rank,
        world_size=world_size, local_rank=local_rank)
    return trainer.train(use_optimal_lr=args_dict["use_optimal_lr"])


def weatherformer_sinusoid_training_loop(args_dict):
    return _former_family_loop(WeatherFormerSinusoid, WeatherFormerSinusoidTrainer, args_dict)
