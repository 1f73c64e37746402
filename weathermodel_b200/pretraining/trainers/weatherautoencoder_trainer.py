"""WeatherAutoencoderTrainer: WeatherBertTrainer's MSE with WeatherFormer's whole-feature masks
(reference src/pretraining/trainers/weatherautoencoder_trainer.py:8-68)."""
import logging

import torch

from ...utils.constants import TOTAL_WEATHER_VARS
from ..models.weatherautoencoder import WeatherAutoencoder
from .weatherbert_trainer import WeatherBertTrainer


class WeatherAutoencoderTrainer(WeatherBertTrainer):
    def __init__(self, model: WeatherAutoencoder, masking_prob: float, n_masked_features: int, **kwargs):
        super().__init__(model, masking_prob, n_masked_features, **kwargs)
        self.masking_function = "weatherformer"


def _bert_family_loop(model_cls, trainer_cls, args_dict):
    rank, world_size, local_rank = (args_dict.get(k, d) for k, d in (("rank", 0), ("world_size", 1), ("local_rank", 0)))
    device = torch.device(f"cuda:{local_rank}" if torch.cuda.is_available() else "cpu")
    model = model_cls(weather_dim=TOTAL_WEATHER_VARS, output_dim=TOTAL_WEATHER_VARS, device=device,
                      **args_dict["model_size_params"]).to(device)
    if rank == 0:
        logging.info(str(model))
    trainer = trainer_cls(
        model=model, batch_size=args_dict["batch_size"], num_epochs=args_dict["n_epochs"],
        init_lr=args_dict["init_lr"], num_warmup_epochs=args_dict["n_warmup_epochs"],
        decay_factor=args_dict["decay_factor"], pretrained_model_path=args_dict["pretrained_model_path"],
        masking_prob=args_dict["masking_prob"], n_masked_features=args_dict["n_masked_features"],
        resume_from_checkpoint=args_dict.get("resume_from_checkpoint"), rank=rank, world_size=world_size,
        local_rank=local_rank)
    return trainer.train(use_optimal_lr=args_dict["use_optimal_lr"])


def weatherautoencoder_training_loop(args_dict):
    return _bert_family_loop(WeatherAutoencoder, WeatherAutoencoderTrainer, args_dict)
