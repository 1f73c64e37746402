"""WeatherBertTrainer: masked-feature pretraining (reference src/pretraining/trainers/weatherbert_trainer.py).

Loss = MSE over the masked elements of the batch (the reference's nn.MSELoss over boolean gathers, :55-60 --
MSE, not MAE), computed by the fused wm_loss_bert kernel without the gathers' host sync."""
import logging
import random
from typing import Dict, Tuple

import torch
from torch.utils.data import DataLoader

from ...base_trainer.base_trainer import BaseTrainer
from ...engine import bert_masked_mse
from ...utils.constants import TOTAL_WEATHER_VARS
from ..dataloader.pretraining_dataloader import streaming_dataloader
from ..models.weatherbert import WeatherBERT

logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s")
random.seed(1234)  # the reference reseeds at import of its trainers
torch.manual_seed(1234)


class WeatherBertTrainer(BaseTrainer):
    _graph_capturable = True  # loss stays on the device end to end

    def __init__(self, model: WeatherBERT, masking_prob: float, n_masked_features: int, **kwargs):
        super().__init__(model, **kwargs)
        self.masking_function = "weatherbert"
        self.masking_prob = masking_prob
        self.n_masked_features = n_masked_features
        cfg = self.output_json["model_config"]
        cfg["masking_function"], cfg["masking_prob"], cfg["n_masked_features"] = "weatherbert", masking_prob, n_masked_features

    def _loss(self, data, coords, year, interval, feature_mask) -> Dict[str, torch.Tensor]:
        y_pad = self.model.forward_raw(data, coords, year, interval, feature_mask)  # (the data-parallel wrapper forwards it)
        return {"total_loss": bert_masked_mse(y_pad, data, feature_mask)}

    def compute_train_loss(self, data, coords, year, interval, feature_mask) -> Dict[str, torch.Tensor]:
        return self._loss(data, coords, year, interval, feature_mask)

    def compute_validation_loss(self, data, coords, year, interval, feature_mask) -> Dict[str, torch.Tensor]:
        return self._loss(data, coords, year, interval, feature_mask)

    def get_dataloaders(self, shuffle: bool = True) -> Tuple[DataLoader, DataLoader]:
        n_masked = self._get_n_masked_features(self.current_epoch, self.n_masked_features)
        common = dict(masking_function=self.masking_function, masking_prob=self.masking_prob,
                      n_masked_features=n_masked, world_size=self.world_size, rank=self.rank)
        return (streaming_dataloader(self.batch_size, split="train", shuffle=shuffle, **common),
                streaming_dataloader(self.batch_size, split="validation", shuffle=False, **common))


def weatherbert_training_loop(args_dict):
    rank, world_size, local_rank = (args_dict.get(k, d) for k, d in (("rank", 0), ("world_size", 1), ("local_rank", 0)))
    device = torch.device(f"cuda:{local_rank}" if torch.cuda.is_available() else "cpu")
    model = WeatherBERT(weather_dim=TOTAL_WEATHER_VARS, output_dim=TOTAL_WEATHER_VARS, device=device,
                        **args_dict["model_size_params"]).to(device)
    if rank == 0:
        logging.info(str(model))
    trainer = WeatherBertTrainer(
        model=model, batch_size=args_dict["batch_size"], num_epochs=args_dict["n_epochs"],
        init_lr=args_dict["init_lr"], num_warmup_epochs=args_dict["n_warmup_epochs"],
        decay_factor=args_dict["decay_factor"], pretrained_model_path=args_dict["pretrained_model_path"],
        masking_prob=args_dict["masking_prob"], n_masked_features=args_dict["n_masked_features"],
        resume_from_checkpoint=args_dict["resume_from_checkpoint"], rank=rank, world_size=world_size,
        local_rank=local_rank)
    return trainer.train(use_optimal_lr=args_dict["use_optimal_lr"])
