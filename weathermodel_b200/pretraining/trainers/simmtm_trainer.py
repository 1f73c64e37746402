"""SimMTMTrainer: WeatherBertTrainer's MSE under contiguous-segment masks
(reference src/pretraining/trainers/simmtm_trainer.py:8-68)."""
from ..models.simmtm import SimMTM
from .weatherautoencoder_trainer import _bert_family_loop
from .weatherbert_trainer import WeatherBertTrainer


class SimMTMTrainer(WeatherBertTrainer):
    def __init__(self, model: SimMTM, masking_prob: float, n_masked_features: int, **kwargs):
        super().__init__(model, masking_prob, n_masked_features, **kwargs)
        self.masking_function = "simmtm"


def simmtm_training_loop(args_dict):
    return _bert_family_loop(SimMTM, SimMTMTrainer, args_dict)
