"""BaseModel with the reference's interface (src/base_models/base_model.py:10-33)."""
import logging
from abc import abstractmethod

import torch.nn as nn


class BaseModel(nn.Module):
    def __init__(self, name: str):
        super().__init__()
        self.name = name
        self.logger = logging.getLogger(f"models.{name}")
        logging.info(f"Initializing {self.name} model")

    def total_params(self) -> int:
        return sum(p.numel() for p in self.parameters())

    def total_params_formatted(self) -> str:
        n = self.total_params()
        return f"{n / 10**6:.1f}m" if n > 10**6 else f"{n / 10**3:.1f}k"

    @abstractmethod
    def load_pretrained(self, pretrained_model: "BaseModel"):
        """copy weights of a pretrained model into this one"""
