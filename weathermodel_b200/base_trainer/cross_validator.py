"""CrossValidator: k folds, each a freshly seeded model + trainer (reference src/base_trainer/cross_validator.py:46-116).
Every fold reseeds Python / numpy / torch to 1234 and switches torch to deterministic algorithms; the B200 kernels
themselves are deterministic (fixed-order split-K and reductions, no atomics)."""
import logging
import os
import random
from typing import Any, Dict, List, Type

import numpy as np
import torch

from ..base_models.base_model import BaseModel
from .base_trainer import BaseTrainer


class CrossValidator:
    def __init__(self, model_class: Type[BaseModel], model_kwargs: Dict[str, Any], trainer_class: Type[BaseTrainer],
                 trainer_kwargs: Dict[str, Any], k_folds: int = 5):
        self.model_class, self.model_kwargs = model_class, model_kwargs
        self.trainer_class, self.trainer_kwargs = trainer_class, trainer_kwargs
        self.k_folds = k_folds
        self.logger = logging.getLogger(__name__)

    @staticmethod
    def _reseed():
        os.environ["CUBLAS_WORKSPACE_CONFIG"] = ":4096:8"
        random.seed(1234)
        np.random.seed(1234)
        torch.manual_seed(1234)
        torch.cuda.manual_seed(1234)
        torch.use_deterministic_algorithms(True)

    def run_cross_validation(self, use_optimal_lr: bool = True) -> Dict[str, Any]:
        self.logger.info(f"Starting {self.k_folds}-fold cross validation")
        fold_results: List[float] = []
        for fold in range(self.k_folds):
            self.logger.info(f"Starting fold {fold + 1}/{self.k_folds}")
            self._reseed()
            model = self.model_class(**self.model_kwargs)
            if fold == 0:
                self.logger.info(str(model))
            trainer = self.trainer_class(model=model, **self.trainer_kwargs)
            best = trainer.train(use_optimal_lr=use_optimal_lr)
            fold_results.append(best)
            self.logger.info(f"Fold [{fold + 1} / {self.k_folds}] completed. Best val loss: {best:.4f}")
        return self._aggregate_results(fold_results, float(sum(fold_results)))

    def _aggregate_results(self, fold_results: List[float], total_best_val_loss: float) -> Dict[str, Any]:
        return {"avg_best_val_loss": np.mean(fold_results), "std_best_val_loss": np.std(fold_results),
                "fold_results": fold_results, "n_folds": len(fold_results)}
