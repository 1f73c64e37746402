"""Crop-yield fine-tune CLI with the reference's flags and defaults (src/crop_yield/yield_main.py:12-253):

    python -m src.crop_yield.yield_main --model weatherformer --model-size mini --n-past-years 6

Supported --model values are the two encoder families of the B200 hot path; the reference's other yield baselines
(CNN-RNN, GNN-RNN, linear, Chronos, ablations) are out of scope and rejected by name (the reference imports all of
them unconditionally, which fails without dgl / chronos installed)."""
import argparse
import logging
import os
import random

import numpy as np
import torch

from ..utils.constants import CROP_YIELD_STATS
from ..utils.utils import parse_args, setup_logging

_CROPS = ["soybean", "corn", "wheat", "sunflower", "cotton", "sugarcane", "beans"]
parser = argparse.ArgumentParser()
for _flag, _kw in [
    ("--model", dict(default="weatherformer", type=str, help="weatherformer or weatherbert")),
    ("--batch-size", dict(default=64, type=int, help="batch size")),
    ("--n-past-years", dict(default=6, type=int, help="number of past years to look at")),
    ("--n-epochs", dict(default=40, type=int, help="number of training epochs")),
    ("--init-lr", dict(default=0.0005, type=float, help="initial learning rate for Adam")),
    ("--decay_factor", dict(default=None, type=float, help="learning rate exponential decay factor")),
    ("--n-warmup-epochs", dict(default=10, type=int, help="number of warmup epochs")),
    ("--pretrained-model-path", dict(default=None, type=str, help="path to pretrained model weights")),
    ("--model-size", dict(default="small", type=str, help="mini, small, medium or large")),
    ("--n-train-years", dict(default=5, type=int, help="number of years of training data to use")),
    ("--beta", dict(default=1e-4, type=float, help="weight of the KL term (WeatherFormer)")),
    ("--use-optimal-lr", dict(action="store_true", default=False, help="run the LR range test first")),
    ("--seed", dict(default=1234, type=int, help="seed for random number generators")),
    ("--n-mixture-components", dict(default=1, type=int, help="accepted for CLI compatibility; unused here")),
    ("--crop-type", dict(default="soybean", type=str, choices=_CROPS, help="crop to predict")),
    ("--country", dict(default="usa", type=str, choices=["usa", "argentina", "brazil"], help="dataset")),
    ("--test-year", dict(default=None, type=int, help="single test year instead of 5-fold cross validation")),
    ("--test-type", dict(default="extreme", type=str, choices=["extreme", "overall", "ahead_pred"], help="fold years")),
]:
    parser.add_argument(_flag, **_kw)


def main(args_dict=None):
    setup_logging(rank=0)
    if args_dict is None:
        args_dict = parse_args(parser)
    seed = args_dict["seed"]
    os.environ["CUBLAS_WORKSPACE_CONFIG"] = ":4096:8"
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    torch.use_deterministic_algorithms(True)
    if args_dict["n_train_years"] < args_dict["n_past_years"] + 1:
        logging.warning(f"Setting n_past_years to {args_dict['n_train_years'] - 1} "
                        f"(only {args_dict['n_train_years']} training years).")
        args_dict["n_past_years"] = args_dict["n_train_years"] - 1
    kind = args_dict["model"].lower()
    if kind == "weatherformer":
        from .trainers.weatherformer_yield_trainer import weatherformer_yield_training_loop as loop
    elif kind == "weatherbert":
        from .trainers.weatherbert_yield_trainer import weatherbert_yield_training_loop as loop
    else:
        raise ValueError(f"Unknown model type: {kind}. The B200 hot path implements 'weatherbert' and 'weatherformer'.")
    results = loop(args_dict, use_cropnet=False)
    log = logging.getLogger(__name__)
    log.info("Training completed successfully!")
    stds = CROP_YIELD_STATS[args_dict["crop_type"]]["std"]
    rmse = [r * s for r, s in zip(results["fold_results"], stds)]  # back to bu/acre
    r2 = [1 - (e / s) ** 2 for e, s in zip(rmse, stds)]
    avg_rmse, std_rmse, avg_r2, std_r2 = float(np.mean(rmse)), float(np.std(rmse)), float(np.mean(r2)), float(np.std(r2))
    log.info(f"Final average best RMSE for {args_dict['crop_type']}: {avg_rmse:.3f} ± {std_rmse:.3f}")
    log.info(f"Final average R² for {args_dict['crop_type']}: {avg_r2:.3f} ± {std_r2:.3f}")
    return avg_rmse, std_rmse, avg_r2, std_r2, r2


if __name__ == "__main__":
    try:
        main()
    except Exception as e:  # noqa: BLE001
        logging.getLogger(__name__).error(f"Training failed with error: {e}")
        raise
