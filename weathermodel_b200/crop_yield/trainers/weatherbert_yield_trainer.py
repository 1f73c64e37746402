"""WeatherBERTYieldTrainer + the shared yield-training plumbing
(reference src/crop_yield/trainers/weatherbert_yield_trainer.py). Loss: MSE on the standardised yield; validation
reports RMSE. Folds test on the reference's fixed year lists."""
import os
from typing import Dict, Optional, Tuple

import pandas as pd
import torch
import torch.nn as nn
from torch.utils.data import DataLoader

from ...base_trainer.base_trainer import BaseTrainer
from ...base_trainer.cross_validator import CrossValidator
from ...utils.constants import DATA_DIR, TOTAL_WEATHER_VARS
from ..dataloader.yield_dataloader import get_train_test_loaders, read_non_us_dataset, read_usa_dataset
from ..models.weatherbert_yield_model import WeatherBERTYieldModel

TEST_YEARS = [2014, 2015, 2016, 2017, 2018]
FOLD_IDX = 0
EXTREME_YEARS = {
    "usa": {"corn": [2002, 2004, 2009, 2012, 2014], "soybean": [2003, 2004, 2009, 2012, 2016]},
    "argentina": {"corn": [2004, 2005, 2007, 2009, 2015], "soybean": [2003, 2006, 2007, 2009, 2015],
                  "wheat": [2002, 2003, 2005, 2009, 2011], "sunflower": [2002, 2007, 2008, 2009, 2011]},
    "brazil": {"corn": [2001, 2003, 2007, 2010, 2015], "soybean": [2001, 2003, 2005, 2011, 2017],
               "sugarcane": [2002, 2003, 2008, 2012, 2017], "wheat": [2001, 2003, 2010, 2015, 2016],
               "cotton": [2004, 2008, 2013, 2017, 2018]},
    "mexico": {"beans": [2016, 2017, 2018, 2021, 2023], "corn": [2014, 2017, 2019, 2022, 2023],
               "sugarcane": [2013, 2014, 2018, 2020, 2021], "wheat": [2013, 2021, 2022, 2023, 2024]},
}


def _reset_fold_index():
    global FOLD_IDX
    FOLD_IDX = 0


class WeatherBERTYieldTrainer(BaseTrainer):
    _graph_capturable = True  # MSE (+ beta * KL) in device-side torch ops; the fused head is one kernel each way
    def __init__(self, crop_df: pd.DataFrame, country: str, n_past_years: int, n_train_years: int, beta: float,
                 use_cropnet: bool, crop_type: str, test_year: Optional[int] = None, test_type: str = "extreme",
                 **kwargs):
        super().__init__(**kwargs)
        self.crop_df, self.country, self.crop_type = crop_df, country, crop_type
        self.n_past_years, self.n_train_years, self.beta = n_past_years, n_train_years, beta
        self.use_cropnet, self.test_type = use_cropnet, test_type
        self.output_json["model_config"]["beta"] = beta
        self.criterion = nn.MSELoss(reduction="mean")
        if use_cropnet:
            raise NotImplementedError("the CropNet loader is outside the B200 hot-path scope (SURVEY.md 2.1)")
        if test_type == "extreme":
            years = EXTREME_YEARS.get(country, {}).get(crop_type)
            if years is None:
                raise ValueError(f"No extreme years found for {crop_type} in {country}.")
        elif test_type in ("overall", "ahead_pred"):
            years = TEST_YEARS
        else:
            raise ValueError(f"Unknown test_type: {test_type}. Choose 'extreme', 'overall', or 'ahead_pred'.")
        self.logger.info(f"Test type: {test_type} on years {years}")
        if self.rank == 0:
            self.model_dir = DATA_DIR + "trained_models/crop_yield/"
            os.makedirs(self.model_dir, exist_ok=True)
        if test_year is not None:
            self.test_year = test_year
        else:
            global FOLD_IDX
            if FOLD_IDX >= len(years):
                raise ValueError(f"FOLD_IDX ({FOLD_IDX}) exceeds TEST_YEARS length ({len(years)}). "
                                 "Call _reset_fold_index() before starting new cross-validation.")
            self.test_year = years[FOLD_IDX]
            FOLD_IDX += 1
        self.logger.info(f"Testing on year: {self.test_year}")
        self.train_loader: Optional[DataLoader] = None
        self.test_loader: Optional[DataLoader] = None

    def get_dataloaders(self, shuffle: bool = False) -> Tuple[DataLoader, DataLoader]:
        if self.train_loader is None or self.test_loader is None:
            self.train_loader, self.test_loader = get_train_test_loaders(
                self.crop_df, self.n_train_years, self.test_year, self.n_past_years, self.batch_size, shuffle,
                num_workers=0, crop_type=self.crop_type, country=self.country,
                test_gap=4 if self.test_type == "ahead_pred" else 0)
        return self.train_loader, self.test_loader

    def _mse(self, padded_weather, coord, year, interval, mask, y_past, target):
        pred = self.model(padded_weather, coord, year, interval, mask, y_past)
        return self.criterion(pred.squeeze(), target.squeeze())

    def compute_train_loss(self, padded_weather, coord_processed, year_expanded, interval, weather_feature_mask,
                           practices, soil, y_past, target_yield) -> Dict[str, torch.Tensor]:
        return {"total_loss": self._mse(padded_weather, coord_processed, year_expanded, interval, weather_feature_mask,
                                        y_past, target_yield)}

    def compute_validation_loss(self, padded_weather, coord_processed, year_expanded, interval, weather_feature_mask,
                                practices, soil, y_past, target_yield) -> Dict[str, torch.Tensor]:
        with torch.no_grad():
            mse = self._mse(padded_weather, coord_processed, year_expanded, interval, weather_feature_mask, y_past,
                            target_yield)
        return {"total_loss": mse ** 0.5}  # RMSE is what the paper tables compare

    def _current_beta(self):
        return self.beta


def _create_yield_training_setup(args_dict, use_cropnet: bool):
    if use_cropnet:
        raise NotImplementedError("the CropNet loader is outside the B200 hot-path scope (SURVEY.md 2.1)")
    country = args_dict["country"]
    crop_df = read_usa_dataset(DATA_DIR) if country == "usa" else read_non_us_dataset(DATA_DIR, country)
    return {
        "rank": args_dict.get("rank", 0), "world_size": args_dict.get("world_size", 1),
        "local_rank": args_dict.get("local_rank", 0),
        "device": torch.device("cuda" if torch.cuda.is_available() else "cpu"),
        "crop_df": crop_df,
        "cross_validation_k": 1 if args_dict.get("test_year") is not None else len(TEST_YEARS),
        "beta": args_dict["beta"], "use_cropnet": use_cropnet, "test_year": args_dict.get("test_year"),
        "test_type": args_dict.get("test_type", "extreme"),
    }


def _run_yield_cross_validation(setup_params, model_class, trainer_class, model_name, args_dict,
                                extra_trainer_kwargs=None, extra_model_kwargs=None):
    if not setup_params["use_cropnet"] and setup_params["test_year"] is None:
        _reset_fold_index()
    model_kwargs = {"name": model_name, "device": setup_params["device"], "weather_dim": TOTAL_WEATHER_VARS,
                    "n_past_years": args_dict["n_past_years"], **args_dict["model_size_params"]}
    model_kwargs.update(extra_model_kwargs or {})
    trainer_kwargs = {
        "crop_df": setup_params["crop_df"], "country": args_dict["country"], "n_past_years": args_dict["n_past_years"],
        "n_train_years": args_dict["n_train_years"], "beta": args_dict["beta"], "use_cropnet": setup_params["use_cropnet"],
        "crop_type": args_dict["crop_type"], "test_year": setup_params["test_year"], "test_type": setup_params["test_type"],
        "batch_size": args_dict["batch_size"], "num_epochs": args_dict["n_epochs"], "init_lr": args_dict["init_lr"],
        "num_warmup_epochs": args_dict["n_warmup_epochs"], "decay_factor": args_dict["decay_factor"],
        "pretrained_model_path": args_dict["pretrained_model_path"],
        "resume_from_checkpoint": args_dict.get("resume_from_checkpoint"), "rank": setup_params["rank"],
        "world_size": setup_params["world_size"], "local_rank": setup_params["local_rank"],
    }
    trainer_kwargs.update(extra_trainer_kwargs or {})
    cv = CrossValidator(model_class=model_class, model_kwargs=model_kwargs, trainer_class=trainer_class,
                        trainer_kwargs=trainer_kwargs, k_folds=setup_params["cross_validation_k"])
    return cv.run_cross_validation(use_optimal_lr=args_dict["use_optimal_lr"])


def weatherbert_yield_training_loop(args_dict, use_cropnet: bool):
    return _run_yield_cross_validation(
        setup_params=_create_yield_training_setup(args_dict, use_cropnet), model_class=WeatherBERTYieldModel,
        trainer_class=WeatherBERTYieldTrainer, model_name=f"weatherbert_{args_dict['crop_type']}_yield",
        args_dict=args_dict)
