"""County-yield data for the fine-tune path (reference src/crop_yield/dataloader/yield_dataloader.py).

Same sample semantics as the reference CropDataset (:114-302): one sample = (n_past_years + 1) consecutive yearly
records of a county ending at the target year; 6 weekly weather series [52] are scattered into feature slots
[7, 8, 11, 1, 2, 29] of the 31-feature encoder input (the other 25 features are masked and imputed by the encoder),
sequence length (n_past_years + 1) * 52 <= 365, year[t] = year + week/52, interval 7, the target year's yield is
replaced by the previous year's in y_past. Standardisation and the train/test split by year follow :314-409.
The reference builds samples with a per-row pandas `apply`; here the history check and the tensor assembly are
grouped by county (same result, no quadratic scan)."""
import json
import logging
import os
from typing import List, Tuple

import numpy as np
import pandas as pd
import torch
from torch.utils.data import DataLoader, Dataset

from ...utils.constants import CROP_YIELD_STATS, DATA_DIR, DRY_RUN, MAX_CONTEXT_LENGTH, TOTAL_WEATHER_VARS

logger = logging.getLogger(__name__)

WEATHER_COLS = [f"W_{i}_{j}" for i in range(1, 7) for j in range(1, 53)]
PRACTICE_COLS = [f"P_{i}" for i in range(1, 15)]
_SOIL = ["bdod", "cec", "cfvo", "clay", "nitrogen", "ocd", "ocs", "phh2o", "sand", "silt", "soc"]
_DEPTHS = ["0-5cm", "5-15cm", "15-30cm", "30-60cm", "60-100cm", "100-200cm"]
SOIL_COLS = [f"{m}_mean_{d}" for m in _SOIL for d in _DEPTHS]
# encoder feature slots of the six weekly series: precipitation, solar radiation, snow depth, T max, T min, vapour pressure
WEATHER_INDICES = [7, 8, 11, 1, 2, 29]


class CropDataset(Dataset):
    def __init__(self, data, start_year, test_year, test_dataset=False, n_past_years=5, test_gap=0, crop_type="soybean"):
        self.crop_type = crop_type
        self.yield_col = f"{crop_type}_yield"
        self.weather_cols, self.practice_cols, self.soil_cols = WEATHER_COLS, PRACTICE_COLS, SOIL_COLS
        self.weather_indices = torch.tensor(WEATHER_INDICES)
        start_year -= test_gap
        n_hist = n_past_years + 1
        if n_hist * 52 > MAX_CONTEXT_LENGTH:
            raise ValueError(f"n_years * seq_len = {n_hist * 52} is greater than MAX_CONTEXT_LENGTH = {MAX_CONTEXT_LENGTH}")

        data = data.sort_values(["loc_ID", "year"]).reset_index(drop=True)
        if test_dataset:
            wanted = data["year"] == test_year
        else:
            wanted = (data["year"] >= start_year) & (data["year"] < test_year - test_gap)
        # position of every record within its county's chronological list: a target needs n_hist records up to itself
        pos = data.groupby("loc_ID").cumcount()
        cand = data.index[wanted & (pos >= n_hist - 1)]
        # the reference keeps candidates in the order of the (unsorted-by-us) frame it was given; that frame is
        # sorted by (loc_ID, year) by read_usa_dataset already, so the order is identical
        name = "test" if test_dataset else "train"
        logger.info(f"Creating {name} dataloader with {len(cand)} samples using {crop_type} yield.")
        self.index = data.loc[cand, ["year", "loc_ID"]].reset_index(drop=True)
        n_use = len(cand) // 20 if DRY_RUN else len(cand)
        self.data: List[Tuple[torch.Tensor, ...]] = []
        if len(cand) == 0:
            logger.warning(f"No samples found for {name} dataset!")
            return

        w_all = data[WEATHER_COLS].to_numpy(dtype=np.float32)
        p_all = data[PRACTICE_COLS].to_numpy(dtype=np.float32) if set(PRACTICE_COLS) <= set(data.columns) else None
        s_all = data[SOIL_COLS].to_numpy(dtype=np.float32) if set(SOIL_COLS) <= set(data.columns) else None
        yr_all = data["year"].to_numpy(dtype=np.float32)
        ll_all = data[["lat", "lng"]].to_numpy(dtype=np.float32)
        y_all = data[self.yield_col].to_numpy(dtype=np.float32)
        week = torch.arange(1, 53, dtype=torch.float32) / 52
        mask_row = torch.ones(TOTAL_WEATHER_VARS, dtype=torch.bool)
        mask_row[self.weather_indices] = False
        for end in cand[:n_use]:
            rows = slice(end - n_hist + 1, end + 1)  # consecutive records of the same county (sorted frame)
            weather = torch.from_numpy(w_all[rows].reshape(n_hist, 6, 52).transpose(0, 2, 1).reshape(n_hist * 52, 6).copy())
            padded = torch.zeros((n_hist * 52, TOTAL_WEATHER_VARS))
            padded[:, self.weather_indices] = weather
            year_expanded = (torch.from_numpy(yr_all[rows].copy()).unsqueeze(1) + week.unsqueeze(0)).reshape(n_hist * 52)
            y_past = y_all[rows].copy()
            target = y_past[-1:].copy()
            y_past[-1] = y_past[-2]  # the target year's own yield is unknown at prediction time
            practices = p_all[rows].reshape(n_hist, 14).copy() if p_all is not None else np.zeros((n_hist, 14), np.float32)
            soil = s_all[rows].reshape(n_hist, 11, 6).copy() if s_all is not None else np.zeros((n_hist, 11, 6), np.float32)
            self.data.append((padded, torch.from_numpy(ll_all[end - n_hist + 1].copy()), year_expanded,
                              torch.full((1,), 7, dtype=torch.float32), mask_row.unsqueeze(0).expand(n_hist * 52, -1),
                              practices, soil, y_past, target))

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        return self.data[idx]

    def get_data_loader(self, batch_size=32, shuffle=False, num_workers=4):
        return DataLoader(self, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers, pin_memory=True)


def load_weather_scalers_from_json(json_path: str):
    """weekly per-parameter scalers -> {W_<slot>_<week>: {mean, std}} (reference :20-62)"""
    slot = {"T2M_MAX": 1, "T2M_MIN": 2, "PRECTOTCORR": 7, "ALLSKY_SFC_SW_DWN": 8, "SNODP": 11, "VAP": 29}
    with open(json_path) as f:
        raw = json.load(f)
    out = {}
    for key, wk in raw.items():
        if not key.startswith("week_"):
            continue
        w = int(key.split("_")[1])
        for pname, s in slot.items():
            if pname in wk["param_means"] and pname in wk["param_stds"]:
                out[f"W_{s}_{w}"] = {"mean": wk["param_means"][pname], "std": wk["param_stds"][pname]}
    return out


def standardize_weather_cols(data: pd.DataFrame, country: str) -> pd.DataFrame:
    out = data.copy()
    cols = [c for c in WEATHER_COLS if c in out.columns]
    if country.lower() != "usa":
        path = os.path.join(DATA_DIR, "khaki_soybeans", "weekly_weather_param_scalers.json")
        if not os.path.exists(path):
            raise FileNotFoundError(f"JSON scalers file not found at {path}")
        scalers = load_weather_scalers_from_json(path)
        for c in cols:
            sc = scalers.get(c)
            if sc:
                out[c] = (out[c] - sc["mean"]) / sc["std"] if sc["std"] > 0 else 0
    elif cols:
        out[cols] = ((out[cols] - out[cols].mean()) / out[cols].std()).fillna(0)
    return out


def split_train_test_by_year(soybean_df: pd.DataFrame, n_train_years: int, test_year: int, standardize: bool,
                             n_past_years: int, crop_type: str, country: str, test_gap: int = 0):
    start_year = test_year - n_train_years
    yield_col = f"{crop_type}_yield"
    data = soybean_df[soybean_df["year"] > 1981.0].copy()
    before = len(data)
    data = data.dropna(subset=[yield_col])
    if len(data) < before:
        logger.warning(f"Dropped {before - len(data)} rows with missing {crop_type} yield values")
    data = data.fillna(0)
    if standardize:
        data = standardize_weather_cols(data, country)
        keep = {"loc_ID", "year", "State", "County", "lat", "lng", yield_col, *WEATHER_COLS}
        others = [c for c in data.columns if c not in keep]
        if others:
            data[others] = ((data[others] - data[others].mean()) / data[others].std()).fillna(0)
        train_rows = data[(data["year"] >= start_year) & (data["year"] < test_year)]
        mean, std = train_rows[yield_col].mean(), train_rows[yield_col].std()
        data[yield_col] = (data[yield_col] - mean) / std
        logger.info(f"Saving mean ({mean:.3f}) and std ({std:.3f}) from training data for {crop_type}")
        CROP_YIELD_STATS[crop_type]["mean"].append(mean)
        CROP_YIELD_STATS[crop_type]["std"].append(std)
    mk = lambda is_test: CropDataset(data.copy(), start_year, test_year, test_dataset=is_test,  # noqa: E731
                                     n_past_years=n_past_years, test_gap=test_gap, crop_type=crop_type)
    return mk(False), mk(True)


def read_usa_dataset(data_dir: str):
    return pd.read_csv(data_dir + "khaki_soybeans/khaki_multi_crop_yield.csv").sort_values(["loc_ID", "year"])


def read_non_us_dataset(data_dir: str, country: str):
    df = pd.read_csv(data_dir + f"khaki_soybeans/khaki_{country}_multi_crop.csv")
    if country == "brazil":
        df = df[df["State"].isin(["Goiás", "Mato Grosso", "Mato Grosso do Sul", "Paraná", "Rio Grande do Sul"])].copy()
    return df.sort_values(["loc_ID", "year"])


def get_train_test_loaders(crop_df: pd.DataFrame, n_train_years: int, test_year: int, n_past_years: int,
                           batch_size: int, shuffle: bool, num_workers: int, crop_type: str, country: str,
                           test_gap: int = 0) -> Tuple[DataLoader, DataLoader]:
    if n_train_years <= 1:
        raise ValueError(f"Not enough training data for current year + n_past_years. Required: {n_past_years + 1}. "
                         f"Available training years: {n_train_years}.")
    if n_train_years < n_past_years + 1:
        logger.warning(f"Setting n_past_years to {n_train_years - 1} (only {n_train_years} training years).")
        n_past_years = n_train_years - 1
    train_ds, test_ds = split_train_test_by_year(crop_df, n_train_years, test_year, standardize=True,
                                                 n_past_years=n_past_years, crop_type=crop_type, country=country,
                                                 test_gap=test_gap)
    if n_past_years < 1:
        raise ValueError("Not enough training data for current year + n_past_years.")
    return (train_ds.get_data_loader(batch_size, shuffle=shuffle, num_workers=num_workers),
            test_ds.get_data_loader(batch_size, shuffle=shuffle, num_workers=num_workers))
