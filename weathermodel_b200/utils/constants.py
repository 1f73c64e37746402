"""Constants of the reference's src/utils/constants.py (same names and values; env switches DRY_RUN / STDOUT)."""
import os

import torch

try:  # the reference loads a .env file if python-dotenv is present
    from dotenv import load_dotenv

    load_dotenv()
except Exception:  # pragma: no cover - optional dependency
    pass


def _env_flag(name: str) -> bool:
    return os.environ.get(name, "False").lower() in ("true", "1", "t")


DATA_DIR = "data/"
WEATHER_FILE_PATH = DATA_DIR + "nasa_power/train_dataset_weekly.pth"
DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")
STDOUT = _env_flag("STDOUT")
DRY_RUN = _env_flag("DRY_RUN")

_CROPS = ("soybean", "corn", "wheat", "sunflower", "cotton", "sugarcane", "beans")
CROP_YIELD_STATS = {crop: {"mean": [], "std": []} for crop in _CROPS}

TOTAL_WEATHER_VARS = 31
MAX_GRANULARITY_DAYS = 31
MAX_CONTEXT_LENGTH = 365
NUM_DATASET_PARTS = 119
VALIDATION_CHUNK_IDS = [7, 30, 56, 59, 93, 106, 110, 24]
DRY_RUN_TRAIN_CHUNK_IDS = [1, 34, 53, 72, 81]
