"""Host utilities with the reference's names and behaviour (src/utils/utils.py): LR schedule,
input normalisation, torchrun bootstrap, model-size table, argument parsing."""
import logging
import math
import os
from argparse import ArgumentParser

import torch
import torch.distributed as dist
import torch.optim as optim

_SIZES = {
    "mini": (4, 2, 12),
    "small": (10, 4, 20),
    "medium": (12, 6, 28),
    "large": (16, 8, 36),
}


def get_model_params(model_size: str):
    """reference src/utils/utils.py:112-123"""
    try:
        heads, layers, factor = _SIZES[model_size]
    except KeyError:
        raise ValueError(f"Unknown model size: {model_size}") from None
    return {"num_heads": heads, "num_layers": layers, "hidden_dim_factor": factor}


def lr_multiplier(epoch, num_warmup_epochs, total_epochs, decay_factor=None):
    """LambdaLR factor of reference src/utils/utils.py:11-60: linear warm-up (0 at epoch 0), then cosine
    (decay_factor None) or exponential decay."""
    if epoch < num_warmup_epochs:
        return float(epoch) / float(max(1, num_warmup_epochs))
    done = epoch - num_warmup_epochs
    if decay_factor is None:
        return 0.5 * (1.0 + math.cos(math.pi * done / (total_epochs - num_warmup_epochs)))
    return decay_factor ** done


def get_scheduler(optimizer, num_warmup_epochs, total_epochs, decay_factor=None):
    return optim.lr_scheduler.LambdaLR(
        optimizer, lambda e: lr_multiplier(e, num_warmup_epochs, total_epochs, decay_factor))


def normalize_year_interval_coords(year, interval, coords):
    """reference src/utils/utils.py:63-74 (host-side version; the CUDA embedding kernel does the same
    arithmetic in-register). Inputs are not modified."""
    scaled = coords.clone()
    scaled[:, 0] = scaled[:, 0] / 360
    scaled[:, 1] = scaled[:, 1] / 180
    return (year - 1970) / 100.0, interval / 30.0, scaled


def setup_distributed():
    """torchrun environment -> (rank, world_size, local_rank); NCCL over NVLink when launched with
    more than one process (reference src/utils/utils.py:77-93)."""
    if "RANK" not in os.environ or "WORLD_SIZE" not in os.environ:
        return 0, 1, 0
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    backend = "nccl" if torch.cuda.is_available() else "gloo"
    if not dist.is_initialized():
        dist.init_process_group(backend=backend)
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, world, local_rank


def cleanup_distributed():
    if dist.is_initialized():
        dist.destroy_process_group()


def setup_logging(rank):
    level = logging.INFO if rank == 0 else logging.WARNING
    logging.basicConfig(level=level, format="%(asctime)s - %(levelname)s - %(message)s")


def parse_args(parser: ArgumentParser) -> dict:
    args_dict = vars(parser.parse_args())
    log = logging.getLogger(__name__)
    log.info("Command-line arguments:")
    for key, value in args_dict.items():
        log.info(f"{key}: {value}")
    args_dict["model_size_params"] = get_model_params(args_dict["model_size"].lower())
    return args_dict
