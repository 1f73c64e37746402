"""Pretraining CLI with the reference's flags and defaults (src/pretraining/pretraining_main.py:22-131):

    python -m src.pretraining.pretraining_main --model weatherformer --model-size small --batch-size 128
    torchrun --nnodes=1 --nproc-per-node=8 -m src.pretraining.pretraining_main --model weatherformer ...

`--model` accepts every model that is built on the transformer encoder of the B200 path: weatherformer,
weatherbert and their siblings weatherautoencoder, simmtm, weatherformersinusoid, weatherformermixture (they
inherit the encoder; their priors / masks are small torch-op additions). The MLP / CNN baselines are rejected
with a clear message instead of silently running something else."""
import argparse
import logging

from ..utils import utils

_FLAGS = [
    ("--model", dict(default="weatherformer", type=str, help="weatherformer or weatherbert")),
    ("--resume-from-checkpoint", dict(default=None, type=str, help="path to resume from checkpoint")),
    ("--pretrained-model-path", dict(default=None, type=str, help="pretrained model to load before training")),
    ("--batch-size", dict(default=256, type=int, help="GLOBAL batch size (divided by the number of ranks)")),
    ("--n-masked-features", dict(default=10, type=int, help="number of masked features (weatherformer)")),
    ("--n-epochs", dict(default=100, type=int, help="number of training epochs")),
    ("--init-lr", dict(default=0.0005, type=float, help="initial learning rate")),
    ("--use-optimal-lr", dict(action="store_true", help="enable learning rate finding")),
    ("--n-warmup-epochs", dict(default=10, type=float, help="number of warm-up epochs")),
    ("--decay-factor", dict(default=0.99, type=float, help="exponential LR decay factor after warm-up")),
    ("--model-size", dict(default="small", type=str, help="mini (60k), small (2M), medium (8M), large (32M)")),
    ("--masking-prob", dict(default=0.30, type=float, help="fraction of elements to mask (weatherbert)")),
    ("--n-mixture-components", dict(default=1, type=int, help="mixture / sinusoid components (weatherformermixture, weatherformersinusoid)")),
    ("--beta", dict(default=0.5, type=float, help="weight of the KL term (weatherformer)")),
]

parser = argparse.ArgumentParser()
for _flag, _kw in _FLAGS:
    parser.add_argument(_flag, **_kw)


def main():
    rank, world_size, local_rank = utils.setup_distributed()
    utils.setup_logging(rank)
    try:
        args_dict = utils.parse_args(parser)
        args_dict.update(rank=rank, world_size=world_size, local_rank=local_rank)
        kind = args_dict["model"].lower()
        if kind == "weatherformer":
            from .trainers.weatherformer_trainer import weatherformer_training_loop as loop
        elif kind == "weatherbert":
            from .trainers.weatherbert_trainer import weatherbert_training_loop as loop
        elif kind == "weatherautoencoder":
            from .trainers.weatherautoencoder_trainer import weatherautoencoder_training_loop as loop
        elif kind == "simmtm":
            from .trainers.simmtm_trainer import simmtm_training_loop as loop
        elif kind == "weatherformersinusoid":
            from .trainers.weatherformer_sinusoid_trainer import weatherformer_sinusoid_training_loop as loop
        elif kind == "weatherformermixture":
            from .trainers.weatherformer_mixture_trainer import weatherformer_mixture_training_loop as loop
        else:
            raise ValueError(f"model '{kind}' is outside the B200 hot path (supported: weatherformer, weatherbert, "
                             "weatherautoencoder, simmtm, weatherformersinusoid, weatherformermixture)")
        best = loop(args_dict)
        if rank == 0:
            logging.getLogger(__name__).info(f"best validation loss: {best}")
    finally:
        utils.cleanup_distributed()


if __name__ == "__main__":
    main()
