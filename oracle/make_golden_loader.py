"""Generate tests/golden/loader_stream_*.npz: the SAMPLE STREAM of the unmodified reference loader.

Run in the build container only (the GPU box has no /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_loader.py
Synthetic chunk files in the reference's on-disk format (written by `write_chunks` below, which
tests/test_host_logic.py repeats verbatim) are read through the reference's own
`src.pretraining.dataloader.pretraining_dataloader.streaming_dataloader` (DRY_RUN chunk ids, CPU generator,
torch.manual_seed(1234), random.seed(99)) and every batch the DataLoader yields is dumped: sample identity
(weather[:, 0, 0] and a per-sample sum), coords, years, interval, the packed mask bits and the batch sizes.
Every third sample of a chunk starts late enough (index 2) to be dropped by the cutoff-year filter.
"""
import os
import random
import sys
import tempfile

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

IDS = [1, 34, 53, 72, 81]  # DRY_RUN_TRAIN_CHUNK_IDS (src/utils/constants.py:54)
N_PER_CHUNK, LATE_EVERY, BATCH = 21, 5, 8


def write_chunks(base, ids, n, seed=0, late_every=0):
    g = torch.Generator().manual_seed(seed)
    os.makedirs(base, exist_ok=True)
    for cid in ids:
        w = torch.randn(n, 365, 31, generator=g)
        coords = torch.stack([torch.rand(n, generator=g) * 120 - 60, torch.rand(n, generator=g) * 360 - 180], 1)
        idx = torch.randint(0, 2, (n,), generator=g).float()
        if late_every:
            idx[::late_every] = 2.0  # 1984 + ((2*365+364)*7)/365 > 2002: dropped by the cutoff filter
        index = torch.stack([idx, torch.full((n,), 7.0)], 1)
        torch.save(torch.utils.data.TensorDataset(w, coords, index), os.path.join(base, f"weather_dataset_weekly_{cid}.pt"))


def main():
    tmp = tempfile.mkdtemp(prefix="wm_loader_golden_")
    os.chdir(tmp)
    os.environ["DRY_RUN"] = "1"
    write_chunks("data/nasa_power/processed/", IDS, N_PER_CHUNK, late_every=LATE_EVERY)
    import src.pretraining.dataloader.pretraining_dataloader as ref_dl  # the UNMODIFIED reference

    assert ref_dl.__file__.startswith(REF)
    for kind, shuffle in (("weatherbert", True), ("weatherformer", True), ("weatherformer", False)):
        random.seed(99)
        torch.manual_seed(1234)
        loader = ref_dl.streaming_dataloader(BATCH, split="train", shuffle=shuffle, masking_function=kind,
                                             masking_prob=0.3, n_masked_features=7)
        assert loader.dataset.device == "cpu"
        batches = [tuple(t.clone() for t in b) for b in loader]
        cat = [torch.cat([b[i] for b in batches]) for i in range(5)]
        n = cat[0].shape[0]
        assert n == len(IDS) * (N_PER_CHUNK - (N_PER_CHUNK + LATE_EVERY - 1) // LATE_EVERY)
        out = {
            "batch_sizes": np.array([b[0].shape[0] for b in batches], dtype=np.int64),
            "key": cat[0][:, 0, 0].numpy(),
            "weather_sum": cat[0].double().sum((1, 2)).numpy(),
            "coords": cat[1].numpy(), "year": cat[2].numpy(), "interval": cat[3].numpy(),
            "mask_bits": np.packbits(cat[4].numpy().reshape(n, -1), axis=1),
            "chunk_order": np.array([int(p.split("_")[-1].split(".")[0]) for p in loader.dataset.file_paths[1::3]]),
        }
        name = f"loader_stream_{kind}_{'shuffle' if shuffle else 'ordered'}.npz"
        np.savez_compressed(os.path.join(OUT, name), **out)
        print(name, "samples", n, "batches", len(batches), "mask mean", float(cat[4].float().mean()))


if __name__ == "__main__":
    main()
