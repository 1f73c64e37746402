"""The batched on-device loader on a real B200.

The reference loader cannot run on the GPU box (no /root/reference there), so the chain of evidence is:
  reference loader == literal per-sample restatement   on the CPU generator: tests/test_host_logic.py against
                                                        tests/golden/loader_stream_*.npz, dumped from the reference
  restatement      == this loader                       on the CUDA generator: HERE, bit for bit -- masks drawn by
                                                        torch.rand / argsort on the device, torch.randperm on the
                                                        device, per-sample cutoff filter, per-sample collate.
Also records the loader's own throughput (samples/s with the chunk files cached by the OS).
"""
import os
import random
import sys
import time

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


def _write_chunks(base, ids, n, seed=0, late_every=0):
    g = torch.Generator().manual_seed(seed)
    os.makedirs(base, exist_ok=True)
    for cid in ids:
        w = torch.randn(n, 365, 31, generator=g)
        coords = torch.stack([torch.rand(n, generator=g) * 120 - 60, torch.rand(n, generator=g) * 360 - 180], 1)
        idx = torch.randint(0, 2, (n,), generator=g).float()
        if late_every:
            idx[::late_every] = 2.0
        index = torch.stack([idx, torch.full((n,), 7.0)], 1)
        torch.save(torch.utils.data.TensorDataset(w, coords, index), os.path.join(base, f"weather_dataset_weekly_{cid}.pt"))


def _reference_style_stream_cuda(paths, kind, p, n_masked, shuffle, batch, cutoff=2002.0):
    """src/pretraining/dataloader/pretraining_dataloader.py:186-301 restated literally with device='cuda:0': per-sample
    year loop, mask from torch.rand / argsort on the device, randperm on the device, per-sample filter, then the
    default collate of `batch` samples (torch.stack) the reference DataLoader applies."""
    samples = []
    for i in range(0, len(paths), 3):
        data = list(torch.load(paths[i + 1], weights_only=False, map_location=DEV))
        n = len(data)
        weather = torch.zeros(n, 365, 31, device=DEV)
        coords = torch.zeros(n, 2, device=DEV)
        years = torch.zeros(n, 365, device=DEV)
        interval = torch.zeros(n, 1, device=DEV)
        for j, (w, c, index) in enumerate(data):
            weather[j] = w
            coords[j] = c
            interval[j, 0] = index[1]
            years[j] = 1984.0 + ((index[0] * 365 + torch.arange(365, dtype=torch.float32, device=DEV)) * index[1]) / 365
        if kind == "weatherbert":
            mask = torch.rand(n, 365, 31, device=DEV) < p
        else:
            mask = (torch.argsort(torch.rand(n, 31, device=DEV), dim=-1) < n_masked).unsqueeze(1).expand(-1, 365, -1)
        if shuffle and n > 1:
            perm = torch.randperm(n, device=DEV)
            weather, coords, years, interval, mask = weather[perm], coords[perm], years[perm], interval[perm], mask[perm]
        for j in range(n):
            if torch.max(years[j]) >= cutoff:
                continue
            samples.append((weather[j], coords[j], years[j], interval[j], mask[j]))
    return [tuple(torch.stack([s[k] for s in samples[b0:b0 + batch]]) for k in range(5))
            for b0 in range(0, len(samples), batch)]


@pytest.mark.parametrize("kind,shuffle", [("weatherbert", True), ("weatherformer", True), ("weatherformer", False)])
def test_cuda_loader_stream_is_bit_identical_to_the_reference_algorithm(tmp_path, monkeypatch, kind, shuffle):
    import src.pretraining.dataloader.pretraining_dataloader as dl

    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(dl, "DRY_RUN", True)
    _write_chunks("data/nasa_power/processed/", [1, 34, 53, 72, 81], n=37, late_every=6)
    random.seed(99)
    torch.manual_seed(1234)  # seeds the CUDA generators too
    loader = dl.streaming_dataloader(16, split="train", shuffle=shuffle, masking_function=kind, masking_prob=0.3,
                                     n_masked_features=7)
    assert loader.dataset.device == DEV
    paths = loader.dataset.file_paths
    got = [tuple(t.clone() for t in b) for b in loader]
    off_ours = torch.cuda.default_generators[0].get_offset()
    torch.manual_seed(1234)
    want = _reference_style_stream_cuda(paths, kind, 0.3, 7, shuffle, 16)
    assert torch.cuda.default_generators[0].get_offset() == off_ours  # the generator ends in the same state
    assert [b[0].shape[0] for b in got] == [b[0].shape[0] for b in want]
    assert sum(b[0].shape[0] for b in got) == 5 * (37 - 7)
    for bg, bw in zip(got, want):
        for a, b_ in zip(bg, bw):
            assert a.is_cuda and a.dtype == b_.dtype and torch.equal(a, b_)


def test_cuda_loader_throughput(tmp_path, monkeypatch):
    """4,096-sample chunks, batches of 512 (BASELINE configs[3] per-GPU batch): samples/s of the loader alone."""
    import src.pretraining.dataloader.pretraining_dataloader as dl

    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(dl, "DRY_RUN", True)
    _write_chunks("data/nasa_power/processed/", [1, 34, 53, 72, 81], n=4096)
    random.seed(1)
    torch.manual_seed(0)
    for rep in range(2):  # second pass: chunk files in the page cache
        loader = dl.streaming_dataloader(512, split="train", shuffle=True, masking_function="weatherformer",
                                         n_masked_features=10)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 0
        for batch in loader:
            n += batch[0].shape[0]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    assert n == 5 * 4096
    line = f"loader alone: {n} samples in {dt * 1e3:.1f} ms = {n / dt:,.0f} sequences/s (5 chunks of 4096, batches of 512, weatherformer masks)"
    print(line)
    keep = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(keep):
        with open(os.path.join(keep, "r02_loader_throughput.txt"), "a") as fh:
            fh.write(line + "\n")
    assert n / dt > 20000  # far above the 8-9k sequences/s the training step consumes on one GPU
