"""Data-parallel host logic with world_size 2 on gloo (CPU): parameter broadcast at wrap time, per-bucket
all-reduce in the order backward finalises the flat gradient buffer, DDP-style averaging, and that the buckets
tile the whole gradient buffer exactly once. The kernels themselves are not involved (no GPU here): each rank
writes a rank-dependent pattern into its flat gradient buffer where the C backward would."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from src.pretraining.models.weatherformer import WeatherFormer
        from weathermodel_b200.data_parallel import BucketedDataParallel

        torch.manual_seed(100 + rank)  # different initial weights per rank: the wrapper must broadcast rank 0's
        model = WeatherFormer(31, 31, torch.device("cpu"), num_heads=4, num_layers=3, hidden_dim_factor=12)
        ddp = BucketedDataParallel(model, bucket_cap_mb=0.2, overlap=True)
        rt = model.runtime
        assert ddp.module is model and rt.grad_ready_hook is not None
        # (1) broadcast
        flat0 = rt.flat_params.clone()
        gathered = [torch.empty_like(flat0) for _ in range(world)]
        dist.all_gather(gathered, flat0)
        assert all(torch.equal(g, gathered[0]) for g in gathered), "parameters not broadcast from rank 0"
        for (name, p), off in zip(rt._named, rt.offsets):  # parameters are still views of the flat buffer
            assert p.data_ptr() == rt.flat_params.data_ptr() + 4 * off, name
        # (2) bucket schedule tiles [0, total) exactly once, head first, embedding last
        sched = rt.bucket_schedule()
        assert sched[0][0] == "head" and sched[-1][0] == "embed" and rt.layers_per_bucket >= 1
        covered = torch.zeros(rt.flat_grads.numel(), dtype=torch.int32)
        for _, _, _, lo, hi in sched:
            covered[lo:hi] += 1
        assert bool((covered == 1).all())
        layer_buckets = [(hi, lo) for kind, hi, lo, _, _ in sched if kind == "layers"]
        assert layer_buckets[0][0] == 3 and layer_buckets[-1][1] == 0
        # (3) simulate backward: rank r writes (r + 1) * pattern, hooks fire per bucket, then join
        pattern = torch.arange(rt.flat_grads.numel(), dtype=torch.float32) % 97 + 1
        rt.flat_grads.copy_(pattern * (rank + 1))
        for _, _, _, lo, hi in sched:
            rt.grad_ready_hook(lo, hi)
        ddp.finish_gradient_sync()
        expect = pattern * (sum(r + 1 for r in range(world)) / world)  # DDP semantics: average over ranks
        assert torch.allclose(rt.flat_grads, expect, rtol=1e-6, atol=0)
        assert not ddp._pending
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_bucketed_data_parallel_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_loader_partition_is_disjoint_across_ranks(monkeypatch):
    sys.path.insert(0, ROOT)
    import src.pretraining.dataloader.pretraining_dataloader as dl

    monkeypatch.setattr(dl, "DRY_RUN", False)
    for world in (2, 4, 8):
        seen = []
        for r in range(world):
            ld = dl.streaming_dataloader(4, split="train", shuffle=False, masking_function="weatherbert", world_size=world, rank=r)
            ids = [int(p.split("_")[-1][:-3]) for p in ld.dataset.file_paths[1::3]]
            seen.append(ids)
        assert len({len(x) for x in seen}) == 1  # same number of chunks on every rank
        flat = sum(seen, [])
        assert len(flat) == len(set(flat))
