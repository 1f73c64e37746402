"""Worker of tests/test_gpu_multirank.py (one process per GPU under torchrun, NCCL). Not collected by pytest.

Checks, on real hardware, what the reference gets from DistributedDataParallel (src/base_trainer/base_trainer.py:311-315):
  (1) after backward through BucketedDataParallel every rank's flat gradient buffer equals the AVERAGE of the single-GPU
      gradients of the ranks' batches (recomputed locally, without the wrapper, on the same weights);
  (2) after FusedAdam all ranks hold bit-identical parameters;
  (3) the same with dropout live (per-rank dropout streams differ, parameters still agree).
Rank 0 writes a JSON report to argv[1].
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    out_path = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    import copy

    import torch.nn as nn
    import wm_oracle as O

    from weathermodel_b200 import engine, ops
    from weathermodel_b200.data_parallel import BucketedDataParallel
    from weathermodel_b200.optim import FusedAdam
    from weathermodel_b200.pretraining.models.weatherformer import WeatherFormer

    report = {"world": world}
    hp = O.get_model_params("small")
    B = 4  # per rank
    torch.manual_seed(100 + rank)  # different initial weights per rank: the wrapper must broadcast rank 0's
    model = WeatherFormer(31, 31, dev, **hp).to(dev).train()
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0
    # several layer buckets at this size; WM_DP_OVERLAP (set by the test) picks overlapped buckets or the single all-reduce
    ddp = BucketedDataParallel(model, bucket_cap_mb=2.0)
    report["overlap"] = ddp.overlap
    rt = model.runtime
    report["buckets"] = len(rt.bucket_schedule())
    # every rank can rebuild every rank's batch
    batches = []
    for r in range(world):
        w, c, y, iv = (torch.from_numpy(a).to(dev) for a in O.synthetic_batch(B, 365, seed=50 + r))
        g = torch.Generator().manual_seed(70 + r)
        mask = (torch.rand(B, 31, generator=g) < 0.3).to(dev).unsqueeze(1).expand(-1, 365, -1)
        batches.append((w, c, y, iv, mask))

    def grads_of(net, batch):
        w, c, y, iv, mask = batch
        net.zero_grad()
        loss = engine.former_elbo(net.forward_raw(w, c, y, iv, mask), w, mask, 0.5)["total_loss"]
        loss.backward()
        return loss.item()

    # (1) reference: local model with the broadcast weights, no wrapper, each rank's batch in turn
    local_model = copy.deepcopy(model)
    assert local_model.runtime is not rt and local_model.runtime.grad_ready_hook is None
    ref = None
    ref_losses = []
    for r in range(world):
        ref_losses.append(grads_of(local_model, batches[r]))
        g = local_model.runtime.flat_grads.double()
        ref = g.clone() if ref is None else ref + g
    ref = (ref / world).float()
    # through the wrapper
    w, c, y, iv, mask = batches[rank]
    opt = FusedAdam(ddp.parameters(), lr=1e-3, runtime=rt)
    opt.zero_grad()
    loss = engine.former_elbo(ddp.forward_raw(w, c, y, iv, mask), w, mask, 0.5)["total_loss"]
    loss.backward()
    ddp.finish_gradient_sync()
    torch.cuda.synchronize()
    got = rt.flat_grads
    err = ((got.double() - ref.double()).norm() / ref.double().norm()).item()
    max_abs = (got - ref).abs().max().item()
    report["loss_matches_local"] = abs(loss.item() - ref_losses[rank]) <= 1e-6 * abs(ref_losses[rank])
    report["grad_rel_err_vs_average_of_single_gpu"] = err
    report["grad_max_abs_err"] = max_abs
    report["grad_scale"] = ref.abs().max().item()
    # (2) identical parameters after the optimiser step
    opt.step()
    torch.cuda.synchronize()

    def params_identical():
        flat = rt.flat_params
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        return all(torch.equal(g, gathered[0]) for g in gathered)

    report["params_identical_after_adam"] = params_identical()
    # (3) dropout live: two more steps
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.1
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.1
    seed63 = rt.seed & 0x7FFFFFFFFFFFFFFF  # (the seed is a uint64: int64 tensors take 63 bits of it)
    seeds = [torch.tensor([seed63], dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(seeds, torch.tensor([seed63], dtype=torch.int64, device=dev))
    report["dropout_seeds_differ_per_rank"] = len({int(s.item()) for s in seeds}) == world
    for _ in range(2):
        opt.zero_grad()
        loss = engine.former_elbo(ddp.forward_raw(w, c, y, iv, mask), w, mask, 0.5)["total_loss"]
        loss.backward()
        ddp.finish_gradient_sync()
        opt.step()
    torch.cuda.synchronize()
    report["params_identical_with_dropout"] = params_identical()
    report["device_error"] = ops.device_error()
    errs = [None] * world
    dist.all_gather_object(errs, report)
    if rank == 0:
        with open(out_path, "w") as fh:
            json.dump({"ranks": errs}, fh, indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
