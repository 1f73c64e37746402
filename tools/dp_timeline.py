"""GPU timeline of one DATA-PARALLEL training step on rank 0 (CUPTI through torch.profiler), WeatherFormer large,
512 sequences per GPU: where the bucketed NCCL all-reduces sit relative to the backward kernels, how much of them is
exposed, and what the step costs next to the same step without communication.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dp_timeline.py
Diagnostic only: nothing measured under a profiler is reported as a bench number."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from weathermodel_b200 import engine, ops  # noqa: E402
from weathermodel_b200.data_parallel import BucketedDataParallel  # noqa: E402
from weathermodel_b200.optim import FusedAdam  # noqa: E402
from weathermodel_b200.pretraining.models.weatherformer import WeatherFormer  # noqa: E402


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, S, F = 512, 365, 31
    torch.manual_seed(1234)
    net = WeatherFormer(weather_dim=F, output_dim=F, device=dev, **bench.size_params("large")).to(dev).train()
    model = BucketedDataParallel(net) if world > 1 else net
    opt = FusedAdam(model.parameters(), lr=5e-4, runtime=net.runtime)
    g = torch.Generator().manual_seed(rank)
    w = torch.randn(B, S, F, generator=g).to(dev)
    c = torch.rand(B, 2, generator=g).to(dev)
    y = (torch.rand(B, S, generator=g) + 1990).to(dev)
    iv = torch.full((B, 1), 7.0, device=dev)

    def step():
        mask = ops.mask_former(S, F, B, 10, device=dev)
        opt.zero_grad()
        loss = engine.former_elbo(net.forward_raw(w, c, y, iv, mask), w, mask, 0.5)["total_loss"]
        loss.backward()
        if world > 1:
            model.finish_gradient_sync()
        opt.step()

    def timed(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    for _ in range(4):
        step()
    ms_dp = timed(10)
    # the same step with the bucket hook switched off (no collective at all): what communication costs in total
    hook = net.runtime.grad_ready_hook
    net.runtime.grad_ready_hook = None
    if world > 1:
        model.world_size, ws = 1, model.world_size  # (finish_gradient_sync then issues nothing)
    ms_nocomm = timed(10)
    if world > 1:
        model.world_size = ws
    net.runtime.grad_ready_hook = hook
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            step()
        torch.cuda.synchronize()
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None),
                 key=lambda e: e.time_range.start)
    starts = [i for i, e in enumerate(evs) if "mask_former" in e.name]
    evs = evs[starts[-1]:]
    t0 = evs[0].time_range.start
    nccl = [e for e in evs if "nccl" in e.name.lower()]
    comp = [e for e in evs if "nccl" not in e.name.lower()]
    t1 = max(e.time_range.end for e in evs)
    print(f"# WM_DP_OVERLAP={os.environ.get('WM_DP_OVERLAP', '0')} NCCL_ALGO={os.environ.get('NCCL_ALGO', '-')} NCCL_PROTO={os.environ.get('NCCL_PROTO', '-')}")
    print(f"# world {world}: step {ms_dp:.2f} ms with the gradient all-reduce, {ms_nocomm:.2f} ms with the hook off (max over ranks, "
          f"10 steps, CUDA events) -> communication costs {ms_dp - ms_nocomm:+.2f} ms = {100 * (ms_dp / ms_nocomm - 1):+.1f} %")
    print(f"# profiled step on rank 0: span {(t1 - t0) / 1e3:.2f} ms, {len(comp)} compute kernels {sum(e.time_range.end - e.time_range.start for e in comp) / 1e3:.2f} ms, "
          f"{len(nccl)} NCCL kernels {sum(e.time_range.end - e.time_range.start for e in nccl) / 1e3:.2f} ms")
    # exposed NCCL time: parts of NCCL kernel intervals during which no compute kernel runs
    ivs = sorted((e.time_range.start, e.time_range.end) for e in comp)
    def covered(a, b):
        tot = 0
        for s, e in ivs:
            if e <= a:
                continue
            if s >= b:
                break
            tot += min(e, b) - max(s, a)
        return tot
    adam_start = next((e.time_range.start for e in comp if "adam" in e.name), t1)
    last_bwd_end = max((e.time_range.end for e in comp if e.time_range.end <= adam_start and "adam" not in e.name), default=t0)
    print(f"# gap between the last backward kernel and Adam: {(adam_start - last_bwd_end):.1f} us")
    print(f"{'NCCL kernel':50s} {'start ms':>9s} {'dur us':>8s} {'exposed us':>10s}   overlapping compute kernels")
    for e in nccl:
        a, b = e.time_range.start, e.time_range.end
        names = sorted({c_.name.split("(")[0].replace("void ", "").replace("wm::", "")[:22] for c_ in comp if c_.time_range.end > a and c_.time_range.start < b})
        print(f"{e.name[:50]:50s} {(a - t0) / 1e3:9.2f} {b - a:8.1f} {(b - a) - covered(a, b):10.1f}   {', '.join(names)[:110]}")
    # compute kernels that ran while an NCCL kernel was resident, vs their average when alone
    agg = {}
    nv = sorted((e.time_range.start, e.time_range.end) for e in nccl)
    for e in comp:
        a, b = e.time_range.start, e.time_range.end
        during = any(s < b and en > a for s, en in nv)
        name = e.name.split("(")[0].replace("void ", "")[:48]
        d = agg.setdefault(name, [0, 0.0, 0, 0.0])
        d[2 if during else 0] += 1
        d[3 if during else 1] += b - a
    print(f"{'compute kernel':50s} {'alone n':>8s} {'avg us':>8s} {'w/ NCCL n':>10s} {'avg us':>8s}")
    for name, (n0, t0_, n1, t1_) in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][3]))[:12]:
        if n1:
            print(f"{name:50s} {n0:8d} {t0_ / max(n0, 1):8.1f} {n1:10d} {t1_ / n1:8.1f}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
