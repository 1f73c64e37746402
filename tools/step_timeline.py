"""GPU timeline of one training step (CUPTI through torch.profiler): kernel time, idle gaps, the largest gaps.
    python tools/step_timeline.py [--workload large]
The bench's own numbers are never taken under a profiler; this is a diagnostic for launch gaps only."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from weathermodel_b200 import engine, ops  # noqa: E402
from weathermodel_b200.optim import FusedAdam  # noqa: E402
from weathermodel_b200.pretraining.models.weatherformer import WeatherFormer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--size", default="large", choices=["mini", "small", "medium", "large"])
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, S, F = a.batch or {"mini": 64, "small": 128, "medium": 256, "large": 512}[a.size], 365, 31
    torch.manual_seed(1234)
    model = WeatherFormer(weather_dim=F, output_dim=F, device=dev, **bench.size_params(a.size)).to(dev).train()
    opt = FusedAdam(model.parameters(), lr=5e-4, runtime=model.runtime)
    w = torch.randn(B, S, F, device=dev)
    c = torch.rand(B, 2, device=dev)
    y = torch.rand(B, S, device=dev) + 1990
    iv = torch.full((B, 1), 7.0, device=dev)

    def step():
        mask = ops.mask_former(S, F, B, 10, device=dev)
        opt.zero_grad()
        y_pad = model.forward_raw(w, c, y, iv, mask)
        loss = engine.former_elbo(y_pad, w, mask, 0.5)["total_loss"]
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    evs = sorted(evs, key=lambda e: e.time_range.start)
    # second step only: from the second mask kernel on
    starts = [i for i, e in enumerate(evs) if "mask_former" in e.name]
    evs = evs[starts[-1]:]
    t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
    busy = sum(e.time_range.end - e.time_range.start for e in evs)
    gaps = []
    end = evs[0].time_range.end
    for prev, e in zip(evs, evs[1:]):
        g = e.time_range.start - end
        if g > 0:
            gaps.append((g, prev.name[:60], e.name[:60]))
        end = max(end, e.time_range.end)
    print(f"step span {(t1 - t0) / 1e3:.2f} ms, kernel time {busy / 1e3:.2f} ms, {len(evs)} kernels, idle {sum(g for g, _, _ in gaps) / 1e3:.2f} ms")
    print(f"gaps > 20 us: {sum(1 for g, _, _ in gaps if g > 20)}; median gap {sorted(g for g, _, _ in gaps)[len(gaps) // 2]:.1f} us")
    for g, a_, b_ in sorted(gaps, reverse=True)[:15]:
        print(f"  {g:8.1f} us  after {a_}  before {b_}")
    # per-kernel totals INSIDE the step (power-capped clocks, warm caches: what the step actually pays, unlike the
    # serialised boost-clock figures of an ncu launch list)
    agg = {}
    for e in evs:
        name = e.name.split("(")[0][:70]
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + (e.time_range.end - e.time_range.start))
    print(f"{'kernel':72s} {'n':>5s} {'total ms':>9s} {'avg us':>8s} {'share':>6s}")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"{name:72s} {n:5d} {t / 1e3:9.3f} {t / n:8.1f} {100 * t / busy:5.1f}%")


if __name__ == "__main__":
    main()
