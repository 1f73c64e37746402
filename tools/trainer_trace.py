"""Where a trainer epoch spends host time: wraps the loader handed to <Model>Trainer._train_epoch and logs, per batch,
how long the trainer waited for the loader (fetch) and how long it kept the batch (step issue). Diagnostic for the gap
between bench.py's `value` and `e2e_trainer`.   python tools/trainer_trace.py [size] [batch] [weatherformer|weatherbert]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


class Traced:
    def __init__(self, loader):
        self.loader, self.fetch, self.hold = loader, [], []

    def __iter__(self):
        it = iter(self.loader)
        while True:
            t0 = time.perf_counter()
            try:
                b = next(it)
            except StopIteration:
                return
            t1 = time.perf_counter()
            yield b
            self.fetch.append(t1 - t0)
            self.hold.append(time.perf_counter() - t1)


def main():
    size = sys.argv[1] if len(sys.argv) > 1 else "large"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else {"mini": 64, "small": 128, "medium": 256, "large": 512}[size]
    # re-use bench.trainer_leg's set-up by monkeypatching the epoch call
    import weathermodel_b200.base_trainer.base_trainer as bt
    real = bt.BaseTrainer._train_epoch
    state = {"n": 0}

    def traced_epoch(self, loader):
        state["n"] += 1
        if state["n"] < 2:
            return real(self, loader)
        tr = Traced(loader)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = real(self, tr)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        print(f"epoch wall {t * 1e3:.1f} ms, {len(tr.fetch)} batches; loader wait total {sum(tr.fetch) * 1e3:.1f} ms, step issue total {sum(tr.hold) * 1e3:.1f} ms")
        print("fetch ms :", " ".join(f"{x * 1e3:.1f}" for x in tr.fetch))
        print("issue ms :", " ".join(f"{x * 1e3:.1f}" for x in tr.hold))
        return out

    bt.BaseTrainer._train_epoch = traced_epoch
    kind = sys.argv[3] if len(sys.argv) > 3 else "weatherformer"
    print(bench.trainer_leg(kind, size, B, torch.device("cuda:0")))


if __name__ == "__main__":
    main()
