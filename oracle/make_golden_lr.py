"""Golden vectors for the LR range finder: runs the UNMODIFIED reference `find_optimal_lr`
(/root/reference/src/base_trainer/find_optimal_lr.py) on a stub trainer that replays prescribed loss curves and
records the learning rate it returns. Test infrastructure only (tests/test_host_logic.py compares
weathermodel_b200.base_trainer.find_optimal_lr with these).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_lr.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def loss_curves():
    """name -> per-iteration loss values (100 entries each; the finder stops early on divergence)."""
    rs = np.random.RandomState(0)
    it = np.arange(100, dtype=np.float64)
    curves = {
        # slow decline, steep decline around iteration 55, divergence after 80
        "classic": 2.0 - 0.002 * it - 1.2 / (1 + np.exp(-(it - 55) / 4.0)) + np.where(it > 80, (it - 80) ** 2 * 0.5, 0.0),
        # noisy plateau, no divergence
        "plateau": 1.5 + 0.01 * rs.randn(100),
        # diverges almost immediately
        "early_divergence": np.concatenate([[1.0, 0.98, 0.97], 1.0 * 3.0 ** np.arange(1, 98)]),
        # monotone decline to the end
        "monotone": 3.0 * np.exp(-it / 30.0) + 0.2,
        # steepest decline in the very first step
        "front": np.concatenate([[5.0, 1.0], 1.0 - 0.001 * np.arange(98)]),
    }
    return {k: np.asarray(v, dtype=np.float64) for k, v in curves.items()}


class StubTrainer:
    """Just enough of BaseTrainer for find_optimal_lr: an optimizer with param_groups, a model with train(),
    compute_train_loss replaying `curve`."""

    def __init__(self, curve):
        self.curve = list(curve)
        self.calls = 0
        self.device = torch.device("cpu")
        self.is_distributed = False
        self.param = torch.nn.Parameter(torch.zeros(1))
        self.optimizer = torch.optim.SGD([self.param], lr=123.0)
        self.model = torch.nn.Linear(1, 1)

    def compute_train_loss(self, *batch):
        v = self.curve[min(self.calls, len(self.curve) - 1)]
        self.calls += 1
        return {"total_loss": self.param.sum() * 0.0 + float(v)}


def main():
    sys.path.insert(0, "/root/reference")
    from src.base_trainer.find_optimal_lr import find_optimal_lr  # the reference's own function

    out = {}
    loader = [(torch.zeros(2, 1),)] * 7  # shorter than num_iter: the reference restarts the iterator
    for name, curve in loss_curves().items():
        for start_lr in (1e-5, 5e-4):
            tr = StubTrainer(curve)
            lr = find_optimal_lr(tr, loader, start_lr=start_lr)
            key = f"{name}/{start_lr:g}"
            out[key + "/curve"] = curve
            out[key + "/lr"] = np.float64(lr)
            out[key + "/calls"] = np.int64(tr.calls)
            out[key + "/restored_lr"] = np.float64(tr.optimizer.param_groups[0]["lr"])
            print(key, lr, tr.calls)
    np.savez(os.path.join(ROOT, "tests", "golden", "lr_finder.npz"), **out)


if __name__ == "__main__":
    main()
