"""LR range test used by --use-optimal-lr (reference src/base_trainer/find_optimal_lr.py:18-168).

Same protocol and the same selection rule as the reference, so `--use-optimal-lr` picks the same learning rate for
the same loss curve (tests/test_host_logic.py replays curves recorded from the reference's own function):
  * the LR grows geometrically from start_lr to end_lr over num_iter real training steps (weights and optimiser
    state are NOT restored afterwards -- the reference trains on from the swept weights, :69-84), the iterator of
    the loader is restarted when it runs out (:66-70);
  * a step whose loss exceeds 5x the best loss so far ends the sweep on every rank (MAX-reduced flag, :93-115);
  * the choice: d loss / d iteration by central differences (np.gradient) over the steps before the loss first
    exceeds 4x its minimum; LR at the most negative slope, divided by 10, floored at 10 * start_lr; 10 * start_lr
    when there is nothing to search (:125-158);
  * the optimiser's LR is put back to what it was (:160-162).
Deviation: none in the result. The device-side difference is the trainer's own step (fused kernels instead of
torch ops); the loss is read back once per step, as in the reference, because the divergence test needs it.
"""
from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist

from ..utils.constants import DRY_RUN


def select_lr(lrs: Sequence[float], losses: Sequence[float], start_lr: float) -> float:
    """The reference's rule (find_optimal_lr.py:125-158) on a recorded sweep."""
    losses = list(losses)
    floor = start_lr * 10
    if not losses:
        return floor
    lo = min(losses)
    first = losses.index(lo)
    stop = len(losses)
    for i in range(first, len(losses)):
        if losses[i] > 4 * lo:
            stop = i
            break
    if stop == 0:
        return floor
    slopes = np.gradient(losses) if len(losses) > 1 else np.zeros(1)
    steepest = int(np.argmin(slopes[:stop]))
    return max(lrs[steepest] / 10, floor)


def find_optimal_lr(trainer, loader, start_lr: float = 1e-5, end_lr: float = 1.0, num_iter: int = None) -> float:
    if num_iter is None:
        num_iter = 5 if DRY_RUN else 100
    groups = trainer.optimizer.param_groups
    original_lr = groups[0]["lr"]
    growth = (end_lr / start_lr) ** (1.0 / (num_iter - 1))
    distributed = dist.is_available() and dist.is_initialized()
    lrs: List[float] = []
    losses: List[float] = []
    best = None
    lr = start_lr
    for g in groups:
        g["lr"] = lr
    it = iter(loader)
    trainer.model.train()
    for _ in range(num_iter):
        try:
            batch = next(it)
        except StopIteration:  # every rank keeps stepping: nobody leaves the collectives below early
            it = iter(loader)
            batch = next(it)
        trainer.optimizer.zero_grad()
        loss = trainer.compute_train_loss(*[t.to(trainer.device) for t in batch])["total_loss"]
        loss.backward()
        if hasattr(trainer.model, "finish_gradient_sync"):
            trainer.model.finish_gradient_sync()
        trainer.optimizer.step()
        value = loss.item()
        lrs.append(lr)
        losses.append(value)
        if best is None or value < best:
            best = value
        stop = value > 5 * best
        if distributed:
            flag = torch.tensor(1.0 if stop else 0.0, device=trainer.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            stop = flag.item() > 0.5
        if stop:
            break
        lr *= growth
        for g in groups:
            g["lr"] = lr
    chosen = select_lr(lrs, losses, start_lr)
    for g in groups:
        g["lr"] = original_lr
    if distributed:
        dist.barrier()
    return chosen
