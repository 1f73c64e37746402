"""LR range test used by --use-optimal-lr (reference src/base_trainer/find_optimal_lr.py): exponentially
sweep the LR over up to 100 training steps through the trainer's own compute_train_loss, pick the LR at
the steepest smoothed loss descent, restore model and optimiser state. Ranks stop together (MAX flag)."""
import copy
import math

import torch
import torch.distributed as dist

from ..utils.constants import DRY_RUN


def find_optimal_lr(trainer, loader, start_lr=1e-6, end_lr=1.0, num_iter=100, beta=0.98):
    if DRY_RUN:
        num_iter = 5
    net = trainer._get_underlying_model()
    saved_model = copy.deepcopy(net.state_dict())
    saved_opt = copy.deepcopy(trainer.optimizer.state_dict())
    gamma = (end_lr / start_lr) ** (1.0 / max(1, num_iter - 1))
    lr, avg, best, lrs, losses = start_lr, 0.0, float("inf"), [], []
    trainer.model.train()
    it = 0
    for batch in loader:
        if it >= num_iter:
            break
        it += 1
        for group in trainer.optimizer.param_groups:
            group["lr"] = lr
        trainer.optimizer.zero_grad()
        loss = trainer.compute_train_loss(*[t.to(trainer.device) for t in batch])["total_loss"]
        value = loss.item()
        avg = beta * avg + (1 - beta) * value
        smooth = avg / (1 - beta ** it)
        stop = 1.0 if (not math.isfinite(value) or (it > 1 and smooth > 4 * best)) else 0.0
        if trainer.is_distributed:
            flag = torch.tensor(stop, device=trainer.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            stop = flag.item()
        if stop:
            break
        best = min(best, smooth)
        lrs.append(lr)
        losses.append(smooth)
        loss.backward()
        if hasattr(trainer.model, "finish_gradient_sync"):
            trainer.model.finish_gradient_sync()
        trainer.optimizer.step()
        lr *= gamma
    net.load_state_dict(saved_model)
    trainer.optimizer.load_state_dict(saved_opt)
    if trainer.is_distributed:
        dist.barrier()
    if len(losses) < 3:
        return start_lr
    slopes = [(losses[i + 1] - losses[i - 1]) / (math.log(lrs[i + 1]) - math.log(lrs[i - 1])) for i in range(1, len(losses) - 1)]
    return lrs[1 + min(range(len(slopes)), key=slopes.__getitem__)]
