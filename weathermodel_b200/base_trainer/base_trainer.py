"""BaseTrainer: the reference's training shell (src/base_trainer/base_trainer.py) on the B200 engine.

Public surface kept: constructor arguments, train(), save_checkpoint()/load_checkpoint(), the abstract
get_dataloaders / compute_train_loss / compute_validation_loss, checkpoint dict keys and file names,
output JSON, `self.model` being the model or a wrapper with `.module`.
What differs underneath: BucketedDataParallel instead of DDP, FusedAdam instead of optim.Adam (same
state_dict format), loss scalars accumulated on the device and read back once per epoch instead of one
`.item()` per key per step, and a sequences/sec figure logged per epoch.
"""
import json
import logging
import os
import time
from abc import ABC, abstractmethod
from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.distributed as dist
from torch.utils.data import DataLoader

from ..base_models.base_model import BaseModel
from ..data_parallel import BucketedDataParallel
from ..optim import FusedAdam
from ..utils import utils
from ..utils.constants import DATA_DIR, DRY_RUN


class BaseTrainer(ABC):
    # trainers whose compute_train_loss has no host round trip (.item(), boolean gathers) opt in to recorded steps
    _graph_capturable = False

    def __init__(self, model: BaseModel, batch_size: int, num_epochs: int, init_lr: float = 1e-4,
                 num_warmup_epochs: int = 5, decay_factor: Optional[float] = None,
                 pretrained_model_path: Optional[str] = None, resume_from_checkpoint: Optional[str] = None,
                 rank: int = 0, world_size: int = 1, local_rank: int = 0):
        self.model: Union[BaseModel, BucketedDataParallel]
        self.logger = logging.getLogger(__name__)
        self.rank, self.world_size, self.local_rank = rank, world_size, local_rank
        self.is_distributed = world_size > 1
        self.batch_size, self.num_epochs, self.current_epoch = batch_size, num_epochs, None
        self.device = torch.device(f"cuda:{local_rank}" if torch.cuda.is_available() else "cpu")
        self.model = model.to(self.device)
        # pretrained weights must be in place before the wrapper broadcasts and the optimiser binds
        self._setup_model_directory()
        self._load_pretrained_model(pretrained_model_path)
        if self.is_distributed:
            self.model = BucketedDataParallel(self.model)
            self.batch_size = batch_size // world_size  # --batch-size is the GLOBAL batch (reference :315)
        self._setup_training_components(init_lr, num_warmup_epochs, decay_factor)
        self._setup_logging_and_output()
        self.start_epoch = 0
        if resume_from_checkpoint and os.path.exists(resume_from_checkpoint):
            self.load_checkpoint(resume_from_checkpoint)
        self.last_epoch_seq_per_s = float("nan")

    # ---------------------------------------------------------------- public API
    def get_model_name(self) -> str:
        m = self._get_underlying_model()
        return f"{m.name}_{m.total_params_formatted()}"

    def get_current_epoch(self) -> Optional[int]:
        return self.current_epoch

    def get_num_epochs(self) -> int:
        return self.num_epochs

    def train(self, use_optimal_lr: bool = False) -> float:
        if use_optimal_lr and self.start_epoch == 0:
            self._find_and_set_optimal_lr()
        for epoch in range(self.start_epoch, self.num_epochs):
            train_loader, val_loader = self.get_dataloaders(shuffle=True)
            self.current_epoch = epoch
            train_loss = self._train_epoch(train_loader)
            val_loss = self._validate_epoch(val_loader)
            if val_loss < self.best_val_loss:
                self.best_val_loss = val_loss
                self._save_best_model()
            if self.rank == 0:
                self.logger.info(
                    f"Epoch [{epoch + 1} / {self.num_epochs}]: Train loss: {train_loss:.3f} Validation loss: "
                    f"{val_loss:.3f} Best Val loss: {self.best_val_loss:.3f} ({self.last_epoch_seq_per_s:.0f} seq/s/rank)")
                if epoch % 5 == 1 or epoch == self.num_epochs - 1:
                    self.save_checkpoint(epoch, val_loss)
                self._save_output_json()
        if self.rank == 0:
            self._cleanup_numbered_checkpoints()
        return self.best_val_loss

    def save_checkpoint(self, epoch: int, val_loss: float):
        if self.rank != 0:
            return
        net = self._get_underlying_model()
        ckpt = {
            "epoch": epoch + 1,
            "model_state_dict": net.state_dict(),
            "optimizer_state_dict": self.optimizer.state_dict(),
            "scheduler_state_dict": self.scheduler.state_dict(),
            "best_val_loss": self.best_val_loss,
            "output_json": self.output_json,
            # extra key (ignored by the reference loader): position of the dropout streams, so that a resumed run does
            # not replay the masks of its first steps
            "wm_dropout_steps": [m.runtime.step_counter for m in net.modules() if hasattr(m, "runtime")],
        }
        stem = self.model_dir + self.get_model_name()
        numbered = [f"{stem}_epoch_{epoch}_checkpoint.pth", f"{stem}_epoch_{epoch}.pth"]
        torch.save(ckpt, numbered[0])
        torch.save(ckpt, f"{stem}_latest_checkpoint.pth")
        torch.save(net, numbered[1])  # whole pickled module, as the reference does
        torch.save(net, f"{stem}_latest.pth")
        self.saved_checkpoint_files.extend(numbered)

    def load_checkpoint(self, checkpoint_path: str):
        ckpt = torch.load(checkpoint_path, map_location=self.device, weights_only=False)
        self._get_underlying_model().load_state_dict(ckpt["model_state_dict"])
        self.optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        self.scheduler.load_state_dict(ckpt["scheduler_state_dict"])
        self.start_epoch = ckpt["epoch"]
        self.best_val_loss = ckpt["best_val_loss"]
        self.output_json = ckpt["output_json"]
        owners = [m for m in self._get_underlying_model().modules() if hasattr(m, "runtime")]
        for m, steps in zip(owners, ckpt.get("wm_dropout_steps", [])):  # absent in reference-written checkpoints
            m.runtime.step_counter = int(steps)
        if self.rank == 0:
            self.logger.info(f"Loaded checkpoint from {checkpoint_path}, resuming from epoch {self.start_epoch}")

    # ---------------------------------------------------------------- to be implemented by subclasses
    @abstractmethod
    def get_dataloaders(self, shuffle: bool = True, cross_validation_k: Optional[int] = None
                        ) -> Tuple[DataLoader, DataLoader]:
        ...

    @abstractmethod
    def compute_train_loss(self, *input_data: torch.Tensor) -> Dict[str, torch.Tensor]:
        ...

    @abstractmethod
    def compute_validation_loss(self, *input_data: torch.Tensor) -> Dict[str, torch.Tensor]:
        ...

    # ---------------------------------------------------------------- epochs
    def _masked_mean(self, tensor: torch.Tensor, mask: torch.Tensor, dim: Tuple[int, ...]):
        return (tensor * mask).sum(dim=dim) / (mask.sum(dim=dim).clamp(min=1))

    def _run_epoch(self, loader, split: str, training: bool) -> Dict[str, float]:
        keys = list(self.output_json["losses"][split])
        sums = torch.zeros(len(keys), dtype=torch.float64, device=self.device)
        steps, seqs = 0, 0
        t0 = time.perf_counter()
        losses = None
        for batch in loader:
            batch = [t.to(self.device) for t in batch]
            if training:
                losses = None  # (drop the previous step's autograd graph before a step may be recorded)
                losses = self._captured_step(batch)  # small shapes: the whole step body as one CUDA-graph replay
                if losses is None:
                    self.optimizer.zero_grad()
                    losses = self.compute_train_loss(*batch)
                    loss = losses["total_loss"]
                    if self.rank == 0 and DRY_RUN:
                        print(f"Train loss: {loss.item()}")
                    loss.backward()
                    if isinstance(self.model, BucketedDataParallel):
                        self.model.finish_gradient_sync()
                    self.optimizer.step()
            else:
                with torch.no_grad():
                    losses = self.compute_validation_loss(*batch)
            # stays on the device: one host read-back per epoch instead of one .item() per key per step
            sums += torch.stack([losses[k].detach().double().reshape(()) for k in keys])
            steps += 1
            seqs += int(batch[0].shape[0])
        if training:
            self.scheduler.step()
        if self.is_distributed:
            dist.barrier()
        totals = sums.cpu().tolist()
        if training and steps:
            self.last_epoch_seq_per_s = seqs / max(1e-9, time.perf_counter() - t0)
        return self._average_losses({k: v for k, v in zip(keys, totals)}, steps)

    # ---------------------------------------------------------------- recorded steps (launch-bound shapes)
    def _graph_eligible(self, batch) -> bool:
        """WM_CUDA_GRAPH=0 never, =1 whenever possible, unset: for batch tokens x model width up to 2^27 -- every BASELINE
        config at its batch size. The small shapes are launch-bound (mini: 1.12 -> 0.64 ms per step); WeatherFormer large at
        512 sequences per GPU is 55 ms of kernels per step and still gains ~0.8 ms end to end: the launches that follow
        each step's loss read-back no longer wait for Python."""
        mode = os.environ.get("WM_CUDA_GRAPH", "auto")
        if mode == "0" or not self._graph_capturable or self.is_distributed or DRY_RUN or self.device.type != "cuda":
            return False
        runtimes = [m.runtime for m in self._get_underlying_model().modules() if hasattr(m, "runtime")]
        if len(runtimes) != 1 or getattr(self.optimizer, "runtime", None) is not runtimes[0]:
            return False
        if not all(p.requires_grad for _, p in runtimes[0].named_flat_params()):
            return False  # frozen encoder phases of the yield trainers stay on the eager path
        if mode == "1":
            return True
        D = runtimes[0]._dims()[0]
        return batch[0].dim() == 3 and batch[0].shape[0] * batch[0].shape[1] * D <= (1 << 27)

    def _captured_step(self, batch):
        """Returns the step's loss dict after running it as a CUDA-graph replay, or None (caller runs the eager step).
        The first two batches of every shape run eagerly (engine handles, GEMM tuning, optimiser state); the third is
        recorded while it trains (graph_step.CapturedTrainStep); a failed recording disables the feature for this trainer."""
        if getattr(self, "_graph_disabled", False) or not self._graph_eligible(batch):
            return None
        from ..graph_step import CapturedTrainStep

        if not hasattr(self, "_graph_steps"):
            self._graph_steps, self._graph_seen = {}, {}
        key = tuple((tuple(t.shape), t.dtype) for t in batch)
        step = self._graph_steps.get(key)
        if step is not None:
            return step(*batch)
        seen = self._graph_seen.get(key, 0)
        self._graph_seen[key] = seen + 1
        if seen < 2 or len(self._graph_steps) >= 4:
            return None
        try:
            step = CapturedTrainStep(self.optimizer, self.compute_train_loss, batch)
        except Exception as e:  # noqa: BLE001 -- any failure: keep training eagerly
            self._graph_disabled = True
            if os.environ.get("WM_GRAPH_DEBUG"):
                import traceback
                traceback.print_exc()
            from ..graph_step import repair_default_generator
            repair_default_generator(self.device)
            self.logger.warning(f"CUDA-graph recording of the training step failed ({type(e).__name__}: {e}); staying eager")
            return None
        self._graph_steps[key] = step
        if self.rank == 0:
            self.logger.info(f"training step recorded as a CUDA graph for batch shape {tuple(batch[0].shape)}")
        return step.first_losses

    def _train_epoch(self, loader) -> float:
        self.model.train()
        if self.rank == 0:
            self.logger.info("Started training epoch.")
        avg = self._run_epoch(loader, "train", training=True)
        self._update_output_json_losses("train", avg)
        return avg["total_loss"]

    def _validate_epoch(self, loader) -> float:
        self.model.eval()
        if self.rank == 0:
            self.logger.info("Started validation epoch.")
        avg = self._run_epoch(loader, "val", training=False)
        self._update_output_json_losses("val", avg)
        return avg["total_loss"]

    # ---------------------------------------------------------------- set-up helpers
    def _setup_training_components(self, init_lr: float, num_warmup_epochs: int, decay_factor: Optional[float]):
        self.init_lr, self.num_warmup_epochs, self.decay_factor = init_lr, num_warmup_epochs, decay_factor
        self.best_val_loss = float("inf")
        runtimes = [m.runtime for m in self._get_underlying_model().modules() if hasattr(m, "runtime")]
        self.optimizer = FusedAdam(self.model.parameters(), lr=init_lr, runtime=runtimes[0] if runtimes else None)
        if self.rank == 0:
            self.logger.info("using cosine annealing" if decay_factor is None
                             else f"using exponential annealing with decay factor {decay_factor}")
        self.scheduler = utils.get_scheduler(self.optimizer, num_warmup_epochs, self.num_epochs, decay_factor)

    def _setup_logging_and_output(self):
        net = self._get_underlying_model()
        if self.rank == 0:
            self.logger.info(f"Total number of parameters: {net.total_params_formatted()}")
            self.logger.info(f"Distributed training: {self.is_distributed}, World size: {self.world_size}")
            self.logger.info(f"Batch size per GPU: {self.batch_size}")
        self.output_json = {
            "model_config": {
                "total_params": net.total_params(),
                "batch_size": self.batch_size * self.world_size,
                "batch_size_per_gpu": self.batch_size,
                "world_size": self.world_size,
                "init_lr": self.init_lr,
                "num_warmup_epochs": self.num_warmup_epochs,
                "decay_factor": self.decay_factor,
                "model_layers": str(net),
            },
            "losses": {"train": {"total_loss": []}, "val": {"total_loss": []}},
        }
        self.saved_checkpoint_files: List[str] = []

    def _setup_model_directory(self):
        if self.rank == 0:
            self.model_dir = DATA_DIR + "trained_models/pretraining/"
            os.makedirs(self.model_dir, exist_ok=True)

    def _load_pretrained_model(self, pretrained_model_path: Optional[str]):
        if pretrained_model_path and os.path.exists(pretrained_model_path):
            pretrained = torch.load(pretrained_model_path, weights_only=False)
            self.logger.info(f"Loading pretrained model from: {pretrained_model_path}")
            self._get_underlying_model().load_pretrained(pretrained)

    def _get_underlying_model(self) -> BaseModel:
        return self.model.module if isinstance(self.model, BucketedDataParallel) else self.model  # type: ignore

    def _find_and_set_optimal_lr(self):
        from .find_optimal_lr import find_optimal_lr

        train_loader, _ = self.get_dataloaders(shuffle=True)
        lr = find_optimal_lr(self, train_loader, start_lr=self.init_lr)
        if self.is_distributed:
            t = torch.tensor(lr, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            lr = t.item() / self.world_size
            dist.barrier()
        if self.rank == 0:
            self.output_json["model_config"]["init_lr"] = lr
            self.logger.info(f"Using averaged optimal learning rate: {lr:.6f}")
        for group in self.optimizer.param_groups:
            group["lr"] = lr

    def _initialize_loss_dict(self, split: str) -> Dict[str, float]:
        return {key: 0.0 for key in self.output_json["losses"][split]}

    def _sync_distributed_training(self):
        if self.is_distributed:
            dist.barrier()

    def _average_losses(self, total_loss_dict: Dict[str, float], loader_len: int) -> Dict[str, float]:
        out = {}
        for key, total in total_loss_dict.items():
            out[key] = total / loader_len
            if self.is_distributed:
                t = torch.tensor(out[key], device=self.device)
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                out[key] = t.item() / self.world_size
        return out

    def _update_output_json_losses(self, split: str, avg_loss_dict: Dict[str, float]):
        if self.rank == 0:
            for key in self.output_json["losses"][split]:
                self.output_json["losses"][split][key].append(avg_loss_dict[key])

    def _save_output_json(self):
        if self.rank != 0:
            return
        with open(self.model_dir + f"{self.get_model_name()}_output.json", "w") as f:
            json.dump(self.output_json, f, indent=2)

    def _cleanup_numbered_checkpoints(self):
        if self.rank != 0:
            return
        for path in self.saved_checkpoint_files:
            try:
                if os.path.exists(path):
                    os.remove(path)
            except OSError as e:
                self.logger.warning(f"Failed to remove {os.path.basename(path)}: {e}")
        self.saved_checkpoint_files.clear()

    def _get_n_masked_features(self, epoch, initial_n_masked_features=1):
        """+2 masked features every 5 epochs, capped at 25 (reference :517-523)."""
        if epoch is None:
            return initial_n_masked_features
        return min(initial_n_masked_features + 2 * (epoch // 5), 25)

    def _save_best_model(self):
        if self.rank != 0:
            return
        path = self.model_dir + f"{self.get_model_name()}_best.pth"
        torch.save(self._get_underlying_model(), path)
        self.logger.info(f"Saved best model with validation loss {self.best_val_loss:.4f} to {path}")
