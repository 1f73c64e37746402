"""BucketedDataParallel: the data-parallel wrapper that replaces torch DDP for this path
(reference: src/base_trainer/base_trainer.py:311-315).

Semantics kept from DDP: parameters broadcast from rank 0 at wrap time, gradients AVERAGED over ranks
(of per-rank mean losses -- not re-normalised globally), `.module` gives the wrapped model.
Mechanism: the encoder's backward publishes gradients straight into one flat fp32 buffer; as soon as the
kernels producing a bucket (one or more encoder layers, top first) are enqueued, that slice is all-reduced
with NCCL on its own stream (async_op) while the next layers' backward kernels run; all handles are waited
once before the optimiser step. No per-parameter hooks, no gradient copies.
"""
import os
from typing import List

import torch
import torch.distributed as dist
import torch.nn as nn


class BucketedDataParallel(nn.Module):
    def __init__(self, module: nn.Module, process_group=None, bucket_cap_mb: float = 25.0, overlap=None):
        """overlap: True = all-reduce every bucket as soon as its gradients are enqueued (DDP's scheme); False = one
        all-reduce over the whole flat gradient buffer when backward is done. Default from WM_DP_OVERLAP (unset: False).
        Why the default is NOT to overlap on this path (profiles/r02_dp8_timeline.txt, 8 x B200): the payload is tiny --
        128 MB per 60 ms step, ~0.5 ms at NVSwitch speed -- but every big kernel here is a persistent one-CTA-per-SM grid
        with static tile striding. An NCCL kernel that holds even a few SMs while such a kernel starts pushes that
        kernel's displaced CTAs into a second wave (gemm_tn2 311 -> 791 us, gemm_wgrad 239 -> 480 us when an all-reduce
        was resident), and NCCL itself, starved of SMs, stays resident for 1-3 ms per 16 MB bucket. Overlapped buckets
        cost +1.8 ms per step (+2.9 %) at 8 GPUs; one exposed all-reduce at the end costs less than a third of that."""
        super().__init__()
        self.module = module
        self.process_group = process_group
        self.overlap = (os.environ.get("WM_DP_OVERLAP", "0") == "1") if overlap is None else bool(overlap)
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._pending: List = []
        self.bucket_cap_mb = bucket_cap_mb
        self._owners = [m for m in module.modules() if hasattr(m, "runtime")]
        if not self._owners:
            raise ValueError("BucketedDataParallel wraps models built on weathermodel_b200's encoder")
        self._backend = dist.get_backend(process_group) if dist.is_initialized() else "none"
        self._attach()
        self._broadcast_parameters()

    @property
    def _runtimes(self):
        """The owners' CURRENT runtimes: load_pretrained() / deepcopy make a module rebuild its runtime lazily."""
        return [m.runtime for m in self._owners]

    def _attach(self):
        """(Re)install the bucket hook on every runtime that lacks it. Cheap; called at wrap time and before every
        forward, so a runtime rebuilt after wrapping cannot drop out of the gradient all-reduce silently."""
        dev = next(self.module.parameters()).device
        for rt in self._runtimes:
            if getattr(rt, "_dp_owner", None) is self and rt.grad_ready_hook is not None:
                continue
            rt.ensure_flat(dev)
            # bucket = as many whole layers as fit the cap (DDP default 25 MiB)
            per_layer = (rt.offsets[14] - rt.offsets[2]) * 4 if len(rt.offsets) > 14 else rt.flat_params.numel() * 4
            rt.layers_per_bucket = max(1, int(self.bucket_cap_mb * 2 ** 20 // max(1, per_layer)))
            rt.grad_ready_hook = self._make_hook(rt)
            rt._dp_owner = self
            if dist.is_initialized():  # every rank draws its own dropout masks (one generator per process in torch DDP, too)
                rt.seed = (rt.base_seed ^ (0x9E3779B97F4A7C15 * (dist.get_rank(self.process_group) + 1))) & 0xFFFFFFFFFFFFFFFF

    def _broadcast_parameters(self):
        if self.world_size == 1:
            return
        for rt in self._runtimes:
            dist.broadcast(rt.flat_params, src=0, group=self.process_group)
            rt.mark_weights_dirty()
        flat_ids = {id(p) for rt in self._runtimes for _, p in rt._named}
        for p in self.module.parameters():
            if id(p) not in flat_ids:
                dist.broadcast(p.data, src=0, group=self.process_group)
        for b in self.module.buffers():
            dist.broadcast(b, src=0, group=self.process_group)

    def _make_hook(self, rt):
        def hook(lo: int, hi: int):
            if self.world_size == 1 or not self.overlap:
                return
            sl = rt.flat_grads[lo:hi]
            if self._backend == "nccl":
                work = dist.all_reduce(sl, op=dist.ReduceOp.AVG, group=self.process_group, async_op=True)
                self._pending.append((work, None))
            else:
                work = dist.all_reduce(sl, op=dist.ReduceOp.SUM, group=self.process_group, async_op=True)
                self._pending.append((work, sl))
        return hook

    def finish_gradient_sync(self):
        """Join every in-flight bucket all-reduce (stream-level wait; no host sync on NCCL) and average
        parameters that live outside the flat buckets."""
        for rt in self._runtimes:
            if self.world_size > 1 and getattr(rt, "_dp_owner", None) is not self:
                raise RuntimeError("BucketedDataParallel: an encoder runtime was rebuilt after wrapping and ran its "
                                   "backward without the bucket hook (call the wrapper's forward, or re-wrap)")
        if self.world_size > 1 and not self.overlap:  # one collective over the whole flat buffer, after backward
            for rt in self._runtimes:
                if self._backend == "nccl":
                    dist.all_reduce(rt.flat_grads, op=dist.ReduceOp.AVG, group=self.process_group)
                else:
                    dist.all_reduce(rt.flat_grads, op=dist.ReduceOp.SUM, group=self.process_group)
                    rt.flat_grads.div_(self.world_size)
        for work, sl in self._pending:
            work.wait()
            if sl is not None:
                sl.div_(self.world_size)
        self._pending.clear()
        if self.world_size > 1:
            flat_ids = {id(p) for rt in self._runtimes for _, p in rt._named}
            for p in self.module.parameters():
                if id(p) not in flat_ids and p.grad is not None:
                    dist.all_reduce(p.grad, group=self.process_group)
                    p.grad.div_(self.world_size)

    def forward(self, *args, **kwargs):
        self._attach()
        return self.module(*args, **kwargs)

    def forward_raw(self, *args, **kwargs):
        """The wrapped encoder's padded raw head output (the trainers' fused-loss path)."""
        self._attach()
        return self.module.forward_raw(*args, **kwargs)
