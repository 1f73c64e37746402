"""FusedAdam: torch.optim.Adam semantics (the optimiser the reference builds at
src/base_trainer/base_trainer.py:337) as ONE kernel launch over the flat parameter bucket of an
EncoderRuntime, plus a per-tensor launch for any parameter that lives outside it (e.g. yield heads).

state_dict()/load_state_dict() keep torch.optim.Adam's format: per-parameter `step`, `exp_avg`,
`exp_avg_sq` (views into flat state buffers here), so reference checkpoints resume and vice versa.
"""
from typing import Optional

import torch

from . import ops
from .engine import EncoderRuntime


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 runtime: Optional[EncoderRuntime] = None, allow_host_params: bool = False):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam supports a single param group (the reference uses one)")
        # the runtime is looked up through its module at every step: load_pretrained() / deepcopy drop a module's
        # runtime and a new one (new flat buffers) is built lazily -- a stored reference would go stale silently
        self._runtime_owner = runtime.module_ref if runtime is not None else None
        self._step_t = torch.tensor(0.0)  # the `step` entry shared by all flat parameters' state (torch format)
        # host (CPU) parameters are refused unless a test of the state-dict / bookkeeping logic asks for them:
        # there is no CPU path in the product
        self.allow_host_params = allow_host_params
        self._flat_m = None
        self._flat_v = None
        self._step = 0
        # set by graph_step.CapturedTrainStep while it records a step: a device tensor {lr, 1 - b1^t, sqrt(1 - b2^t)} the
        # update kernels read instead of taking the numbers as launch parameters; step counters are then kept by the caller
        self._captured_hyper = None

    @property
    def runtime(self) -> Optional[EncoderRuntime]:
        return self._runtime_owner.runtime if self._runtime_owner is not None else None

    # ---- flat state bound to the runtime's flat parameter buffer ------------------------------------
    def _bind_flat_state(self):
        rt = self.runtime
        if rt is None or rt.flat_params is None:
            return False
        n = rt.flat_params.numel()
        if (self._flat_m is None or self._flat_m.numel() != n or self._flat_m.device != rt.flat_params.device
                or getattr(self, "_bound_to", None) is not rt):
            self._bound_to = rt
            self._flat_m = torch.zeros(n, dtype=torch.float32, device=rt.flat_params.device)
            self._flat_v = torch.zeros(n, dtype=torch.float32, device=rt.flat_params.device)
            for (_, p), off in zip(rt._named, rt.offsets):
                st = self.state[p]
                m = self._flat_m[off:off + p.numel()].view(p.shape)
                v = self._flat_v[off:off + p.numel()].view(p.shape)
                if "exp_avg" in st:  # state loaded from a checkpoint: adopt its values
                    m.copy_(st["exp_avg"])
                    v.copy_(st["exp_avg_sq"])
                    self._step = max(self._step, int(st["step"]))
                st["exp_avg"], st["exp_avg_sq"] = m, v
                st["step"] = self._step_t
            self._step_t.fill_(float(self._step))
        return True

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._flat_m = None  # re-adopt loaded tensors into flat buffers at the next step
        self._flat_v = None
        # the step count comes from the loaded state (or restarts at 0 for an empty one, as torch.optim.Adam does):
        # keeping the old count would bias-correct fresh zero moments as if they were `_step` steps old
        self._step = 0

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        group = self.param_groups[0]
        lr, (b1, b2), eps, wd = group["lr"], group["betas"], group["eps"], group["weight_decay"]
        rt = self.runtime
        flat_ids = set()
        if rt is not None and rt.flat_params is not None and rt.flat_params.is_cuda:
            rt.ensure_flat(rt.flat_params.device)
            in_flat = [p for _, p in rt._named]
            # the flat launch is valid only if every flat parameter received its gradient as a view of
            # the flat gradient buffer in this step (frozen / unused tensors fall back to per-tensor)
            if self._captured_hyper is not None and (self._flat_m is None or getattr(self, "_bound_to", None) is not rt):
                raise RuntimeError("FusedAdam: record a step only after an eager step has bound the optimiser state")
            if all(p.grad is not None and p.grad.data_ptr() == rt.grad_view(i).data_ptr()
                   for i, p in enumerate(in_flat)) and self._bind_flat_state():
                shadow = rt.shadow_ptr()  # the kernel also writes the bf16 copy the GEMMs read (no separate cast pass)
                with torch.cuda.device(rt.flat_params.device):
                    if self._captured_hyper is not None:
                        ops.adam_fused_dev(rt.flat_params, rt.flat_grads, self._flat_m, self._flat_v, self._captured_hyper,
                                           b1, b2, eps, wd, shadow=shadow)
                    else:
                        self._step += 1
                        ops.adam_fused(rt.flat_params, rt.flat_grads, self._flat_m, self._flat_v, self._step, lr, b1, b2,
                                       eps, wd, shadow=shadow)
                        self._step_t.fill_(float(self._step))  # one host tensor shared by every flat parameter's state
                rt.mark_weights_dirty(shadow_written=shadow is not None)
                flat_ids = {id(p) for p in in_flat}
        for p in group["params"]:
            if id(p) in flat_ids or p.grad is None:
                continue
            st = self.state[p]
            if "exp_avg" not in st:
                if self._captured_hyper is not None:
                    raise RuntimeError("FusedAdam: record a step only after an eager step has created the optimiser state")
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if self._captured_hyper is not None:  # recorded step: same device-side scalars as the flat bucket
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
                    raise RuntimeError("FusedAdam: a captured step needs contiguous fp32 CUDA parameters")
                with torch.cuda.device(p.device):
                    ops.adam_fused_dev(p, p.grad, st["exp_avg"], st["exp_avg_sq"], self._captured_hyper, b1, b2, eps, wd)
                if rt is not None:
                    rt.mark_weights_dirty()
                continue
            st["step"] += 1
            k = int(st["step"])
            if p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous():
                with torch.cuda.device(p.device):
                    ops.adam_fused(p, p.grad, st["exp_avg"], st["exp_avg_sq"], k, lr, b1, b2, eps, wd)
            else:
                if not self.allow_host_params:
                    raise RuntimeError("FusedAdam: parameter is not a contiguous fp32 CUDA tensor (no CPU fallback; "
                                       "pass allow_host_params=True only to test host-side bookkeeping)")
                g = p.grad if wd == 0 else p.grad.add(p, alpha=wd)
                st["exp_avg"].lerp_(g, 1 - b1)
                st["exp_avg_sq"].mul_(b2).addcmul_(g, g, value=1 - b2)
                denom = (st["exp_avg_sq"].sqrt() / (1 - b2 ** k) ** 0.5).add_(eps)
                p.addcdiv_(st["exp_avg"], denom, value=-lr / (1 - b1 ** k))
            if rt is not None:
                rt.mark_weights_dirty()
        return loss
