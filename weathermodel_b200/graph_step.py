"""CapturedTrainStep: the body of BaseTrainer's step -- zero_grad -> compute_train_loss -> backward -> optimizer.step()
(reference src/base_trainer/base_trainer.py:239-252) -- recorded ONCE per batch shape in a CUDA graph and replayed.

Why: the small BASELINE configs (WeatherBERT mini, WeatherFormer small, the yield fine-tune) are launch-bound -- a
step is ~80-150 kernels of a few microseconds each, issued through Python + ctypes. A replay is one launch.

What makes a replay a NEW step although every launch parameter is frozen:
  * inputs        copied into static device buffers before the replay (one small copy kernel per tensor);
  * dropout       three fresh words per replay: the host writes them into a slot of a pinned ring and enqueues a small
                  H2D copy in front of the replay; the recorded wm_step_params_apply installs them and every dropout
                  site folds them into its keys (csrc/wm_common.cuh g_wm_drop_mix). Forward and backward of one replay
                  see the same words. (The copy is NOT part of the graph: a recorded copy would read the pinned buffer
                  when the replay executes, by which time the host may have written the next step's values.)
  * Adam          learning rate and the two bias corrections travel in the same slot and are read from device memory
                  by wm_adam_fused_dev; the host keeps the optimiser's step counters in line;
  * torch RNG     draws inside the step (the yield model's randn_like) are registered with the graph by torch itself.
Losses come back as static device tensors (overwritten by the next replay: consume them in stream order). Single-GPU only
(the bucketed NCCL all-reduce stays on the eager path). The mix words stay installed after a replay; eager kernels that
follow simply fold them in as well (forward and backward still agree) -- ops.step_params_apply(None) restores zeros.
"""
import threading
from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import ops

_MASK64 = (1 << 64) - 1
# Held while a step is being recorded. Helper threads that issue CUDA work of their own (the loader's chunk prefetch:
# pinned allocation + H2D copy on a side stream) take it around that work, so nothing foreign runs during a capture.
CAPTURE_LOCK = threading.RLock()


def _splitmix(z: int) -> int:
    z = (z + 0x9E3779B97F4A7C15) & _MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK64
    return z ^ (z >> 31)


def repair_default_generator(device) -> None:
    """torch leaves the device's default generator flagged as "capturing" when a capture ends with an error
    (CUDAGraph::capture_end throws before the generators' epilogue runs); every later eager draw -- the loader's
    randperm, a dropout -- then raises "Offset increment outside graph capture encountered unexpectedly". Swap in a
    fresh state object carrying the same seed and offset. Best effort; called on the failure path only."""
    try:
        dev = torch.device(device)
        gen = torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]
        fresh = torch.Generator(device=dev)
        fresh.manual_seed(gen.initial_seed())
        fresh.set_offset(gen.get_offset())
        gen.graphsafe_set_state(fresh.graphsafe_get_state())
    except Exception:  # noqa: BLE001
        pass


class CapturedTrainStep:
    def __init__(self, optimizer, loss_fn: Callable[..., Dict[str, torch.Tensor]], example_batch: Sequence[torch.Tensor]):
        """optimizer: FusedAdam bound to the model's EncoderRuntime; loss_fn(*batch) -> {"total_loss": ..., ...}.
        Call this only after a couple of EAGER steps of the same batch shape (they create the engine handle, tune the GEMM
        variants and run torch's lazy initialisations; BaseTrainer does so with the first batches of an epoch).
        Constructing performs exactly ONE training step on example_batch: the recording executes nothing, the first
        replay is that step; its losses are in `first_losses`."""
        rt = optimizer.runtime
        if rt is None:
            raise ValueError("CapturedTrainStep needs a FusedAdam bound to an EncoderRuntime")
        self.opt, self.rt, self.loss_fn = optimizer, rt, loss_fn
        dev = example_batch[0].device
        self.static_in: List[torch.Tensor] = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in example_batch]
        self._shapes = [(tuple(t.shape), t.dtype) for t in example_batch]
        self._ring = 8                                                             # steps the host may run ahead of the GPU
        self._pinned = torch.zeros(self._ring, 8, dtype=torch.int32).pin_memory()  # rows: [mix0, mix1, mix2, lr, bc1, sqrt(bc2), -, -]
        self._pinned_i = self._pinned.numpy()
        self._pinned_f = self._pinned.view(torch.float32).numpy()
        self._slot_free = [None] * self._ring                                      # event: the slot's copy has been executed
        self._n = 0
        self._dev_words = torch.zeros(8, dtype=torch.int32, device=dev)
        self._dev_hyper = self._dev_words.view(torch.float32)[3:6]
        self.keys: Optional[List[str]] = None
        self.static_out: Optional[torch.Tensor] = None
        # the recorded step must contain the FULL weight refresh (bf16 cast + transposes): every replay follows an update, and
        # a parameter change from outside between replays (load_state_dict, ...) must not meet a stale shadow
        rt._shadow_version = None
        rt._shadow_fresh_ver = None
        self._load(example_batch)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        step_before, counter_before = optimizer._step, rt.step_counter
        self._advance_host_state()
        optimizer._captured_hyper = self._dev_hyper
        launches0 = ops.lib().wm_launch_count()
        try:
            # thread_local: autograd runs the backward launches on its device worker thread; only the recording thread
            # must be barred from capture-unsafe calls
            with CAPTURE_LOCK, torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                ops.step_params_apply(self._dev_words)
                self._body(eager=False)
        except BaseException:
            optimizer._step, rt.step_counter = step_before, counter_before
            optimizer._step_t.fill_(float(step_before))
            for p in optimizer.param_groups[0]["params"]:
                st = optimizer.state.get(p)
                if st and "step" in st and st["step"] is not optimizer._step_t:
                    st["step"] -= 1
            raise
        finally:
            optimizer._captured_hyper = None
        self.kernels_per_replay = int(ops.lib().wm_launch_count() - launches0)  # libwm_b200 kernels inside one replay
        # the recording executes nothing: this replay IS the training step on example_batch
        self.graph.replay()
        self.rt.mark_weights_dirty(shadow_written=self.rt.shadow_ptr() is not None)
        self.first_losses = {k: self.static_out[i] for i, k in enumerate(self.keys)}

    # ------------------------------------------------------------------ pieces
    def matches(self, batch: Sequence[torch.Tensor]) -> bool:
        return len(batch) == len(self._shapes) and all(tuple(t.shape) == s and t.dtype == d for t, (s, d) in zip(batch, self._shapes))

    def _load(self, batch):
        for dst, src in zip(self.static_in, batch):
            dst.copy_(src, non_blocking=True)  # (expands stride-0 mask views into the dense static buffer)

    def _body(self, eager: bool):
        self.opt.zero_grad()
        losses = self.loss_fn(*self.static_in)
        losses["total_loss"].backward()
        self.opt.step()
        if self.keys is None:
            self.keys = list(losses)
        out = torch.stack([losses[k].detach().float().reshape(()) for k in self.keys])
        if not eager:
            self.static_out = out
        return out

    def _advance_host_state(self):
        """What the host would have done in an eager step: next dropout stream, next Adam step, current lr."""
        rt, opt = self.rt, self.opt
        slot = self._n % self._ring
        self._n += 1
        if self._slot_free[slot] is not None:
            self._slot_free[slot].synchronize()  # the copy that last read this slot has run (host at most `ring` steps ahead)
        rt.step_counter += 1
        z = _splitmix((rt.seed ^ (rt.step_counter * 0xD1342543DE82EF95)) & _MASK64)
        z2 = _splitmix(z)
        words = [z & 0xFFFFFFFF, (z >> 32) & 0xFFFFFFFF, z2 & 0xFFFFFFFF]
        for i, w in enumerate(words):
            self._pinned_i[slot, i] = w - (1 << 32) if w >= (1 << 31) else w
        group = opt.param_groups[0]
        b1, b2 = group["betas"]
        opt._step += 1
        k = opt._step
        self._pinned_f[slot, 3] = float(group["lr"])
        self._pinned_f[slot, 4] = 1.0 - b1 ** k
        self._pinned_f[slot, 5] = (1.0 - b2 ** k) ** 0.5
        self._dev_words.copy_(self._pinned[slot], non_blocking=True)  # stream-ordered in front of the replay
        ev = self._slot_free[slot] or torch.cuda.Event()
        ev.record()
        self._slot_free[slot] = ev
        opt._step_t.fill_(float(k))
        for p in group["params"]:  # parameters outside the flat bucket (yield head): per-tensor step counters
            st = opt.state.get(p)
            if st and "step" in st and st["step"] is not opt._step_t:
                st["step"] += 1

    # ------------------------------------------------------------------ one training step
    def __call__(self, *batch: torch.Tensor) -> Dict[str, torch.Tensor]:
        self._load(batch)
        self._advance_host_state()
        self.graph.replay()
        # Adam ran at the end of the replay (and wrote the bf16 shadow): eager forwards (validation) must rebuild the transposes
        self.rt.mark_weights_dirty(shadow_written=self.rt.shadow_ptr() is not None)
        return {k: self.static_out[i] for i, k in enumerate(self.keys)}
