"""Host side of the encoder: flat parameter storage, the C engine handle and the autograd glue.

`EncoderRuntime` belongs to one WeatherBERT / WeatherFormer module (src/pretraining/models/weatherbert.py
in the reference). It
  * keeps every parameter of the module as a VIEW into one flat fp32 buffer laid out as
    wm_encoder_param_layout() says (reference named_parameters() order), and the gradients as views into
    a second flat buffer -- so the optimiser and the gradient all-reduce work on whole buckets;
  * owns the device workspace and creates/caches `wm_encoder` handles per (batch, seq_len);
  * exposes forward/backward as ONE torch.autograd.Function (`_EncoderFn`) that only launches kernels
    of libwm_b200.so. There is no PyTorch implementation of the math here and no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import check, lib

_PER_LAYER = [
    "self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight", "self_attn.out_proj.bias",
    "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias",
    "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias",
]


def encoder_param_names(num_layers: int) -> List[str]:
    """Parameter names in the order of the flat buffer (== reference named_parameters() order)."""
    names = ["in_proj.weight", "in_proj.bias"]
    for l in range(num_layers):
        names += [f"transformer_encoder.layers.{l}.{n}" for n in _PER_LAYER]
    names += ["out_proj.weight", "out_proj.bias"]
    return names


class EncoderRuntime:
    def __init__(self, module: nn.Module, num_heads: int, dropout_p: float = 0.1, ln_eps: float = 1e-5):
        self.module_ref = module
        self.num_heads = int(num_heads)
        self.dropout_p = float(dropout_p)
        self.ln_eps = float(ln_eps)
        self.flat_params: Optional[torch.Tensor] = None
        self.flat_grads: Optional[torch.Tensor] = None
        self.offsets: List[int] = []
        self.workspace: Optional[torch.Tensor] = None
        self._handles: Dict[Tuple[int, int, bool], int] = {}
        self._active: Optional[int] = None
        self._shadow_version = None
        self._shadow_fresh_ver = None
        self.step_counter = 0
        self.base_seed = 0x5EED_2002
        self.seed = self.base_seed  # BucketedDataParallel folds the rank in
        # called as hook(lo_float_offset, hi_float_offset) right after the gradients of that slice of
        # the flat buffer are final (data-parallel bucket all-reduce); set by BucketedDataParallel
        self.grad_ready_hook: Optional[Callable[[int, int], None]] = None
        self.layers_per_bucket = 1

    # ------------------------------------------------------------------ configuration / storage
    def _dims(self):
        m = self.module_ref
        D = m.in_proj.out_features
        L = len(m.transformer_encoder.layers)
        FF = m.transformer_encoder.layers[0].linear1.out_features
        return D, L, FF, m.out_proj.out_features, m.weather_dim

    def _config(self, B: int, S: int, eval_only: bool = False) -> _lib.EncoderConfig:
        D, L, FF, out_dim, F = self._dims()
        return _lib.EncoderConfig(B, S, F, D, self.num_heads, L, FF, out_dim, self.dropout_p, self.ln_eps, int(eval_only))

    def named_flat_params(self):
        D, L, _, _, _ = self._dims()
        params = dict(self.module_ref.named_parameters())
        return [(n, params[n]) for n in encoder_param_names(L)]

    def _views_ok(self) -> bool:
        if self.flat_params is None:
            return False
        base = self.flat_params.data_ptr()
        for (_, p), off in zip(self._named, self.offsets):
            if p.data_ptr() != base + 4 * off:
                return False
        return True

    def ensure_flat(self, device: torch.device):
        """(Re)build the flat buffers if the module's parameters are not (or no longer) views into them --
        e.g. right after construction, .to(device), load_pretrained() or load_state_dict(assign=True)."""
        self._named = self.named_flat_params()
        if self.flat_params is not None and self.flat_params.device == device and self._views_ok():
            return
        cfg = self._config(1, 8)
        total = lib().wm_encoder_param_count(C.byref(cfg))
        if total <= 0:
            raise ValueError("this model shape is not supported by the sm_100a encoder "
                             "(need D % 8 == 0, head_dim % 4 == 0 and <= 48, seq_len <= 384)")
        n = lib().wm_encoder_param_layout(C.byref(cfg), None, 0)
        offs = (C.c_int64 * n)()
        lib().wm_encoder_param_layout(C.byref(cfg), offs, n)
        self.offsets = [int(o) for o in offs]
        if len(self.offsets) != len(self._named):
            raise RuntimeError("parameter layout mismatch between the module and libwm_b200")
        flat = torch.zeros(total, dtype=torch.float32, device=device)
        grads = torch.zeros(total, dtype=torch.float32, device=device)
        with torch.no_grad():
            for (name, p), off in zip(self._named, self.offsets):
                if p.dtype != torch.float32:
                    raise TypeError(f"{name}: master weights must be float32")
                flat[off:off + p.numel()].copy_(p.detach().reshape(-1))
                p.data = flat[off:off + p.numel()].view(p.shape)
                p.grad = None
        self.flat_params, self.flat_grads = flat, grads
        self._shadow_version = None

    def grad_view(self, index: int) -> torch.Tensor:
        p = self._named[index][1]
        off = self.offsets[index]
        return self.flat_grads[off:off + p.numel()].view(p.shape)

    def layer_slice(self, layer_lo: int, layer_hi: int) -> Tuple[int, int]:
        """[lo, hi) float offsets of the flat buffer covering encoder layers layer_lo .. layer_hi-1."""
        lo = self.offsets[2 + 12 * layer_lo]
        hi = self.offsets[2 + 12 * layer_hi]
        return lo, hi

    # ------------------------------------------------------------------ engine handles
    def _handle(self, B: int, S: int, eval_only: bool = False) -> int:
        """C handle for this (batch, seq_len). A forward that no backward follows (torch.no_grad(): validation,
        inference) reuses the training handle of the same shape if one exists, else gets an eval-only handle whose
        workspace holds one layer of activations instead of all of them (include/wm_b200.h: eval_only)."""
        if eval_only and (B, S, False) in self._handles:
            eval_only = False
        key = (B, S, eval_only)
        cfg = self._config(B, S, eval_only)
        need = lib().wm_encoder_workspace_bytes(C.byref(cfg))
        if need == 0:
            raise ValueError(f"unsupported encoder shape B={B} S={S}")
        if self.workspace is None or self.workspace.numel() < need or self.workspace.device != self.flat_params.device:
            self._destroy_handles()
            self.workspace = torch.empty(need, dtype=torch.uint8, device=self.flat_params.device)
        if key not in self._handles:
            out = C.c_void_p()
            with torch.cuda.device(self.flat_params.device):
                check(lib().wm_encoder_create(C.byref(cfg), self.workspace.data_ptr(), self.workspace.numel(),
                                              C.byref(out)), "wm_encoder_create")
            self._handles[key] = out.value
            # once per shape: pick the faster (bit-identical) kernel variant for each of the layer's GEMM call sites
            D, _, FF, _, _ = self._dims()
            ops.tune_gemm_sites(B * S, D, FF, self.dropout_p, self.flat_params.device)
        return self._handles[key]

    def _destroy_handles(self):
        for h in self._handles.values():
            lib().wm_encoder_destroy(h)
        self._handles.clear()
        self._shadow_version = None
        self._shadow_fresh_ver = None

    def __del__(self):
        try:
            self._destroy_handles()
        except Exception:
            pass

    def mark_weights_dirty(self, shadow_written: bool = False):
        """The fp32 parameters changed behind torch's back. shadow_written: the caller (FusedAdam) also wrote the bf16
        shadow through shadow_ptr(), so the next refresh only has to rebuild the transposed copies -- valid as long as
        nobody else touches the parameters in between (their version counters are remembered)."""
        self._shadow_version = None
        self._shadow_fresh_ver = sum(p._version for _, p in self._named) if shadow_written else None

    def shadow_ptr(self) -> Optional[int]:
        """Device address of the bf16 parameter shadow (None before the first handle exists). All handles carve the same
        workspace prefix, so any of them answers."""
        for h in self._handles.values():
            return lib().wm_encoder_shadow(h)
        return None

    def _refresh_if_needed(self, handle: int):
        # all handles carve the same workspace, so the bf16 shadows written through one are valid for
        # another only if the layout prefix is identical -- it is (weights are carved first)
        # in-place updates by any torch optimiser bump the parameters' version counters; the fused Adam
        # kernel writes behind torch's back and calls mark_weights_dirty() instead
        ver = sum(p._version for _, p in self._named)
        if self._shadow_version != ver:
            if getattr(self, "_shadow_fresh_ver", None) == ver:  # the optimiser wrote the bf16 shadow itself
                check(lib().wm_encoder_refresh_transposes(handle, self.flat_params.data_ptr(), ops._stream()),
                      "wm_encoder_refresh_transposes")
            else:
                check(lib().wm_encoder_refresh_weights(handle, self.flat_params.data_ptr(), ops._stream()),
                      "wm_encoder_refresh_weights")
            self._shadow_version = ver
            self._shadow_fresh_ver = None

    # ------------------------------------------------------------------ forward / backward
    def forward(self, weather, coords, year, mask, training: bool, save_for_backward: bool = True) -> torch.Tensor:
        """Returns the padded raw head output: fp32 [B, S, 32|64]. save_for_backward=False is the lean schedule for
        torch.no_grad() callers (BaseTrainer._validate_epoch: model.eval() + no_grad, reference base_trainer.py:262-285)."""
        B, S, F = weather.shape
        dev = weather.device
        self.ensure_flat(dev)
        h = self._handle(B, S, eval_only=not save_for_backward)
        with torch.cuda.device(dev):
            self._refresh_if_needed(h)
            m = self.module_ref
            outP = 32 if m.out_proj.out_features <= 32 else 64
            y = torch.empty((B, S, outP), dtype=torch.float32, device=dev)
            mask, msb, mss = ops.mask_strides(mask)
            self.step_counter += 1
            pe = m.positional_encoding.pos_encoding
            check(lib().wm_encoder_forward(h, self.flat_params.data_ptr(), weather.data_ptr(), mask.data_ptr(), msb, mss,
                                           year.data_ptr(), coords.data_ptr(), pe.data_ptr(), y.data_ptr(),
                                           int(training), int(save_for_backward), self.seed, self.step_counter,
                                           ops._stream()),
                  "wm_encoder_forward")
        self._active = h
        return y

    def bucket_schedule(self) -> List[Tuple[str, int, int, int, int]]:
        """Order in which backward finalises slices of the flat gradient buffer (top of the network first):
        (kind, layer_hi, layer_lo, float_lo, float_hi) with kind in {'head', 'layers', 'embed'}."""
        L = len(self.module_ref.transformer_encoder.layers)
        sched = [("head", L, L, self.offsets[-2], self.flat_grads.numel())]
        hi = L
        while hi > 0:
            lo = max(0, hi - self.layers_per_bucket)
            sched.append(("layers", hi, lo) + self.layer_slice(lo, hi))
            hi = lo
        sched.append(("embed", 0, 0, 0, self.offsets[2]))
        return sched

    def backward(self, handle: int, dy_bf16: torch.Tensor):
        """dy_bf16: bf16 [B*S, 32|64] gradient of the loss w.r.t. the padded head output. Fills flat_grads
        (and each parameter's .grad as a view of it); fires grad_ready_hook per bucket, top layers first."""
        fg, fp = self.flat_grads.data_ptr(), self.flat_params.data_ptr()
        st = ops._stream()
        hook = self.grad_ready_hook
        with torch.cuda.device(self.flat_params.device):
            for kind, hi, lo, f_lo, f_hi in self.bucket_schedule():
                if kind == "head":
                    check(lib().wm_encoder_backward_head(handle, dy_bf16.data_ptr(), fg, st), "wm_encoder_backward_head")
                elif kind == "layers":
                    check(lib().wm_encoder_backward_layers(handle, fp, hi, lo, fg, st), "wm_encoder_backward_layers")
                else:
                    check(lib().wm_encoder_backward_embed(handle, fg, st), "wm_encoder_backward_embed")
                if hook:
                    hook(f_lo, f_hi)

    def grads_are_live(self) -> bool:
        """True if some parameter's .grad already IS its slice of the flat gradient buffer, i.e. a backward ran and
        nobody called zero_grad() since: the next backward has to accumulate, as torch does."""
        if self.flat_grads is None:
            return False
        return any(p.grad is not None and p.grad.data_ptr() == self.grad_view(i).data_ptr()
                   for i, (_, p) in enumerate(self._named))

    def publish_grads(self):
        for i, (_, p) in enumerate(self._named):
            if not p.requires_grad:
                continue
            g = self.grad_view(i)
            if p.grad is None or p.grad.data_ptr() == g.data_ptr():
                p.grad = g
            else:  # a foreign gradient tensor (e.g. produced through torch ops): add ours to it, out of place
                p.grad = p.grad + g


# The fused loss heads hand their bf16 gradient to the encoder backward directly: autograd insists that the
# gradient of the fp32 head output be an fp32 tensor of the same shape, which would cost a bf16 -> fp32 -> bf16 round
# trip of [B*S, 64] values per step. The loss backward therefore returns a stride-0 expand of a
# zero scalar and parks the real gradient on the runtime, keyed by the forward step it belongs to.
_ZERO_SCALARS: Dict[torch.device, torch.Tensor] = {}


def _placeholder_grad(shape, device) -> torch.Tensor:
    z = _ZERO_SCALARS.get(device)
    if z is None:
        z = _ZERO_SCALARS[device] = torch.zeros((), dtype=torch.float32, device=device)
    return z.expand(shape)


def _is_placeholder(t: torch.Tensor) -> bool:
    z = _ZERO_SCALARS.get(t.device)
    return z is not None and t.data_ptr() == z.data_ptr() and all(s == 0 for s in t.stride())


class _EncoderFn(torch.autograd.Function):
    """y_pad = encoder(weather, coords, year, mask); backward launches the C backward schedule and
    publishes parameter gradients as views of the flat gradient buffer (returned grads are None)."""

    @staticmethod
    def forward(ctx, anchor, runtime: EncoderRuntime, weather, coords, year, mask, training):
        y = runtime.forward(weather, coords, year, mask, training, save_for_backward=True)
        ctx.runtime = runtime
        ctx.handle = runtime._active
        ctx.step = runtime.step_counter
        return y

    @staticmethod
    def backward(ctx, dy):
        rt: EncoderRuntime = ctx.runtime
        if ctx.step != rt.step_counter:
            raise RuntimeError("encoder backward called after another forward reused the activation workspace")
        B, S, P = dy.shape
        side = rt.__dict__.pop("_side_dy", None)
        if side is not None and side[0] != ctx.step:
            side = None
        if side is not None and _is_placeholder(dy):
            dyb = side[1]  # bf16 [B*S, P] straight from the fused loss kernel
        else:
            dyf = dy.reshape(B * S, P)
            if side is not None:  # the head output fed the fused loss AND other torch ops: add both gradients
                dyf = dyf + side[1].float()
            dyb = dyf.to(torch.bfloat16).contiguous()
        backup = rt.flat_grads.clone() if rt.grads_are_live() else None  # second backward without zero_grad()
        rt.backward(ctx.handle, dyb)
        if backup is not None:
            rt.flat_grads.add_(backup)
        rt.publish_grads()
        return (None,) * 7


def encoder_apply(runtime: EncoderRuntime, weather, coords, year, mask, training: bool) -> torch.Tensor:
    for t in (weather, coords, year, mask):
        if not t.is_cuda:
            raise RuntimeError("WeatherBERT/WeatherFormer (weathermodel_b200) run on CUDA sm_100a only; "
                               "the reference PyTorch model is the CPU path")
    weather = weather.contiguous().float()
    coords = coords.contiguous().float()
    year = year.contiguous().float()
    runtime.ensure_flat(weather.device)
    if torch.is_grad_enabled() and any(p.requires_grad for _, p in runtime._named):
        # A fresh leaf per call makes autograd track the node. (A parameter in this role drags its AccumulateGrad node
        # along: that node remembers the stream of the step it was created in for as long as any earlier graph is alive,
        # and a backward recorded into a CUDA graph then fails on a dependency to that uncaptured stream.)
        anchor = torch.empty((), dtype=torch.float32, device=weather.device).requires_grad_()
        y = _EncoderFn.apply(anchor, runtime, weather, coords, year, mask, training)
        y._wm_src = (runtime, runtime.step_counter)  # lets a fused loss park its bf16 gradient for this forward
        return y
    return runtime.forward(weather, coords, year, mask, training, save_for_backward=False)


# ---------------------------------------------------------------------------------------------
# fused loss heads on the raw (padded) encoder output
# ---------------------------------------------------------------------------------------------
def _hand_over(src, dy_bf16: torch.Tensor, shape, device):
    """Park the bf16 loss gradient for the encoder backward of the forward it came from; the fp32 copy autograd
    would otherwise demand is made only when the head output did not come straight from the encoder."""
    if src is not None:
        rt, step = src
        cur = rt.__dict__.get("_side_dy")
        if rt.step_counter == step and (cur is None or cur[0] != step):  # (a second fused loss on the same output adds in fp32)
            rt.__dict__["_side_dy"] = (step, dy_bf16)
            return _placeholder_grad(shape, device)
    return dy_bf16.view(shape).float()


class _BertLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_pad, weather, mask, src):
        B, S, P = y_pad.shape
        F = weather.shape[-1]
        w2, m2 = weather.reshape(B * S, F).contiguous(), mask.reshape(B * S, F).contiguous()
        out, scratch = ops.loss_bert_value(y_pad.view(B * S, P), w2, m2)
        ctx.save_for_backward(y_pad, w2, m2, scratch)
        ctx.src = src
        return out[0]

    @staticmethod
    def backward(ctx, g):
        y_pad, w2, m2, scratch = ctx.saved_tensors
        B, S, P = y_pad.shape
        dy = ops.loss_bert_grad(y_pad.view(B * S, P), w2, m2, scratch, g, ld_grad=P)
        return _hand_over(ctx.src, dy, (B, S, P), y_pad.device), None, None, None


class _FormerLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_pad, weather, mask, beta, src):
        B, S, P = y_pad.shape
        out, scratch = ops.loss_former_value(y_pad.view(B * S, P), weather, mask, beta)
        ctx.save_for_backward(y_pad, weather, mask, scratch)
        ctx.beta, ctx.src = beta, src
        ctx.mark_non_differentiable(out)
        return out[0], out

    @staticmethod
    def backward(ctx, g, _g_all):
        y_pad, weather, mask, scratch = ctx.saved_tensors
        B, S, P = y_pad.shape
        dy = ops.loss_former_grad(y_pad.view(B * S, P), weather, mask, ctx.beta, scratch, g, ld_grad=P)
        return _hand_over(ctx.src, dy, (B, S, P), y_pad.device), None, None, None, None


def bert_masked_mse(y_pad, weather, mask) -> torch.Tensor:
    """mean((weather[mask] - y[mask])**2) without the boolean gathers (weatherbert_trainer.py:55-60)."""
    return _BertLossFn.apply(y_pad, weather.contiguous().float(), mask, getattr(y_pad, "_wm_src", None))


def former_elbo(y_pad, weather, mask, beta: float) -> Dict[str, torch.Tensor]:
    """ELBO of weatherformer_trainer.py:68-111 straight from the raw head output [mu | logvar | pad]."""
    total, allv = _FormerLossFn.apply(y_pad, weather.contiguous().float(), mask, float(beta), getattr(y_pad, "_wm_src", None))
    return {"total_loss": total, "reconstruction": allv[1], "kl_term": allv[2]}


# ---------------------------------------------------------------------------------------------
# fused crop-yield head on the raw (padded) encoder output
# ---------------------------------------------------------------------------------------------
class _YieldHeadFn(torch.autograd.Function):
    """(pred, z) = head(y_pad, weather, mask, eps, y_past; 8 head parameters) -- one kernel forward, one backward
    (+ a fixed-order fold of the per-sequence parameter gradients). z is returned for API parity only
    (WeatherFormerYieldModel.forward returns it; no trainer differentiates through it)."""

    @staticmethod
    def forward(ctx, y_pad, weather, mask, eps, y_past, is_former, *params):
        pred, z = ops.yield_head_fwd(y_pad, weather, mask, eps, y_past, params, is_former)
        ctx.save_for_backward(y_pad, mask, eps if eps is not None else y_pad.new_empty(0), z, y_past, *params)
        ctx.is_former = is_former
        ctx.mark_non_differentiable(z)
        return pred, z

    @staticmethod
    def backward(ctx, dpred, _dz):
        y_pad, mask, eps, z, y_past, *params = ctx.saved_tensors
        dy, grads = ops.yield_head_bwd(dpred, y_pad, mask, eps if ctx.is_former else None, z, y_past, params, ctx.is_former)
        return (dy, None, None, None, None, None, *grads)


def yield_head(y_pad, weather, mask, eps, y_past, params, is_former: bool):
    """Imputation (+ reparameterisation), softmax pooling over the sequence and the yield MLP of the reference's
    yield models (src/crop_yield/models/weatherbert_yield_model.py:40-67, weatherformer_yield_model.py:58-60)."""
    return _YieldHeadFn.apply(y_pad, weather.contiguous().float(), mask, eps, y_past, bool(is_former), *params)
