"""torch.autograd.Function wrappers: the only place where autograd meets the C ABI.

PlanFunction runs a whole network plan (plan.py) forward / backward; the loss Functions wrap the
fused loss kernels.  Parameter gradients are accumulated straight into ``param.grad`` by the plan's
backward (the wgrad GEMM output is unpacked into it), so parameters are not Function inputs; a
per-device scalar *anchor* that requires grad keeps the node in the autograd graph when no tensor
input does (first-level generator passes on data)."""
from __future__ import annotations

import torch

from . import lib as L
from . import ops

_anchors = {}


def _anchor(device):
    a = _anchors.get(device)
    if a is None:
        a = _anchors[device] = torch.zeros((), device=device, requires_grad=True)
    return a


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor -- this implementation has no CPU fallback "
                           "(the reference's CPU path is timed by bench.py --impl reference)")


class PlanFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, keep, eps_list, anchor, *inputs):
        outs, run = plan.run_forward(inputs, eps_list, keep)
        ctx.plan, ctx.run = plan, run
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        if ctx.run is None:
            raise RuntimeError("backward through a plan that was run without saving activations")
        from .plan import _STATE
        need = ctx.needs_input_grad[4:]
        if not _STATE.get("input_grads", True):      # discriminator step: stop at the network inputs
            need = tuple(False for _ in need)
        from .plan import queue_flush
        from . import lanes
        lanes.note_stream(torch.cuda.current_stream())      # autograd runs this node on its forward's stream (a lane?)
        gin = ctx.plan.run_backward(ctx.run, grads, need, defer_flush=True)
        queue_flush()       # weight / bias gradient accumulators -> .grad once, after the whole backward pass
        return (None, None, None, None, *gin)


def run_plan(plan, inputs, eps_list=()):
    for x in inputs:
        require_cuda(x, "network input")
        if x.device.index != torch.cuda.current_device():
            # every kernel is enqueued on the CURRENT device's current stream (lib.stream_ptr): tensors of another GPU
            # would be dereferenced in the wrong context
            raise RuntimeError(f"network input lives on {x.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                               "call torch.cuda.set_device() (one process per GPU) first")
    keep = torch.is_grad_enabled()
    return PlanFunction.apply(plan, keep, tuple(eps_list), _anchor(inputs[0].device), *inputs)


# ------------------------------------------------------------------------------ fused losses
def _f32c(t):
    t = t if t.dtype == torch.float32 else t.float()
    return t if t.is_contiguous() else t.contiguous()


# Accumulators of the fused loss kernels (they add block partial sums into a zeroed scalar).  Inside a composite's
# step (Networks._Composite._xy ... _items) all of them come from ONE arena zeroed by one launch at the start of the
# step -- a zero-fill launch per loss call sat on the serial section between the forward and the backward pass (12
# launches of a VAE-CycleGAN step).  Outside a step every call zeroes its own.
_ARENA = {"buf": None, "pos": 0}
_ARENA_FLOATS = 512


def begin_step(device):
    _ARENA["buf"] = ops.zero_(torch.empty(_ARENA_FLOATS, dtype=torch.float32, device=device))
    _ARENA["pos"] = 0


def end_step():
    _ARENA["buf"] = None


def _accumulator(device):
    buf, i = _ARENA["buf"], _ARENA["pos"]
    if buf is None or buf.device != device or i + 4 > buf.numel():
        return ops.zero_(torch.empty(4, dtype=torch.float32, device=device))
    _ARENA["pos"] = i + 4
    return buf[i:i + 4]


class L1MeanFn(torch.autograd.Function):
    """mean |a - b| (nn.L1Loss, Losses.py:21-24): value and sign gradient from one kernel launch."""

    @staticmethod
    def forward(ctx, a, b):
        require_cuda(a, "L1 loss")
        a, b = _f32c(a.detach()), _f32c(b.detach())
        out = _accumulator(a.device)
        need = any(ctx.needs_input_grad)
        g = torch.empty_like(a) if need else None
        ops.l1_fwd_bwd(a, b, out[0:1], g, 1.0 / a.numel())
        ctx.g = g
        return out[0] / a.numel()

    @staticmethod
    def backward(ctx, go):
        g = ctx.g * go
        return (g if ctx.needs_input_grad[0] else None, -g if ctx.needs_input_grad[1] else None)


class MseConstFn(torch.autograd.Function):
    """mean (d - target)^2 against a constant (LSGAN terms, Losses.py:80-81, 99-100)."""

    @staticmethod
    def forward(ctx, d, target):
        require_cuda(d, "GAN loss")
        dd = _f32c(d.detach())
        out = _accumulator(d.device)
        g = torch.empty_like(dd) if ctx.needs_input_grad[0] else None
        ops.mse_const_fwd_bwd(dd, target, out[0:1], g, 1.0 / dd.numel())
        ctx.g = g
        return out[0] / dd.numel()

    @staticmethod
    def backward(ctx, go):
        return ctx.g * go, None


class KlFn(torch.autograd.Function):
    """-0.5 * mean(1 + clamp(lv) - mu^2 - exp(clamp(lv)))  (Losses.py:115-121)."""

    @staticmethod
    def forward(ctx, mu, lv):
        require_cuda(mu, "KL loss")
        m, l = _f32c(mu.detach()), _f32c(lv.detach())
        out = _accumulator(m.device)
        need = any(ctx.needs_input_grad)
        gm = torch.empty_like(m) if need else None
        gl = torch.empty_like(l) if need else None
        ops.kl_fwd_bwd(m, l, out[0:1], gm, gl, 1.0 / m.numel())
        ctx.gm, ctx.gl = gm, gl
        return out[0] * (-0.5 / m.numel())

    @staticmethod
    def backward(ctx, go):
        return ctx.gm * go, ctx.gl * go
