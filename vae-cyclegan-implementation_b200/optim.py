"""FusedAdam: torch.optim.Adam's interface (param_groups, state_dict()/load_state_dict() with the
same keys -- exp_avg, exp_avg_sq, step -- so reference checkpoints interoperate, Networks.py:315-328,
utils.py:22,47) with the update done by ONE multi-tensor kernel launch (csrc/adam.cu) instead of the
foreach kernel sequence of torch/optim/adam.py:457-547.

Gradients live in one flat fp32 buffer (``.grad`` tensors are views into it): zero_grad() is a single
fill, and data-parallel training all-reduces the flat buffer in buckets (dist.py)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import lib as L
from . import ops

CHUNK = 65536


class FusedAdam(torch.optim.Adam):
    def state_dict(self):
        self.finish()
        return super().state_dict()

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, flat_grads=True):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, foreach=False)
        self._flat_grads = flat_grads
        self._flat = None
        self._table = None
        self._table_key = None
        self._dev_state = None      # float32[4] on the device: {step, lr/bc1, sqrt(bc2), -}
        self._dev_step = 0
        self._last = None
        self.grad_scale = 1.0       # data-parallel averaging folded into the update (dist.py sets 1/world)
        self.pre_step_hook = None   # dist.py: starts the all-reduce of the flat gradient buffer (side stream)
        self.pre_update_hook = None  # dist.py: joins the all-reduce right before the Adam kernel
        self.defer = False          # dist.py: step() only starts the all-reduce; finish() joins it and updates, so
                                    # that the exchange overlaps whatever the caller runs in between
        self._deferred = False
        self.overlap = False        # single-process two-optimiser models: step() launches the Adam kernel on a side
                                    # stream so that it overlaps the discriminator backward; finish() joins it
        self._side = None
        self._side_pending = False

    # ---- flat gradient buffer -------------------------------------------------------------
    def _params(self):
        return [p for g in self.param_groups for p in g["params"] if p.requires_grad]

    def flat_grad(self):
        """Allocate (once) the flat gradient buffer and point every .grad into it."""
        ps = self._params()
        if self._flat is None or self._flat.device != ps[0].device:
            total = sum((p.numel() + 3) // 4 * 4 for p in ps)     # keep every view 16-byte aligned
            self._flat = torch.zeros(total, dtype=torch.float32, device=ps[0].device)
            off = 0
            for p in ps:
                p.grad = self._flat[off:off + p.numel()].view_as(p)
                off += (p.numel() + 3) // 4 * 4
        return self._flat

    def zero_grad(self, set_to_none=True):
        self.finish()
        if not self._flat_grads or not self._params() or not self._params()[0].is_cuda:
            return super().zero_grad(set_to_none)
        flat = self.flat_grad()
        ops.zero_(flat)
        off = 0
        for p in self._params():       # re-attach views dropped by an external zero_grad(set_to_none=True)
            if p.grad is None or p.grad.data_ptr() != flat.data_ptr() + 4 * off:
                p.grad = flat[off:off + p.numel()].view_as(p)
            off += (p.numel() + 3) // 4 * 4

    # ---- the update ---------------------------------------------------------------------------
    def _build_table(self, items):
        rows = []
        for p, g, m, v in items:
            n = p.numel()
            for o in range(0, n, CHUNK):
                rows.append((p.data_ptr() + 4 * o, g.data_ptr() + 4 * o, m.data_ptr() + 4 * o, v.data_ptr() + 4 * o,
                             min(CHUNK, n - o)))
        arr = (L.AdamChunk * len(rows))()
        for i, r in enumerate(rows):
            arr[i].p, arr[i].g, arr[i].m, arr[i].v, arr[i].numel = r
        host = torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy())
        return host.to(items[0][0].device), len(rows)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        from . import plan
        plan.flush_grads()          # no-op unless a backward was driven outside autograd
        self.finish()               # a previous deferred step of this optimiser
        if self.pre_step_hook is not None:
            self.pre_step_hook(self)
        if self.defer:
            self._deferred = True
            return loss
        ps = self._params()
        if self.overlap and self.pre_step_hook is None and ps and ps[0].is_cuda:
            if self._side is None:
                self._side = torch.cuda.Stream(device=ps[0].device)
            self._side.wait_stream(torch.cuda.current_stream(ps[0].device))
            with torch.cuda.stream(self._side):
                self._update()
            self._side_pending = True
            return loss
        self._update()
        return loss

    @torch.no_grad()
    def finish(self):
        """Complete a deferred step (join the gradient exchange, run the Adam kernel)."""
        if self._side_pending:
            self._side_pending = False
            torch.cuda.current_stream(self._params()[0].device).wait_stream(self._side)
        if self._deferred:
            self._deferred = False
            self._update()

    def _update(self):
        if self.pre_update_hook is not None:
            self.pre_update_hook(self)
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            items, steps = [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("FusedAdam: CUDA parameters required (no CPU fallback)")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                g = p.grad
                if not g.is_contiguous() or g.dtype != torch.float32:
                    raise RuntimeError("FusedAdam: gradients must be contiguous fp32")
                items.append((p, g, st["exp_avg"], st["exp_avg_sq"]))
                steps.append(st["step"])
            if not items:
                continue
            key = tuple((p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()) for p, g, m, v in items)
            if key != self._table_key or len(self.param_groups) > 1:
                self._table = self._build_table(items)
                self._table_key = key
            before = int(steps[0].item())
            if any(int(s.item()) != before for s in steps[1:]):
                raise RuntimeError("FusedAdam: parameters of one group must share the step count")
            dev = items[0][0].device
            if self._dev_state is None or self._dev_state.device != dev:
                self._dev_state = torch.zeros(4, dtype=torch.float32, device=dev)
                self._dev_step = 0
            if self._dev_step != before:          # first use / after load_state_dict: resync the device counter
                self._dev_state[0:1].fill_(float(before))
            table, nchunks = self._table
            ops.adam_multi(table, nchunks, self._dev_state, group["lr"], beta1, beta2, group["eps"], self.grad_scale)
            self._dev_step = before + 1
            self._last = (steps, [p for p, _, _, _ in items])
            self._bump(steps, self._last[1])

    @staticmethod
    def _bump(steps, params):
        for s in steps:
            s += 1
        for p in params:        # the kernel wrote p behind autograd's back: invalidate packed copies
            p._vcg_epoch = getattr(p, "_vcg_epoch", 0) + 1

    def capture_rollback(self):
        """step() ran under CUDA-graph capture: its kernels were recorded, not executed; undo the host bump."""
        if self._last is not None:
            for st in self._last[0]:
                st -= 1
            self._dev_step -= 1

    def note_replay(self):
        """A captured CUDA graph containing this optimiser's step was replayed: the device-side counter
        and the weights advanced; mirror that in the host-side state (state_dict 'step', cache epochs)."""
        if self._last is not None:
            self._bump(*self._last)
            self._dev_step += 1
