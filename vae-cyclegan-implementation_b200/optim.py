"""FusedAdam: torch.optim.Adam's interface (param_groups, state_dict()/load_state_dict() with the
same keys -- exp_avg, exp_avg_sq, step -- so reference checkpoints interoperate, Networks.py:315-328,
utils.py:22,47) with the update done by multi-tensor kernel launches (csrc/adam.cu) instead of the
foreach kernel sequence of torch/optim/adam.py:457-547.

Layout.  Gradients live in one flat fp32 buffer (``.grad`` tensors are views into it) and so do the two
Adam moments.  The parameters are grouped into *buckets*: the convolutions of one generator /
discriminator (an "owner") in the order in which the backward pass finishes them -- decoder first,
encoder last -- cut every ~24 M parameters.  A bucket is a contiguous range of the flat buffers.

Overlap (the "bucketed, overlapped with backward" exchange of SURVEY.md 8e).  Inside
``with optimizer.track({owner: passes})`` the plan's backward reports every weight-gradient GEMM it
issues; as soon as the last pass that touches a bucket has issued its last contribution, the bucket's
tail runs on the optimiser's side stream, behind exactly those GEMMs (CUDA events):

    accumulators -> .grad (wunpack_multi)  ->  [fp32 -> bf16 wire cast, NCCL all-reduce]  ->  Adam  ->  filter re-pack

while the main stream / lanes keep running the rest of the backward pass and the discriminator step.
``step()`` only completes what has not run yet, ``finish()`` joins the side stream.  The Adam kernel
zeroes the fp32 gradient it consumed, so the next ``zero_grad()`` costs nothing."""
from __future__ import annotations

import contextlib

import numpy as np
import torch

from . import lib as L
from . import ops

CHUNK = 65536
BUCKET_PARAMS = 24 << 20


class _Bucket:
    def __init__(self, owner, holders, params):
        self.owner, self.holders, self.params = owner, holders, params
        self.lo = self.hi = 0
        self.tables = {}          # wire mode -> (key, device table, nchunks)
        self.events = {}          # stream handle -> event behind the last contribution issued on that stream
        self.fired = False
        self.remaining = 0


class _Tracker:
    """Counts the weight-gradient contributions of a tracked backward pass (plan.run_backward calls
    contributed(holder) after it issued the holder's weight-gradient GEMM)."""

    def __init__(self, opt, expect):
        self.opt = opt
        self.count = {}
        self.buckets = []
        for b in opt._buckets:
            n = expect.get(b.owner) if b.owner is not None else None
            if not n or not b.holders:
                continue
            b.remaining, b.events, b.fired = len(b.holders), {}, False
            self.buckets.append(b)
            for h in b.holders:
                self.count[id(h)] = [int(n), b]

    def tracks(self, holder):
        return id(holder) in self.count

    def contributed(self, holder, extra_stream=None):
        """extra_stream: the side stream the weight-gradient GEMM was issued on, when it is not the current one"""
        c = self.count.get(id(holder))
        if c is None:
            return
        b = c[1]
        if b.fired:
            raise RuntimeError("FusedAdam.track: a weight gradient arrived after its bucket was handed to the optimiser "
                               "(more backward passes than announced)")
        for st in (torch.cuda.current_stream(), extra_stream):
            if st is not None:
                ev = torch.cuda.Event()
                ev.record(st)
                b.events[st.cuda_stream] = ev
        c[0] -= 1
        if c[0] == 0:
            b.remaining -= 1
            if b.remaining == 0:
                self.opt._process(b)

    def finish(self):
        for b in self.buckets:
            if not b.fired:
                self.opt._process(b)


class FusedAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, flat_grads=True):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, foreach=False)
        self._flat_grads = flat_grads
        self._flat = self._m = self._v = self._wire = None
        self._offsets = {}
        self._buckets = None
        self._owners = None
        self._dev_state = None      # float32[4] on the device: {step, lr/bc1, sqrt(bc2), -}
        self._dev_step = 0
        self._last = None
        self._ticked = False
        self._zeroed = False        # the flat gradient buffer is known to hold zeros
        self.keep_grads = False     # True: .grad survives step() (tests / tools that read gradients after a step)
        self.grad_scale = 1.0       # data-parallel averaging folded into the update (dist.py sets 1/world)
        self.sync = None            # dist.GradSync when data-parallel
        self.overlap = True         # run the bucket tails on side streams (joined by finish())
        self.side_streams = 3       # buckets go round-robin over this many streams: the all-reduce of one bucket (NVLink)
                                    # overlaps the unpack / Adam / re-pack of its neighbours (HBM)
        self._sides = []
        self._sides_used = set()
        self._next_side = 0
        self._tick_event = None

    # ---- buckets ----------------------------------------------------------------------------
    def set_owners(self, owners, bucket_params=BUCKET_PARAMS):
        """owners: the modules (generators / discriminators) whose passes the training step announces to track().
        Their convolutions are bucketed in backward-completion order (reverse definition order)."""
        self._owners = list(owners)
        self._bucket_params = bucket_params
        self._buckets = None
        self._flat = None

    def _params(self):
        return [p for g in self.param_groups for p in g["params"] if p.requires_grad]

    def _build_buckets(self):
        mine = {id(p) for p in self._params()}
        taken = set()
        buckets = []
        for owner in self._owners or []:
            holders = [m for m in owner.modules() if getattr(m, "_vcg_holder", False)]
            cur_h, cur_p, cur_n = [], [], 0
            for h in reversed(holders):
                ps = [p for p in h.parameters(recurse=False) if id(p) in mine and id(p) not in taken]
                if not ps:
                    continue
                n = sum(p.numel() for p in ps)
                if cur_h and cur_n + n > self._bucket_params:
                    buckets.append(_Bucket(owner, cur_h, cur_p))
                    cur_h, cur_p, cur_n = [], [], 0
                cur_h.append(h)
                cur_p.extend(ps)
                cur_n += n
                taken.update(id(p) for p in ps)
            if cur_h:
                buckets.append(_Bucket(owner, cur_h, cur_p))
        rest = [p for p in self._params() if id(p) not in taken]
        if rest or not buckets:
            # parameters outside the announced owners (or no owners at all): one bucket, flushed by plan.flush_grads()
            buckets.append(_Bucket(None, None, rest))
        self._buckets = buckets

    def flat_grad(self):
        """Allocate (once) the flat gradient / moment buffers and point every .grad into the gradient buffer."""
        ps = self._params()
        if self._flat is None or self._flat.device != ps[0].device:
            self._build_buckets()
            dev = ps[0].device
            off = 0
            self._offsets = {}
            for b in self._buckets:
                b.lo = off
                for p in b.params:
                    self._offsets[id(p)] = off
                    off += (p.numel() + 3) // 4 * 4           # keep every view 16-byte aligned
                b.hi = off
                b.tables = {}
            self._flat = torch.zeros(off, dtype=torch.float32, device=dev)
            # moments: allocated (and zero-filled) HERE, on the calling stream, before any bucket tail can run on a side
            # stream -- the tails of different buckets run on different streams and must not race with a lazy fill
            self._m, self._v, self._wire = torch.zeros_like(self._flat), torch.zeros_like(self._flat), None
            self._zeroed = True
            for p in ps:
                o = self._offsets[id(p)]
                p.grad = self._flat[o:o + p.numel()].view_as(p)
                p._vcg_opt = self
        return self._flat

    def wire_buffer(self):
        """bf16 image of the flat gradient buffer: what travels over NVLink (dist.GradSync)."""
        if self._wire is None or self._wire.device != self._flat.device:
            # (uninitialised on purpose: every slice is written by the cast of its bucket before anything reads it, and a
            #  fill launched from one bucket's stream would race with the casts of the others)
            self._wire = torch.empty(self._flat.numel(), dtype=torch.bfloat16, device=self._flat.device)
        return self._wire

    def reduced_grad(self):
        """the gradient the last update consumed (data-parallel: the all-reduced sum), fp32 copy"""
        src = self._wire if (self.sync is not None and self.sync.world > 1 and self.sync.wire == "bf16") else self._flat
        return src.float().clone()

    def mark_dirty(self):
        self._zeroed = False

    def zero_grad(self, set_to_none=True):
        self.finish()
        ps = self._params()
        if not self._flat_grads or not ps or not ps[0].is_cuda:
            return super().zero_grad(set_to_none)
        flat = self.flat_grad()
        if not self._zeroed:
            ops.zero_(flat)
            self._zeroed = True
        for p in ps:       # re-attach views dropped by an external zero_grad(set_to_none=True)
            o = self._offsets[id(p)]
            if p.grad is None or p.grad.data_ptr() != flat.data_ptr() + 4 * o:
                p.grad = flat[o:o + p.numel()].view_as(p)

    # ---- tracking -------------------------------------------------------------------------------
    @contextlib.contextmanager
    def track(self, expect):
        """expect: {owner module: number of backward passes that will produce its weight gradients}."""
        ps = self._params()
        if not (self.overlap and self._flat_grads and ps and ps[0].is_cuda and self._owners):
            yield None
            return
        from . import plan
        self.flat_grad()
        tr = _Tracker(self, expect)
        plan._TRACKERS.append(tr)
        try:
            yield tr
        finally:
            plan._TRACKERS.remove(tr)
        tr.finish()

    # ---- the update -----------------------------------------------------------------------------
    def _stream(self, dev):
        if not self._sides or self._sides[0].device != dev:
            self._sides = [torch.cuda.Stream(device=dev) for _ in range(max(1, self.side_streams))]
        s = self._sides[self._next_side % len(self._sides)]
        self._next_side += 1
        return s

    def _state_for(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            o = self._offsets.get(id(p))
            if o is not None and p.is_contiguous():
                st["exp_avg"] = self._m[o:o + p.numel()].view_as(p)
                st["exp_avg_sq"] = self._v[o:o + p.numel()].view_as(p)
            else:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        return st

    def _table(self, b, wire):
        rows, key = [], []
        for p in b.params:
            st = self._state_for(p)
            o = self._offsets[id(p)]
            if wire == "bf16":
                g_ptr, g_item = self._wire.data_ptr() + 2 * o, 2
            else:
                g = p.grad
                if g is None or not g.is_contiguous() or g.dtype != torch.float32:
                    raise RuntimeError("FusedAdam: gradients must be contiguous fp32")
                g_ptr, g_item = g.data_ptr(), 4
            m, v = st["exp_avg"], st["exp_avg_sq"]
            key.append((p.data_ptr(), g_ptr, m.data_ptr(), v.data_ptr()))
            n = p.numel()
            for c in range(0, n, CHUNK):
                rows.append((p.data_ptr() + 4 * c, g_ptr + g_item * c, m.data_ptr() + 4 * c, v.data_ptr() + 4 * c, min(CHUNK, n - c)))
        key = tuple(key)
        cached = b.tables.get(wire)
        if cached is not None and cached[0] == key:
            return cached[1], cached[2]
        arr = (L.AdamChunk * max(1, len(rows)))()
        for i, r in enumerate(rows):
            arr[i].p, arr[i].g, arr[i].m, arr[i].v, arr[i].numel = r
        host = torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy())
        dev = host.to(b.params[0].device)
        b.tables[wire] = (key, dev, len(rows))
        return dev, len(rows)

    def _process(self, b):
        """The tail of one bucket: gradient accumulators -> .grad, (all-reduce), Adam, filter re-pack."""
        from . import plan
        if not b.params:
            b.fired = True
            return
        dev = b.params[0].device
        if not b.params[0].is_cuda:
            raise RuntimeError("FusedAdam: CUDA parameters required (no CPU fallback)")
        if dev.index != torch.cuda.current_device():
            raise RuntimeError(f"FusedAdam: parameters live on {dev} but the current CUDA device is "
                               f"cuda:{torch.cuda.current_device()} (kernels are enqueued on the current device's streams)")
        self.flat_grad()
        cur = torch.cuda.current_stream(dev)
        side = self._stream(dev) if self.overlap else None
        if side is not None:
            if b.events:
                for ev in b.events.values():
                    side.wait_event(ev)
            else:
                side.wait_stream(cur)
            self._sides_used.add(side)
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            if b.holders:
                plan.flush_grads(b.holders)
            wire = "fp32"
            if self.sync is not None and self.sync.world > 1:
                wire = self.sync.reduce_bucket(self, b)
            group = self.param_groups[0]
            if len(self.param_groups) > 1:
                raise RuntimeError("FusedAdam: one parameter group")
            for k in ("weight_decay", "amsgrad", "maximize"):
                if group.get(k):
                    raise RuntimeError(f"FusedAdam: param_group option {k} is not supported")
            beta1, beta2 = group["betas"]
            table, nchunks = self._table(b, wire)
            flags = 0
            ticking = not self._ticked
            if ticking:
                self._sync_step_counter(dev)
                flags |= L.ADAM_TICK
                self._ticked = True
            elif side is not None and self._tick_event is not None:
                side.wait_event(self._tick_event)       # the step counter / bias corrections advance once per step
            if wire == "bf16":
                flags |= L.ADAM_GRAD_BF16
            elif not self.keep_grads:
                flags |= L.ADAM_ZERO_GRAD
            ops.adam_multi(table, nchunks, self._dev_state, group["lr"], beta1, beta2, group["eps"], self.grad_scale, flags)
            if ticking and side is not None:
                self._tick_event = torch.cuda.Event()
                self._tick_event.record(side)
            for p in b.params:        # the kernel wrote p behind autograd's back: invalidate packed copies
                p._vcg_epoch = getattr(p, "_vcg_epoch", 0) + 1
            if b.holders:
                plan.refresh_holders(b.holders)
        b.fired = True
        b.events = {}

    def _sync_step_counter(self, dev):
        steps = [self._state_for(p)["step"] for p in self._params()]
        before = int(steps[0].item())
        if any(int(s.item()) != before for s in steps[1:]):
            raise RuntimeError("FusedAdam: parameters of one optimiser must share the step count")
        if self._dev_state is None or self._dev_state.device != dev:
            self._dev_state = torch.zeros(4, dtype=torch.float32, device=dev)
            self._dev_step = 0
        if self._dev_step != before:          # first use / after load_state_dict: resync the device counter
            self._dev_state[0:1].fill_(float(before))
            self._dev_step = before

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        from . import plan
        ps = self._params()
        if not ps:
            return loss
        if not ps[0].is_cuda:
            raise RuntimeError("FusedAdam: CUDA parameters required (no CPU fallback)")
        if not self._flat_grads:
            raise RuntimeError("FusedAdam: flat_grads=False is not supported by the fused update")
        self.flat_grad()
        pending = [b for b in self._buckets if not b.fired]
        if any(b.holders is None for b in pending) or plan.has_pending():
            plan.flush_grads()          # gradients produced outside a tracked backward (on the calling stream)
        for b in pending:
            self._process(b)
        # host-side bookkeeping of one optimiser step
        steps = [self.state[p]["step"] for p in ps if len(self.state[p])]
        for s in steps:
            s += 1
        self._dev_step += 1
        self._last = steps
        self._ticked = False
        self._tick_event = None
        self._next_side = 0
        for b in self._buckets:
            b.fired = False
        self._zeroed = not self.keep_grads      # consumed gradients were zeroed by the Adam / wire-cast kernels
        return loss

    @torch.no_grad()
    def finish(self):
        """Join the side stream: afterwards the calling stream sees the updated weights and their packed copies."""
        if self._sides_used:
            cur = torch.cuda.current_stream(next(iter(self._sides_used)).device)
            for s in self._sides_used:
                cur.wait_stream(s)
            self._sides_used = set()

    def state_dict(self):
        self.finish()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        """torch's load replaces the state tensors; the loaded moments are copied back INTO the existing (flat-buffer)
        tensors so that device tables -- possibly baked into a captured CUDA graph -- keep pointing at live memory."""
        self.finish()
        old = {p: dict(self.state[p]) for p in self._params() if len(self.state[p])}
        super().load_state_dict(state_dict)
        with torch.no_grad():
            for p, o in old.items():
                new = self.state[p]
                for k in ("exp_avg", "exp_avg_sq"):
                    if k in new and k in o and o[k].shape == new[k].shape and o[k].device == new[k].device:
                        o[k].copy_(new[k])
                        new[k] = o[k]
        for b in self._buckets or []:
            b.tables = {}

    def capture_rollback(self):
        """step() ran under CUDA-graph capture: its kernels were recorded, not executed; undo the host bump."""
        if self._last is not None:
            for st in self._last:
                st -= 1
            self._dev_step -= 1

    def note_replay(self):
        """A captured CUDA graph containing this optimiser's step was replayed: the device-side counter
        and the weights advanced; mirror that in the host-side state (state_dict 'step', cache epochs)."""
        if self._last is not None:
            for st in self._last:
                st += 1
            self._dev_step += 1
            for p in self._params():
                p._vcg_epoch = getattr(p, "_vcg_epoch", 0) + 1
