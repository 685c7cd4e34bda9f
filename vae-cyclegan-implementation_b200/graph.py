"""CUDA-graph replay of the whole training step.

A CycleVAEGAN step is ~1300 kernel launches; at small per-GPU batch (8 GPUs x batch 8) the Python /
launch path, not the GPU, would set the step time.  GraphedStep captures `model.training_step` once for
fixed batch shapes -- forward, both backward sweeps (autograd runs at capture time only), the NCCL
gradient all-reduce, both Adam updates (step counters live on the device, csrc/adam.cu) and the metric
reductions -- and then replays it: one host call per step, one D2H read of the packed metrics.

The warm-up steps that precede the capture (allocator warm-up, descriptor and table caches) are real training steps;
their effect is undone before the first replay: parameters, buffers (spectral-norm u / v), Adam moments and step
counters and the CUDA RNG state are snapshotted first and restored in place afterwards, so a graphed run follows the
same trajectory as the eager one and a checkpoint's `step` equals the number of batches seen."""
from __future__ import annotations

import torch


class GraphedStep:
    def __init__(self, model, example_batch, warmup=2, profile=False):
        """profile=True: every kernel launch of the captured step is bracketed by external CUDA events (ops.prof_begin)
        that each replay re-records; kernel_times() reads the last replay's durations (bench.py's roofline)."""
        self.model = model
        self.records = None
        self.x = example_batch["x"].detach().clone()
        self.y = example_batch["y"].detach().clone()
        if not self.x.is_cuda:
            raise RuntimeError("GraphedStep: CUDA tensors required")
        self.opts = [o for o in (getattr(model, n, None) for n in ("optimizer", "optimizer_G", "optimizer_D")) if o is not None]
        batch = {"x": self.x, "y": self.y}
        # two warm-up steps at least: the first one plans and packs inside the passes, the second one takes the
        # steady-state path (all filters packed by one launch before the lanes fork) and builds its device tables --
        # host-to-device table copies are not legal inside the capture
        warmup = max(2, warmup)
        snap = self._snapshot()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                      # allocations, descriptor caches, attribute calls
                model.training_step(batch)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        sink = {}
        model._vcg_capture = sink
        self.graph = torch.cuda.CUDAGraph()
        from . import lib, ops
        if profile:
            ops.prof_begin(external=True)
        n0 = lib.launch_count()
        try:
            with torch.cuda.graph(self.graph):
                model.training_step(batch)
        finally:
            model._vcg_capture = None
            if profile:
                self.records = ops.prof_end()
        self.launches_per_step = lib.launch_count() - n0      # kernels of this library inside one replay
        for o in self.opts:                              # capture recorded the step but did not execute it
            o.capture_rollback()
        self._restore(snap)
        self.keys, self.vals = sink["keys"], sink["vals"]
        # hyper-parameters are kernel ARGUMENTS baked into the captured launches
        self._hyper = [self._hyper_of(o) for o in self.opts]
        self._stage_x = self._stage_y = self._copy_stream = self._staged_ready = self._stage_free = None
        self._staged_key = None

    @staticmethod
    def _hyper_of(opt):
        return [(g["lr"], tuple(g["betas"]), g["eps"]) for g in opt.param_groups]

    def kernel_times(self):
        """[(family, flops or bytes, tag, milliseconds)] of the kernel launches of the LAST replay (profile=True)."""
        if self.records is None:
            raise RuntimeError("GraphedStep was built without profile=True")
        torch.cuda.synchronize()
        return [(kind, work, tag, s.elapsed_time(e)) for kind, work, tag, s, e in self.records]

    # ---- undo the warm-up steps (in place: every pointer baked into the graph stays valid)
    def _snapshot(self):
        snap = {"tensors": [(t, t.detach().clone()) for t in list(self.model.parameters()) + list(self.model.buffers())],
                "rng": torch.cuda.get_rng_state(self.x.device), "opts": []}
        for o in self.opts:
            o.finish()
            st = {}
            for p in o._params():
                s = o.state.get(p) or {}
                st[p] = {k: v.detach().clone() for k, v in s.items()}
            snap["opts"].append((o, st, None if o._dev_state is None else o._dev_state.clone(), o._dev_step))
        return snap

    @torch.no_grad()
    def _restore(self, snap):
        for t, saved in snap["tensors"]:
            t.copy_(saved)
        for o, st, dev_state, dev_step in snap["opts"]:
            for p in o._params():
                cur, old = o.state.get(p) or {}, st[p]
                for k, v in cur.items():
                    if k in old:
                        v.copy_(old[k])
                    else:
                        v.zero_()                       # moments / step created by the warm-up: back to "never stepped"
            if o._dev_state is not None:
                if dev_state is not None:
                    o._dev_state.copy_(dev_state)
                    o._dev_step = dev_step
                else:
                    o._dev_state.zero_()
                    o._dev_step = 0
        # the kernel-layout filter copies still hold the warm-up weights, and the captured step re-packs them only at
        # its END (behind each bucket's Adam): bring them back to the restored masters before the first replay
        prepack = getattr(self.model, "_prepack", None)
        if prepack is not None:
            prepack()
        torch.cuda.set_rng_state(snap["rng"], self.x.device)
        torch.cuda.synchronize()

    def __call__(self, batch, prefetch=None):
        """Run one step on `batch`.  `prefetch` (optional) is the NEXT step's batch in pinned host memory: its
        host-to-device copy is issued on a side stream right after the graph launch, so it overlaps this step's
        compute instead of preceding the next one (what a pinned-memory data loader with non_blocking copies does).
        The caller must not modify the prefetched host tensors before the call that consumes them."""
        if [self._hyper_of(o) for o in self.opts] != self._hyper:
            raise RuntimeError("GraphedStep: an optimiser's lr / betas / eps changed after the capture; they are baked into "
                               "the captured Adam launches -- build a new GraphedStep (e.g. once per lr-schedule step)")
        key = self._batch_key(batch)
        if self._staged_key is not None and key == self._staged_key:
            torch.cuda.current_stream().wait_event(self._staged_ready)
            self.x.copy_(self._stage_x, non_blocking=True)          # device-to-device, ~30 us
            self.y.copy_(self._stage_y, non_blocking=True)
        else:
            self.x.copy_(batch["x"], non_blocking=True)
            self.y.copy_(batch["y"], non_blocking=True)
        self._staged_key = None
        self.graph.replay()
        if prefetch is not None and not prefetch["x"].is_cuda:
            if self._stage_x is None:
                self._stage_x, self._stage_y = torch.empty_like(self.x), torch.empty_like(self.y)
                self._copy_stream = torch.cuda.Stream()
                self._staged_ready = torch.cuda.Event()
                self._stage_free = torch.cuda.Event()
            self._stage_free.record()                               # staging buffers were read by the copies above
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(self._stage_free)
                self._stage_x.copy_(prefetch["x"], non_blocking=True)
                self._stage_y.copy_(prefetch["y"], non_blocking=True)
                self._staged_ready.record()
            self._staged_key = self._batch_key(prefetch)
        for o in self.opts:
            o.note_replay()
        vals = self.vals.tolist()
        if not all(v == v and abs(v) != float("inf") for v in vals):
            # the reference's NaN / Inf guard (Networks.py:356-372) is a host read before the update and cannot run
            # inside a replayed graph: report after the fact
            import warnings
            warnings.warn("GraphedStep: non-finite metric after a replayed step (the update was already applied)")
            out = dict(zip(self.keys, vals))
            out["nan_detected"] = True
            return out
        return dict(zip(self.keys, vals))

    @staticmethod
    def _batch_key(batch):
        return (batch["x"].data_ptr(), batch["y"].data_ptr(), tuple(batch["x"].shape))
