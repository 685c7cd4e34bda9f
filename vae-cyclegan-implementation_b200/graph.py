"""CUDA-graph replay of the whole training step.

A CycleVAEGAN step is ~1300 kernel launches; at small per-GPU batch (8 GPUs x batch 8) the Python /
launch path, not the GPU, would set the step time.  GraphedStep captures `model.training_step` once for
fixed batch shapes -- forward, both backward sweeps (autograd runs at capture time only), the NCCL
gradient all-reduce, both Adam updates (step counters live on the device, csrc/adam.cu) and the metric
reductions -- and then replays it: one host call per step, one D2H read of the packed metrics."""
from __future__ import annotations

import torch


class GraphedStep:
    def __init__(self, model, example_batch, warmup=2):
        self.model = model
        self.x = example_batch["x"].detach().clone()
        self.y = example_batch["y"].detach().clone()
        if not self.x.is_cuda:
            raise RuntimeError("GraphedStep: CUDA tensors required")
        self.opts = [o for o in (getattr(model, n, None) for n in ("optimizer", "optimizer_G", "optimizer_D")) if o is not None]
        batch = {"x": self.x, "y": self.y}
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
        try:
            with torch.cuda.graph(self.graph):
                model.training_step(batch)
        finally:
            model._vcg_capture = None
        for o in self.opts:                              # capture recorded the step but did not execute it
            o.capture_rollback()
        self.keys, self.vals = sink["keys"], sink["vals"]
        self._stage_x = self._stage_y = self._copy_stream = self._staged_ready = self._stage_free = None
        self._staged_key = None

    def __call__(self, batch, prefetch=None):
        """Run one step on `batch`.  `prefetch` (optional) is the NEXT step's batch in pinned host memory: its
        host-to-device copy is issued on a side stream right after the graph launch, so it overlaps this step's
        compute instead of preceding the next one (what a pinned-memory data loader with non_blocking copies does).
        The caller must not modify the prefetched host tensors before the call that consumes them."""
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
        return dict(zip(self.keys, self.vals.tolist()))

    @staticmethod
    def _batch_key(batch):
        return (batch["x"].data_ptr(), batch["y"].data_ptr(), tuple(batch["x"].shape))
