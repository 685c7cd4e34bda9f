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

    def __call__(self, batch):
        self.x.copy_(batch["x"], non_blocking=True)
        self.y.copy_(batch["y"], non_blocking=True)
        self.graph.replay()
        for o in self.opts:
            o.note_replay()
        return dict(zip(self.keys, self.vals.tolist()))
