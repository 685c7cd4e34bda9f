"""Data parallelism for the training step (the reference has none: single process, single device,
train.py:385).  One process per GPU; the batch is sharded along N; weights, Adam state and the
spectral-norm buffers are replicated; the only exchange per optimiser step is an all-reduce(sum) of
the optimiser's flat fp32 gradient buffer (optim.FusedAdam.flat_grad), issued in buckets on a side
stream over NCCL/NVLink, with the 1/world averaging folded into the Adam kernel (grad_scale).

InstanceNorm is per-sample and every loss is a mean over batch x features, so the mean over ranks
of the per-shard gradients equals the reference's global-batch gradient (verified A2)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def shard(t, rank, world):
    """rows [r*B/W, (r+1)*B/W) of a globally ordered batch tensor (SURVEY.md 8e partitioning)."""
    b = t.shape[0]
    if b % world:
        raise ValueError(f"global batch {b} is not divisible by world size {world}")
    per = b // world
    return t[rank * per:(rank + 1) * per]


class GradSync:
    """Bucketed all-reduce of flat gradient buffers.

    reduce(flat) splits the buffer into <= bucket_bytes slices and all-reduces each (sum).  On CUDA the
    collectives run on a side stream so that they overlap whatever the main stream does next (the
    discriminator step does not depend on the generators' reduced gradients); wait() joins."""

    def __init__(self, group=None, bucket_bytes=64 << 20):
        self.group = group
        self.bucket_bytes = bucket_bytes
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._stream = None
        self._pending = []

    def buckets(self, numel, itemsize=4):
        per = max(1, self.bucket_bytes // itemsize)
        return [(o, min(per, numel - o)) for o in range(0, numel, per)]

    def reduce(self, flat):
        if self.world == 1:
            return
        if flat.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=flat.device)
            self._stream.wait_stream(torch.cuda.current_stream(flat.device))
            with torch.cuda.stream(self._stream):
                for o, n in self.buckets(flat.numel(), flat.element_size()):
                    self._pending.append(dist.all_reduce(flat[o:o + n], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            for o, n in self.buckets(flat.numel(), flat.element_size()):
                dist.all_reduce(flat[o:o + n], op=dist.ReduceOp.SUM, group=self.group)

    def wait(self, device=None):
        for w in self._pending:
            w.wait()
        self._pending.clear()
        if self._stream is not None:
            torch.cuda.current_stream(device).wait_stream(self._stream)


def attach(model, sync=None):
    """Make a composite model data-parallel: every FusedAdam it owns all-reduces its flat gradient
    buffer right before the update and averages inside the Adam kernel; metrics are averaged too."""
    sync = sync or GradSync()
    for name in ("optimizer", "optimizer_G", "optimizer_D"):
        opt = getattr(model, name, None)
        if opt is None:
            continue
        opt.grad_scale = 1.0 / sync.world
        opt.pre_step_hook = lambda o, _s=sync: _s.reduce(o.flat_grad())            # async, side stream
        opt.pre_update_hook = lambda o, _s=sync: _s.wait(o.flat_grad().device)     # joined right before the Adam kernel
        # two-optimiser models: the generators' exchange (530 MB) overlaps the discriminator backward; the
        # composite's training_step calls finish() on its optimisers at the end
        opt.defer = name == "optimizer_G"
    model._vcg_sync = sync
    return sync


def broadcast_state(model, src=0):
    """Replicate parameters and buffers from rank `src` (construction is seeded, this is a guard)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src)
