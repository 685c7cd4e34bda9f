"""Data parallelism for the training step (the reference has none: single process, single device,
train.py:385).  One process per GPU; the batch is sharded along N; weights, Adam state and the
spectral-norm buffers are replicated; the only exchange per optimiser step is an all-reduce(sum) of
the optimiser's flat gradient buffer (optim.FusedAdam.flat_grad), bucket by bucket on a side stream over
NCCL/NVLink as soon as a bucket's gradients are final (overlapping the rest of the backward pass and the
discriminator step), as bf16 on the wire, with the 1/world averaging folded into the Adam kernel (grad_scale).

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
    """All-reduce of the optimisers' flat gradient buffers.

    reduce_bucket(opt, bucket) is called by optim.FusedAdam on its side stream as soon as a bucket's gradients are
    final (FusedAdam.track), i.e. while the rest of the backward pass and the discriminator step still run.
    wire='bf16' (default on CUDA): the bucket's slice of the fp32 buffer is cast into a bf16 wire buffer (one
    launch, which also re-zeroes the fp32 slice for the next step) and the bf16 slice is all-reduced (sum) -- half
    the NVLink bytes of fp32; Adam reads the reduced bf16 gradient and folds in the 1/world averaging.
    wire='fp32': the fp32 slice is all-reduced in place (parity tests).
    reduce(flat) / wait() are the plain whole-buffer form (CPU / gloo tests of the host logic)."""

    def __init__(self, group=None, bucket_bytes=64 << 20, wire="bf16"):
        if wire not in ("bf16", "fp32"):
            raise ValueError("GradSync: wire must be 'bf16' or 'fp32'")
        self.group = group
        self.bucket_bytes = bucket_bytes
        self.wire = wire
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._stream = None
        self._pending = []
        self.bytes_sent = 0           # per-rank payload handed to all_reduce so far (diagnostics)

    def buckets(self, numel, itemsize=4):
        per = max(1, self.bucket_bytes // itemsize)
        return [(o, min(per, numel - o)) for o in range(0, numel, per)]

    def reduce_bucket(self, opt, b):
        """-> 'bf16' (the reduced gradient is in opt.wire_buffer()[b.lo:b.hi]) or 'fp32' (reduced in place)."""
        flat = opt.flat_grad()
        n = b.hi - b.lo
        if self.world == 1 or n == 0:
            return "fp32"
        if self.wire == "bf16" and flat.is_cuda:
            from . import ops
            wire = opt.wire_buffer()
            ops.cast_bf16(flat[b.lo:b.hi], wire[b.lo:b.hi], zero_src=not opt.keep_grads)
            for o, m in self.buckets(n, 2):
                dist.all_reduce(wire[b.lo + o:b.lo + o + m], op=dist.ReduceOp.SUM, group=self.group)
            self.bytes_sent += 2 * n
            return "bf16"
        for o, m in self.buckets(n, 4):
            dist.all_reduce(flat[b.lo + o:b.lo + o + m], op=dist.ReduceOp.SUM, group=self.group)
        self.bytes_sent += 4 * n
        return "fp32"

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
    """Make a composite model data-parallel: every FusedAdam it owns all-reduces each gradient bucket on its side
    stream before the bucket's Adam launch and averages inside the Adam kernel; metrics are averaged too."""
    sync = sync or GradSync()
    for name in ("optimizer", "optimizer_G", "optimizer_D"):
        opt = getattr(model, name, None)
        if opt is None:
            continue
        opt.grad_scale = 1.0 / sync.world
        opt.sync = sync
    model._vcg_sync = sync
    return sync


def broadcast_state(model, src=0):
    """Replicate parameters and buffers from rank `src` (construction is seeded, this is a guard)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src)
