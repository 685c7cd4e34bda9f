"""Data-parallel host logic on CPU: world_size-2 gloo (no GPU)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import vcg_b200  # noqa: F401
    from vcg_b200 import dist as vdist
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, _, w = vdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    g = torch.Generator().manual_seed(7)
    x = torch.rand(8, 3, 4, 4, generator=g)
    mine = vdist.shard(x, rank, world)
    assert torch.equal(mine, x[rank * 4:(rank + 1) * 4])
    # bucketed all-reduce of a ragged flat gradient buffer (3 buckets, last one short)
    sync = vdist.GradSync(bucket_bytes=4096)
    flat = torch.arange(2500, dtype=torch.float32) * (rank + 1)
    assert sync.buckets(flat.numel()) == [(0, 1024), (1024, 1024), (2048, 452)]
    sync.reduce(flat)
    sync.wait()
    ok = torch.equal(flat, torch.arange(2500, dtype=torch.float32) * 3)
    # the per-bucket form optim.FusedAdam calls on its side stream: only [lo, hi) of the flat buffer is exchanged
    # (fp32 wire on CPU tensors; the bf16 wire needs the CUDA cast kernel and is covered by tools/dp_check.py)
    class _Opt:
        keep_grads = True

        def __init__(self, t):
            self.t = t

        def flat_grad(self):
            return self.t

    class _Bucket:
        lo, hi = 100, 2300

    flat2 = torch.arange(2500, dtype=torch.float32) * (rank + 1)
    sync2 = vdist.GradSync(bucket_bytes=4096, wire="bf16")          # falls back to fp32 for non-CUDA buffers
    assert sync2.reduce_bucket(_Opt(flat2), _Bucket()) == "fp32"
    expect = torch.arange(2500, dtype=torch.float32) * (rank + 1)
    expect[100:2300] = torch.arange(100, 2300, dtype=torch.float32) * 3
    ok = ok and torch.equal(flat2, expect) and sync2.bytes_sent == 4 * 2200
    # mean over ranks of shard gradients == global-batch gradient for a mean loss
    w_ = torch.ones(3, requires_grad=True)
    (mine.mean(dim=(0, 2, 3)) * w_).sum().backward()
    gsum = w_.grad.clone()
    dist.all_reduce(gsum)
    w2 = torch.ones(3, requires_grad=True)
    (x.mean(dim=(0, 2, 3)) * w2).sum().backward()
    ok = ok and torch.allclose(gsum / world, w2.grad, atol=1e-7)
    with pytest.raises(ValueError):
        vdist.shard(torch.zeros(3, 1), rank, world)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gloo_world2_gradient_exchange():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
