"""Micro-benchmark of the memory-bound transform kernels at the shapes of the batch-64 step.
usage: python tools/bench_xform.py [n] [prefetch distances, e.g. 2,0,4]   -> one line per case: us/launch and GB/s (algorithmic bytes)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vcg_b200  # noqa
from vcg_b200 import lib as L, ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
PF = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1]
dt = torch.bfloat16
FWD = [  # mode, pad, c, h, w, norm, act, res
    (2, 1, 64, 256, 256, 1, 0, 0), (0, 3, 64, 256, 256, 1, 0, 0), (1, 1, 128, 128, 128, 1, 0, 0), (2, 1, 128, 128, 128, 1, 0, 0),
    (1, 1, 256, 64, 64, 1, 0, 0), (2, 1, 256, 64, 64, 1, 0, 0), (0, 1, 1024, 16, 16, 1, 0, 0), (0, 1, 1024, 16, 16, 1, 0, 1),
    (1, 1, 1024, 16, 16, 1, 0, 1), (3, 1, 64, 128, 128, 0, 0, 0), (3, 1, 8, 256, 256, 0, 0, 0), (0, 3, 8, 256, 256, 0, 0, 0),
]
BWD = [  # c, h, w, [(mode, pad)], norm, act
    (64, 256, 256, [(2, 1)], 1, 0), (64, 256, 256, [(0, 3)], 1, 0), (128, 128, 128, [(1, 1)], 1, 0), (128, 128, 128, [(2, 1)], 1, 0),
    (256, 64, 64, [(1, 1)], 1, 0), (256, 64, 64, [(2, 1)], 1, 0), (512, 32, 32, [(1, 1)], 1, 0), (1024, 16, 16, [(0, 1)], 1, 0),
    (1024, 16, 16, [(0, 1), (0, 1)], 1, 0), (64, 128, 128, [(3, 1)], 0, 2), (8, 256, 256, [(0, 3)], 0, 0),
]


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


tot = 0.0
for mode, pad, c, h, w, norm, act, res in FWD:
    src = torch.randn(n, h, w, c, device="cuda").to(dt)
    mr = torch.rand(n * c * 2, device="cuda") + 0.5 if norm else None
    dst = torch.empty(ops.xform_dst_shape(n, h, w, c, mode, pad), dtype=dt, device="cuda")
    resbuf = torch.randn(n, h + 2, w + 2, c, device="cuda").to(dt) if res else None
    ms = timeit(lambda: ops.xform_fwd(src, c, dst, mode, pad, mr, act, resbuf, 1 if res else 0))
    by = src.numel() * 2 + dst.numel() * 2 + (src.numel() * 2 if res else 0)
    tot += ms
    print(f"fwd  mode{mode} pad{pad} c{c:5d} {h:3d}x{w:3d} norm{norm} res{res}: {ms * 1e3:8.1f} us  {by / ms / 1e6:8.0f} GB/s")
for c, h, w, srcs, norm, act in BWD:
    y = torch.randn(n, h, w, c, device="cuda").to(dt)
    mr = torch.rand(n * c * 2, device="cuda") + 0.5 if norm else None
    sl = [(torch.randn(ops.xform_dst_shape(n, h, w, c, m, p), device="cuda").to(dt), m, p, True) for m, p in srcs]
    ms = timeit(lambda: [ops.fold_halo_(t, m, p, h, w, c) for t, m, p, _ in sl])
    tot += ms
    print(f"fold modes{[m for m, _ in srcs]} c{c:5d} {h:3d}x{w:3d}      : {ms * 1e3:8.1f} us")
    dy = torch.zeros(n, h + 2, w + 2, c, dtype=dt, device="cuda")
    gs = torch.zeros(n * c * 2, device="cuda") if norm else None
    db = torch.zeros(c, device="cuda")
    for pf in PF:        # L2 read-ahead distance (vcg_set_l2_prefetch); the first entry is the library default
        L.check(L.load().vcg_set_l2_prefetch(pf), "vcg_set_l2_prefetch")
        ms = timeit(lambda: ops.xform_bwd_gather(sl, y, n, h, w, c, dy, 1, mr, act, 0, gs, None if norm else db))
        by = sum(t[0].numel() * 2 for t in sl) + 2 * y.numel() * 2
        tot += ms if pf == PF[0] else 0.0
        print(f"gath modes{[m for m, _ in srcs]} c{c:5d} {h:3d}x{w:3d} norm{norm} pf{pf}: {ms * 1e3:8.1f} us  {by / ms / 1e6:8.0f} GB/s")
        if norm:
            ms = timeit(lambda: ops.xform_bwd_norm(y, n, h, w, c, dy, 1, mr, gs, 0, db))
            by = 3 * y.numel() * 2
            tot += ms if pf == PF[0] else 0.0
            print(f"norm            c{c:5d} {h:3d}x{w:3d}       pf{pf}: {ms * 1e3:8.1f} us  {by / ms / 1e6:8.0f} GB/s")
    L.check(L.load().vcg_set_l2_prefetch(PF[0]), "vcg_set_l2_prefetch")
print(f"total {tot:.3f} ms")
