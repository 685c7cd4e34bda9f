"""Micro-benchmark of the multi-tensor filter pack / gradient unpack launches at generator scale (66 M params)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vcg_b200  # noqa
from vcg_b200 import lib as L, ops

U, P = L.WMAP_UNSHUFFLE, L.WMAP_PLAIN
SPECS = [(64, 3, 7, P, 8), (128, 256, 3, U, 256), (256, 512, 3, U, 512), (512, 1024, 3, U, 1024), (1024, 2048, 3, U, 2048),
         (1024, 1024, 3, P, 1024), (1024, 1024, 3, P, 1024), (64, 1024, 3, P, 1024), (64, 1024, 3, P, 1024), (64, 64, 3, P, 64),
         (1024, 64, 3, P, 64), (1024, 1024, 3, P, 1024), (1024, 1024, 3, P, 1024), (512, 256, 3, P, 256), (256, 128, 3, P, 128),
         (128, 64, 3, P, 64), (64, 32, 3, P, 32), (3, 64, 7, P, 64)]
specs = [ops.ConvSpec(co, ci, k, k, wm, cp) for co, ci, k, wm, cp in SPECS]
ws = [torch.randn(sp.co, sp.ci, sp.kh, sp.kw, device="cuda") for sp in specs]
nparam = sum(w.numel() for w in ws)
dt = torch.bfloat16
wk = [torch.zeros(sp.packed_shape(False), dtype=dt, device="cuda") for sp in specs]
wkT = [torch.zeros(sp.packed_shape(True), dtype=dt, device="cuda") for sp in specs]
dws = [torch.randn(sp.packed_shape(False), device="cuda") for sp in specs]
grads = [torch.zeros_like(w) for w in ws]
tp = ops.wjob_table([(sp, w, a, b) for sp, w, a, b in zip(specs, ws, wk, wkT)], ws[0].device)
tu = ops.wjob_table([(sp, g, dw, None) for sp, g, dw in zip(specs, grads, dws)], ws[0].device)


def timeit(fn, iters=5):
    fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


t = timeit(lambda: ops.wpack_multi(tp, dt))
print(f"pack   {nparam / 1e6:.1f} M params: {t * 1e3:8.1f} us  {(4 + 2 + 2) * nparam / t / 1e6:7.0f} GB/s (tiles {tp[2]})")
t = timeit(lambda: ops.wunpack_multi(tu))
print(f"unpack {nparam / 1e6:.1f} M params: {t * 1e3:8.1f} us  {16 * nparam / t / 1e6:7.0f} GB/s")
tw = ops.wjob_table([(sp, g, dw, None) for sp, g, dw in zip(specs, grads, dws)], ws[0].device, accumulate=False)
t = timeit(lambda: ops.wunpack_multi(tw))
print(f"unpack (write, no accumulate) {t * 1e3:8.1f} us  {12 * nparam / t / 1e6:7.0f} GB/s")
flat = torch.empty(nparam, device="cuda")
t = timeit(lambda: ops.zero_(flat))
print(f"zero {nparam * 4 / 1e6:.0f} MB: {t * 1e3:8.1f} us  {4 * nparam / t / 1e6:7.0f} GB/s")
