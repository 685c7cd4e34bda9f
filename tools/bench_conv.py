"""Micro-benchmark of the tcgen05 convolution kernels on the dominant layer shapes.
Prints achieved TFLOP/s per kernel (CUDA events, inputs > L2 or L2 flushed between iterations)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vcg_b200  # noqa
from vcg_b200 import lib as L, ops

SHAPES = [  # name, n, H(out), W(out), cin_phys, cout, k
    ("eD1", 16, 128, 128, 256, 128, 3), ("eD2", 16, 64, 64, 512, 256, 3), ("eD3", 16, 32, 32, 1024, 512, 3),
    ("eD4", 16, 16, 16, 2048, 1024, 3), ("eR", 16, 16, 16, 1024, 1024, 3), ("dU1", 16, 32, 32, 256, 512, 3),
    ("dU2", 16, 64, 64, 128, 256, 3), ("dU3", 16, 128, 128, 64, 128, 3), ("eR_b64", 64, 16, 16, 1024, 1024, 3),
    # thin layers at the batch-64 step (cin_phys = 8 for the 3-channel image)
    ("e0_b64", 64, 256, 256, 8, 64, 7), ("d5_b64", 64, 256, 256, 64, 3, 7), ("dU4_b64", 64, 256, 256, 32, 64, 3),
    ("dU3_b64", 64, 128, 128, 64, 128, 3), ("eD1_b64", 64, 128, 128, 256, 128, 3),
    ("dU2_b64", 64, 64, 64, 128, 256, 3), ("dU1_b64", 64, 32, 32, 256, 512, 3),
    # probes: same output size as e0, different K
    ("p_k3", 64, 256, 256, 8, 64, 3), ("p_k5", 64, 256, 256, 8, 64, 5), ("p_c64k1", 64, 256, 256, 64, 64, 1), ("p_c64k3", 64, 256, 256, 64, 64, 3),
]


def timeit(fn, iters=5, flush=None):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    only = sys.argv[1:] or None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = []
    for name, n, H, W, c, co, k in SHAPES:
        if only and name not in only:
            continue
        spec = ops.ConvSpec(co, 3 if c == 8 else c, k, k, L.WMAP_PLAIN, c)
        dt = torch.bfloat16
        xp = torch.randn(n, H + k - 1, W + k - 1, c, device="cuda").to(dt)
        wk = (torch.randn(spec.packed_shape(False), device="cuda") * 0.02).to(dt)
        wkT = (torch.randn(spec.packed_shape(True), device="cuda") * 0.02).to(dt)
        bias = torch.zeros(co, device="cuda")
        y = torch.empty(n, H, W, spec.out_c, dtype=dt, device="cuda")
        acc = torch.zeros(n * co * 2, device="cuda")
        halo = k - 1
        dyp = torch.randn(n, H + 2 * halo, W + 2 * halo, spec.out_c, device="cuda").to(dt)
        dxp = torch.empty_like(xp)
        dw = torch.zeros(spec.packed_shape(False), device="cuda")
        flops = 2.0 * n * H * W * co * spec.ci * k * k
        use_stats = co >= 8
        t_f = timeit(lambda: ops.conv_fwd(spec, xp, wk, bias, y, acc if use_stats else None, L.ACT_RELU if use_stats else L.ACT_NONE), flush=flush)
        t_d = timeit(lambda: ops.conv_dgrad(spec, dyp, wkT, dxp), flush=flush)
        t_w = timeit(lambda: ops.conv_wgrad(spec, xp, dyp, dw), flush=flush)
        rec = {"layer": name, "n": n, "fwd_ms": round(t_f, 4), "dgrad_ms": round(t_d, 4), "wgrad_ms": round(t_w, 4),
               "fwd_tflops": round(flops / t_f / 1e9, 1), "dgrad_tflops": round(flops / t_d / 1e9, 1),
               "wgrad_tflops": round(flops / t_w / 1e9, 1)}
        print(json.dumps(rec), flush=True)
        out.append(rec)
    return out


if __name__ == "__main__":
    main()
