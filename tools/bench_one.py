"""Run ONE conv layer shape a few times (for ncu --set full captures).
usage: python tools/bench_one.py <name> [fwd|dgrad|wgrad] [iters]"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vcg_b200  # noqa
from vcg_b200 import lib as L, ops

SH = {  # name: n, H, W, cin_phys, cout, k, ci_ref
    "eR": (64, 16, 16, 1024, 1024, 3, 1024), "eD1": (64, 128, 128, 256, 128, 3, 256), "e0": (64, 256, 256, 8, 64, 7, 3),
    "d5": (64, 256, 256, 64, 3, 7, 64), "dU4": (64, 256, 256, 32, 64, 3, 32), "dU3": (64, 128, 128, 64, 128, 3, 64),
}
name = sys.argv[1]
what = sys.argv[2] if len(sys.argv) > 2 else "fwd"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n, H, W, c, co, k, ci = SH[name]
spec = ops.ConvSpec(co, ci, k, k, L.WMAP_PLAIN, c)
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
for _ in range(iters):
    if what == "fwd":
        ops.conv_fwd(spec, xp, wk, bias, y, acc if co >= 8 else None, L.ACT_RELU if co >= 8 else L.ACT_NONE)
    elif what == "dgrad":
        ops.conv_dgrad(spec, dyp, wkT, dxp)
    else:
        ops.conv_wgrad(spec, xp, dyp, dw)
torch.cuda.synchronize()
print("ok", name, what)
