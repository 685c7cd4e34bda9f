"""Per-kernel parity tests (-m gpu): every exported kernel family against a plain torch fp32/fp64
reference of the same op on identical seeded inputs.  Tolerances: 1e-5 rel-L2 in fp32 mode, 2e-2 in
bf16 mode (north_star); most bf16 checks are far tighter because inputs are pre-rounded to bf16."""
import pytest
import torch
import torch.nn.functional as F

from util import assert_close, nchw, nhwc, ref_unshuffle_phys, ref_xform, rel_l2

pytestmark = pytest.mark.gpu

DT = {"fp32": torch.float32, "bf16": torch.bfloat16}
TOL = {"fp32": 1e-5, "bf16": 2e-2}


@pytest.fixture(scope="module")
def K(vcg):
    from vcg_b200 import lib, ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return ops, lib


def dev(t):
    return t.cuda()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def q(t, dtype):
    """round to the storage dtype and back so the reference sees exactly what the kernel sees"""
    return t.to(dtype).float()


# --------------------------------------------------------------------------------------------
def test_tmap_window_probe(K):
    """Overlapping-stride ('window') TMA descriptors must be legal: dim0 = kw*c elements, dim1 stride = c."""
    ops, _ = K
    buf = torch.zeros(1 << 20, dtype=torch.bfloat16, device="cuda")
    c, kw, wp, hp, n, wo = 8, 7, 38, 38, 1, 32
    rc = ops.probe_tmap(buf, [kw * c, wo, hp, n], [c * 2, wp * c * 2, hp * wp * c * 2], [64, 32, 4, 1])
    assert rc == 0, "window tensor map rejected by cuTensorMapEncodeTiled"
    rc = ops.probe_tmap(buf, [64, 18, 18, 2], [128, 18 * 128, 18 * 18 * 128], [64, 16, 8, 1])
    assert rc == 0


# --------------------------------------------------------------------------------------------
XF_CASES = [  # (mode, pad, c, h, w, norm, act, res)
    (0, 1, 64, 16, 16, True, 0, True), (0, 1, 64, 16, 16, False, 0, False), (0, 3, 8, 32, 32, True, 1, False),
    (0, 0, 16, 16, 16, True, 2, False), (1, 1, 128, 16, 16, True, 0, False), (2, 1, 32, 32, 32, True, 0, False),
    (3, 1, 16, 32, 32, True, 2, False), (3, 1, 8, 32, 32, False, 0, False), (1, 1, 64, 8, 24, False, 1, False),
    (0, 3, 64, 64, 48, True, 0, False), (2, 1, 64, 64, 64, True, 0, False), (1, 1, 128, 24, 40, True, 0, True),
    (3, 1, 64, 32, 64, False, 2, False), (0, 1, 512, 16, 16, True, 0, True), (2, 1, 24, 16, 16, True, 0, False),
    (0, 2, 8, 6, 5, False, 0, False),
    # CaSb's Tanh / Sigmoid after the norm (generic gather kernel: derivative from the recomputed pre-activation)
    (0, 1, 64, 16, 16, True, 3, False), (0, 3, 16, 32, 32, True, 4, False),
]


@pytest.mark.parametrize("folded", [False, True], ids=["mirrors", "folded"])
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("case", XF_CASES)
def test_xform_fwd_bwd(K, prec, case, folded):
    ops, L = K
    mode, pad, c, h, w, norm, act, res = case
    dt, n = DT[prec], 2
    x = q(rnd(n, c, h, w, seed=1), dt)
    r = q(rnd(n, c, h, w, seed=2), dt) if res else None
    xr = x.clone().requires_grad_(True)
    rr = r.clone().requires_grad_(True) if res else None
    ref = ref_xform(xr.double(), mode, pad, norm, act, rr.double() if res else None)
    # ours
    src = torch.empty(n, h, w, c, dtype=dt, device="cuda")
    ops.pack_nchw(x, src)
    mr, shw = None, 0
    if norm:
        mr = torch.zeros(n * c * 6, dtype=torch.float32, device="cuda")
        ops.in_stats(src, c, mr)
        if folded:      # production bf16 path: RAW {sum, sum of squares} pairs, mean / rstd derived on load (stats_hw)
            xs = nhwc(x).reshape(n, h * w, c).double()
            mr = torch.stack([xs.sum(1), (xs * xs).sum(1)], dim=-1).float().reshape(-1).contiguous()
            shw = h * w
    resbuf = None
    if res:
        resbuf = torch.empty(n, h + 2, w + 2, c, dtype=dt, device="cuda")
        ops.pack_nchw(r, resbuf, halo=1)
    dshape = ops.xform_dst_shape(n, h, w, c, mode, pad)
    dst = torch.full(dshape, float("nan"), dtype=dt, device="cuda")
    ops.xform_fwd(src, c, dst, mode, pad, mr, act, resbuf, 1 if res else 0, stats_hw=shw)
    got = nchw(dst.float())[:, :ref.shape[1]]
    assert_close(got, ref.detach(), 1e-5 if prec == "fp32" else 6e-3, f"xform_fwd {case}", "nchw")
    # backward: gradient of sum(ref * G) w.r.t. x
    G = q(rnd(*ref.shape, seed=3), dt)
    (ref * G.double()).sum().backward()
    dxp = torch.zeros(dshape, dtype=dt, device="cuda")
    ops.pack_nchw(G, dxp)   # G is already in the destination domain
    # folded variant: the gather clears the halo ring itself (clear_halo), so start from NaN
    dy = torch.full((n, h + 2, w + 2, c), float("nan") if folded else 0.0, dtype=dt, device="cuda")
    gs = torch.zeros(n * c * 2, dtype=torch.float32, device="cuda") if norm else None
    if folded:      # adjoint of the reflect pad applied once, in place; the gather then takes its fast path
        ops.fold_halo_(dxp, mode, pad, h, w, c)
    ops.xform_bwd_gather([(dxp, mode, pad, folded)], src, n, h, w, c, dy, 1, mr, act, 0, gs, None, stats_hw=shw,
                         clear_halo=folded)
    if norm:
        ops.xform_bwd_norm(src, n, h, w, c, dy, 1, mr, gs, stats_hw=shw)
    got_dx = nchw(dy[:, 1:-1, 1:-1].float())
    assert_close(got_dx, xr.grad, 2e-5 if prec == "fp32" else 1.5e-2, f"xform_bwd {case}", "nchw")
    assert float(dy[:, 0].abs().max()) == 0.0 and float(dy[:, :, -1].abs().max()) == 0.0, "halo must stay zero"


def test_l2_read_ahead_changes_no_result(K):
    """vcg_set_l2_prefetch only moves HBM -> L2 traffic earlier: the backward transforms return the same bits for every
    read-ahead distance (blocks of several iterations, chunks that wrap image rows, the last partial chunk)."""
    ops, L = K
    n, c, h, w, dt = 3, 64, 72, 52, torch.bfloat16
    y = rnd(n, h, w, c, seed=1).to(dt)
    dxp = rnd(n, h + 2, w + 2, c, seed=2).to(dt)
    mr = torch.stack([y.float().sum((1, 2)), (y.float() ** 2).sum((1, 2))], dim=-1).reshape(-1).contiguous()
    outs = []
    try:
        for dist in (0, 1, 4, 16):
            L.check(L.load().vcg_set_l2_prefetch(dist), "vcg_set_l2_prefetch")
            dy = torch.full((n, h + 2, w + 2, c), float("nan"), dtype=dt, device="cuda")
            gs = torch.zeros(n * c * 2, dtype=torch.float32, device="cuda")
            ops.xform_bwd_gather([(dxp, 0, 1, True)], y, n, h, w, c, dy, 1, mr, 0, 0, gs, None, stats_hw=h * w, clear_halo=True)
            ops.xform_bwd_norm(y, n, h, w, c, dy, 1, mr, gs, stats_hw=h * w)
            torch.cuda.synchronize()
            outs.append(dy.clone())
        assert L.load().vcg_set_l2_prefetch(17) != 0 and L.load().vcg_set_l2_prefetch(-1) != 0
    finally:
        L.check(L.load().vcg_set_l2_prefetch(1), "vcg_set_l2_prefetch")
    assert torch.isfinite(outs[0].float()).all()
    for o in outs[1:]:
        # (the per-channel sums are accumulated with atomics in a launch-dependent order: allow their last bit)
        assert float((o.float() - outs[0].float()).abs().max()) <= 2e-2 * float(outs[0].float().abs().max())
        assert (o == outs[0]).float().mean() > 0.99


# --------------------------------------------------------------------------------------------
CONV_CASES = [  # (name, n, H, W, ci_in, co, k, wmap)
    ("p64", 2, 16, 16, 64, 64, 3, 0), ("p128x256", 1, 32, 32, 128, 256, 3, 0), ("p64x128w64", 3, 64, 64, 64, 128, 3, 0),
    ("p64w128", 1, 128, 128, 64, 64, 3, 0), ("unshuf", 2, 32, 32, 32, 64, 3, 1), ("s2d", 2, 32, 32, 64, 128, 4, 2),
    ("s2d_img", 2, 64, 64, 3, 64, 4, 2), ("e0", 1, 32, 32, 3, 64, 7, 0), ("d5", 1, 32, 32, 64, 3, 7, 0),
    ("dU4", 1, 32, 32, 32, 64, 3, 0), ("bn256", 74, 16, 16, 64, 256, 3, 0), ("longK", 1, 16, 16, 1024, 64, 3, 0),
    ("lat", 2, 16, 16, 64, 1024, 3, 0),
    # wide maps: thin-output forward / data-gradient shapes served by conv_tc_fold.cu (taps folded into N)
    ("d5wide", 2, 24, 160, 64, 3, 7, 0), ("e0wide", 1, 20, 140, 3, 64, 7, 0), ("dU4wide", 1, 40, 136, 32, 64, 3, 0),
    ("d5full", 1, 256, 256, 64, 3, 7, 0),
    # odd number of 128-pixel tiles (CTA-pair kernel: the second CTA of the last pair has no tile)
    ("oddtiles", 1, 24, 16, 128, 128, 3, 0),
    # 80 pair tiles on 74 CTA pairs: the 6 left-over BN = 256 tiles are split into BN = 128 halves
    ("tailsplit", 80, 16, 16, 64, 256, 3, 0),
    # data gradient as interior tiles + border-ring tiles that batch images (partial image groups, odd group counts)
    ("ring16", 9, 16, 16, 128, 128, 3, 0), ("ring16wide", 23, 16, 16, 256, 256, 3, 0), ("ring32", 5, 32, 32, 128, 256, 3, 0), ("ring64", 3, 64, 64, 128, 128, 3, 0),
    # 256-channel-multiple outputs: CTA-pair weight-gradient kernel (K split over pixel blocks)
    ("wgpair", 2, 32, 32, 256, 512, 3, 0),
    # data gradient with 64 output channels on the CTA-pair kernel (flat tiling at 128 x 128)
    ("n64flat", 1, 128, 128, 64, 128, 3, 0),
]


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_layer_fwd_bwd(K, prec, case):
    """pack -> xform -> conv fwd (+stats) / dgrad / wgrad -> unpack against F.conv2d + autograd."""
    ops, L = K
    name, n, H, W, ci_in, co, k, wmap = case
    dt = DT[prec]
    pad = 3 if k == 7 else 1
    x = q(rnd(n, ci_in, H, W, seed=11), dt).requires_grad_(True)
    ci_ref = ci_in * 4 if wmap == 1 else ci_in
    wt = q(rnd(co, ci_ref, k, k, seed=12, scale=(ci_ref * k * k) ** -0.5), dt).requires_grad_(True)
    b = rnd(co, seed=13)
    # reference in fp64
    xd, wd = x.double(), wt.double()
    if wmap == 0:
        ref = F.conv2d(F.pad(xd, (pad,) * 4, mode="reflect"), wd, b.double())
        mode = 0
    elif wmap == 1:
        ref = F.conv2d(F.pad(F.pixel_unshuffle(xd, 2), (1,) * 4, mode="reflect"), wd, b.double())
        mode = 2
    else:
        ref = F.conv2d(F.pad(xd, (1,) * 4, mode="reflect"), wd, b.double(), stride=2)
        mode = 3
    ref = F.relu(ref)
    G = q(rnd(*ref.shape, seed=14), dt)
    (ref * G.double()).sum().backward()

    c_img = ops.rup(ci_in, 8)
    spec = ops.ConvSpec(co, ci_ref, k, k, wmap, c_phys=(4 * c_img if wmap == 2 else 0))
    src = torch.empty(n, H, W, c_img, dtype=dt, device="cuda")
    ops.pack_nchw(x.detach(), src)
    xp = torch.empty(ops.xform_dst_shape(n, H, W, c_img, mode, pad), dtype=dt, device="cuda")
    assert xp.shape[-1] == spec.cin_phys, (xp.shape, spec.cin_phys)
    ops.xform_fwd(src, c_img, xp, mode, pad)
    wk = torch.empty(spec.packed_shape(False), dtype=dt, device="cuda")
    wkT = torch.empty(spec.packed_shape(True), dtype=dt, device="cuda")
    ops.wpack(spec, wt.detach().contiguous(), wk, False)
    ops.wpack(spec, wt.detach().contiguous(), wkT, True)
    Ho, Wo = ref.shape[2], ref.shape[3]
    y = torch.full((n, Ho, Wo, spec.out_c), float("nan"), dtype=dt, device="cuda")
    use_stats = prec == "bf16"
    acc = torch.zeros(n * co * 2, dtype=torch.float32, device="cuda") if use_stats else None
    ops.conv_fwd(spec, xp, wk, b, y, acc, L.ACT_RELU)
    got = nchw(y.float())[:, :co]
    assert_close(got, ref.detach(), 1e-5 if prec == "fp32" else 5e-3, f"conv_fwd {name}", "nchw")
    if spec.out_c > co:
        assert float(y[..., co:].float().abs().max()) == 0.0, "pad output channels must be zero"
    if use_stats:
        mr = torch.empty(n * co * 2, dtype=torch.float32, device="cuda")
        ops.in_finalize(acc, n * co, Ho * Wo, mr)
        mr = mr.view(n, co, 2)
        rd = ref.detach()
        assert_close(mr[..., 0], rd.mean(dim=(2, 3)), 2e-3, f"stats mean {name}", "nc")
        assert_close(mr[..., 1], (rd.var(dim=(2, 3), unbiased=False) + 1e-5).rsqrt(), 2e-3, f"stats rstd {name}", "nc")
    # ---- backward
    halo = spec.pkh - 1
    gy = G * (ref.detach() > 0).float()           # ReLU backward happens in the xform pass in production
    dyp = torch.zeros(n, Ho + 2 * halo, Wo + 2 * halo, spec.out_c, dtype=dt, device="cuda")
    ops.pack_nchw(q(gy, dt).contiguous(), dyp, halo=halo)
    dxp = torch.full(tuple(xp.shape), float("nan"), dtype=dt, device="cuda")
    ops.conv_dgrad(spec, dyp, wkT, dxp)
    dx = torch.zeros(n, H, W, c_img, dtype=dt, device="cuda")
    ops.xform_bwd_gather([(dxp, mode, pad)], None, n, H, W, c_img, dx, 0)
    got_dx = nchw(dx.float())[:, :ci_in]
    # reference gradient uses the bf16-rounded gy too
    x2 = x.detach().double().requires_grad_(True)
    w2 = wt.detach().double().requires_grad_(True)
    if wmap == 0:
        r2 = F.conv2d(F.pad(x2, (pad,) * 4, mode="reflect"), w2)
    elif wmap == 1:
        r2 = F.conv2d(F.pad(F.pixel_unshuffle(x2, 2), (1,) * 4, mode="reflect"), w2)
    else:
        r2 = F.conv2d(F.pad(x2, (1,) * 4, mode="reflect"), w2, stride=2)
    (r2 * q(gy, dt).double()).sum().backward()
    assert_close(got_dx, x2.grad, 1e-5 if prec == "fp32" else 6e-3, f"conv_dgrad {name}", "nchw")
    dw = torch.zeros(spec.packed_shape(False), dtype=torch.float32, device="cuda")
    ops.conv_wgrad(spec, xp, dyp, dw, allow_simt=True)      # odd-width test maps have no tensor-core wgrad kernel
    gw = torch.empty(co, ci_ref, k, k, dtype=torch.float32, device="cuda")
    ops.wunpack_grad(spec, dw, gw)
    assert_close(gw, w2.grad, 1e-5 if prec == "fp32" else 5e-3, f"conv_wgrad {name}", "oihw")
    # accumulate semantics
    ops.conv_wgrad(spec, xp, dyp, dw, allow_simt=True)
    ops.wunpack_grad(spec, dw, gw)
    assert_close(gw, 2 * w2.grad, 1e-5 if prec == "fp32" else 5e-3, f"conv_wgrad accumulate {name}", "oihw")


# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_wpack_multi(K, prec):
    """multi-tensor filter pack (both layouts) / gradient unpack against the per-filter kernels"""
    ops, L = K
    dt = DT[prec]
    specs = [ops.ConvSpec(128, 64, 3, 3), ops.ConvSpec(64, 128, 3, 3, L.WMAP_UNSHUFFLE, 128),
             ops.ConvSpec(128, 64, 4, 4, L.WMAP_S2D, 256), ops.ConvSpec(64, 3, 7, 7, L.WMAP_PLAIN, 8),
             ops.ConvSpec(3, 64, 7, 7), ops.ConvSpec(64, 3, 4, 4, L.WMAP_S2D, 32), ops.ConvSpec(40, 72, 3, 3)]
    ws = [rnd(sp.co, sp.ci, sp.kh, sp.kw, seed=40 + i) for i, sp in enumerate(specs)]
    wk = [torch.zeros(sp.packed_shape(False), dtype=dt, device="cuda") for sp in specs]
    wkT = [torch.zeros(sp.packed_shape(True), dtype=dt, device="cuda") for sp in specs]
    table = ops.wjob_table([(sp, w, a, b) for sp, w, a, b in zip(specs, ws, wk, wkT)], ws[0].device)
    ops.wpack_multi(table, dt)
    for sp, w, a, b in zip(specs, ws, wk, wkT):
        ra, rb = torch.empty_like(a), torch.empty_like(b)
        ops.wpack(sp, w, ra, False)
        ops.wpack(sp, w, rb, True)
        assert torch.equal(a, ra), f"forward layout {sp}"
        assert torch.equal(b, rb), f"data-gradient layout {sp}"
    # unpack: grad += unpack(dw); dw = 0
    dws = [rnd(*sp.packed_shape(False), seed=60 + i) for i, sp in enumerate(specs)]
    grads = [torch.ones_like(w) for w in ws]
    refs = []
    for sp, dw in zip(specs, dws):
        g = torch.ones(sp.co, sp.ci, sp.kh, sp.kw, device="cuda")
        ops.wunpack_grad(sp, dw, g, True)
        refs.append(g)
    ops.wunpack_multi(ops.wjob_table([(sp, g, dw, None) for sp, g, dw in zip(specs, grads, dws)], ws[0].device))
    for sp, g, r, dw in zip(specs, grads, refs, dws):
        assert torch.equal(g, r), f"unpack {sp}"
        # every position that maps to a real weight was re-zeroed (padding positions are never read)
        chk = torch.zeros_like(g)
        ops.wunpack_grad(sp, dw, chk, False)
        assert float(chk.abs().max()) == 0.0, f"accumulator not cleared {sp}"
    # bias-style vector flush
    src = [rnd(16, seed=80), rnd(8, seed=81)]
    dst = [torch.ones(10, device="cuda"), torch.ones(3, device="cuda")]
    exp = [dst[0] + src[0][:10], dst[1] + src[1][:3]]
    ops.vecflush_multi(ops.vecjob_table([(src[0], dst[0], 10), (src[1], dst[1], 3)], src[0].device))
    assert torch.equal(dst[0], exp[0]) and torch.equal(dst[1], exp[1])
    assert float(src[0][:10].abs().max()) == 0.0 and float(src[0][10:].abs().min()) > 0.0


# --------------------------------------------------------------------------------------------
def test_losses(K):
    ops, L = K
    a, b = rnd(4, 3, 64, 64, seed=21), rnd(4, 3, 64, 64, seed=22)
    a.view(-1)[:7] = b.view(-1)[:7]             # exact ties -> sign 0
    out = torch.zeros(4, device="cuda")
    g = torch.empty_like(a)
    ops.l1_fwd_bwd(a, b, out[0:1], g, 0.5)
    assert abs(float(out[0]) / a.numel() - float(F.l1_loss(a, b))) < 1e-6
    assert torch.equal(g, 0.5 * torch.sign(a - b))
    d = rnd(37, seed=23)
    gd = torch.empty_like(d)
    ops.mse_const_fwd_bwd(d, 1.0, out[1:2], gd, 1.0 / 37)
    assert abs(float(out[1]) / 37 - float(F.mse_loss(d, torch.ones_like(d)))) < 1e-6
    assert_close(gd, 2 * (d - 1) / 37, 1e-6, "mse grad", "n")
    mu = rnd(2, 64, 16, 16, seed=24, scale=3.0).requires_grad_(True)
    lv = rnd(2, 64, 16, 16, seed=25, scale=8.0).requires_grad_(True)
    lvc = torch.clamp(lv, -10, 10)
    ref = -0.5 * torch.mean(1 + lvc - mu.pow(2) - lvc.exp())
    ref.backward()
    gmu, glv = torch.empty_like(mu), torch.empty_like(lv)
    ops.kl_fwd_bwd(mu.detach(), lv.detach(), out[2:3], gmu, glv, 1.0 / mu.numel())
    assert abs(-0.5 * float(out[2]) / mu.numel() - float(ref)) < 1e-5 * abs(float(ref))
    assert_close(gmu, mu.grad, 1e-6, "kl gmu", "nchw")
    assert_close(glv, lv.grad, 1e-6, "kl glv", "nchw")


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_reparam(K, prec):
    ops, L = K
    dt, n, c, h, w = DT[prec], 2, 64, 16, 16
    mu = q(rnd(n, c, h, w, seed=31, scale=3.0), dt).requires_grad_(True)
    lv = q(rnd(n, c, h, w, seed=32, scale=8.0), dt).requires_grad_(True)
    eps = rnd(n, c, h, w, seed=33)
    lvc = torch.clamp(lv, -10, 10)
    z = mu + eps * torch.exp(0.5 * lvc)
    G = q(rnd(n, c, h, w, seed=34), dt)
    gmu_e, glv_e = rnd(n, c, h, w, seed=35), rnd(n, c, h, w, seed=36)
    (z * G).sum().backward(retain_graph=True)
    mu_g1, lv_g1 = mu.grad.clone(), lv.grad.clone()
    mu.grad = lv.grad = None
    ((z * G).sum() + (mu * gmu_e).sum() + (lvc * glv_e).sum()).backward()
    # ours: mu / logvar are fp32 NHWC conv outputs in both modes; mu sits in a wider buffer (pitch 2c)
    both = torch.zeros(n, h, w, 2 * c, dtype=torch.float32, device="cuda")
    tmp = torch.empty(n, h, w, c, dtype=torch.float32, device="cuda")
    ops.pack_nchw(mu.detach(), tmp)
    both[..., :c] = tmp
    lvb = torch.empty(n, h, w, c, dtype=torch.float32, device="cuda")
    ops.pack_nchw(lv.detach(), lvb)
    zb = torch.empty(n, h, w, c, dtype=dt, device="cuda")
    mu_o, lv_o = torch.empty(n, c, h, w, device="cuda"), torch.empty(n, c, h, w, device="cuda")
    kl = torch.zeros(1, device="cuda")
    ops.reparam_fwd(both, 0, lvb, 0, eps, c, zb, mu_o, lv_o, kl)
    tol = 1e-6 if prec == "fp32" else 6e-3
    assert_close(nchw(zb.float()), z.detach(), tol, "reparam z", "nchw")
    assert torch.equal(mu_o, mu.detach()) and torch.equal(lv_o, lvc.detach())
    ref_kl = float((1 + lvc - mu.pow(2) - lvc.exp()).sum())
    assert abs(float(kl) - ref_kl) < 1e-4 * abs(ref_kl)
    dzb = torch.empty(n, h, w, c, dtype=dt, device="cuda")
    ops.pack_nchw(G, dzb)
    dmu, dlv = torch.empty_like(zb), torch.empty_like(zb)
    ops.reparam_bwd(both, 0, lvb, 0, eps, dzb, c, dmu, dlv, gmu_e, glv_e, 0.0)
    assert_close(nchw(dmu.float()), mu.grad, tol, "reparam dmu", "nchw")
    assert_close(nchw(dlv.float()), lv.grad, tol, "reparam dlv", "nchw")
    ops.reparam_bwd(both, 0, lvb, 0, eps, dzb, c, dmu, dlv, None, None, 0.0)
    assert_close(nchw(dmu.float()), mu_g1, tol, "reparam dmu (no ext)", "nchw")
    assert_close(nchw(dlv.float()), lv_g1, tol, "reparam dlv (no ext)", "nchw")


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_dhead(K, prec):
    """spectral-normalised head == <x, w/|w|> + b, with the gradient through sigma
    (torch/nn/utils/spectral_norm.py:92-114 with a 1 x K weight matrix)."""
    ops, L = K
    dt, n, c, h, w = DT[prec], 3, 512, 16, 16
    x = q(rnd(n, c, h, w, seed=41), dt).requires_grad_(True)
    wt = rnd(1, c, h, w, seed=42, scale=0.0867).requires_grad_(True)
    b = rnd(1, seed=43)
    sigma = wt.flatten().norm()
    s = F.conv2d(x, wt / sigma, b).view(-1)
    gs = rnd(n, seed=44)
    (s * gs).sum().backward()
    xb = torch.empty(n, h, w, c, dtype=dt, device="cuda")
    ops.pack_nchw(x.detach(), xb)
    w_khwc = wt.detach()[0].permute(1, 2, 0).contiguous().view(-1)
    score, wn2 = torch.empty(n, device="cuda"), torch.empty(1, device="cuda")
    ops.dhead_fwd(xb, w_khwc, b, score, wn2)
    assert_close(score, s.detach(), 1e-5 if prec == "fp32" else 1e-4, "dhead score", "n")
    dx = torch.empty_like(xb)
    dw, db = torch.zeros_like(w_khwc), torch.zeros(1, device="cuda")
    scratch = torch.empty(w_khwc.numel() + 8, device="cuda")
    ops.dhead_bwd(xb, w_khwc, wn2, gs, dx, dw, db, scratch)
    assert_close(nchw(dx.float()), x.grad, 1e-5 if prec == "fp32" else 6e-3, "dhead dx", "nchw")
    assert_close(dw.view(h, w, c).permute(2, 0, 1), wt.grad[0], 2e-5 if prec == "fp32" else 1e-3, "dhead dw", "chw")
    assert abs(float(db) - float(gs.sum())) < 1e-5
    # the same gradient written straight into an OIHW tensor (the .grad view of the flat buffer), accumulating
    g_oihw, db2 = torch.zeros(1, c, h, w, device="cuda"), torch.zeros(1, device="cuda")
    for _ in range(2):
        ops.dhead_bwd(xb, w_khwc, wn2, gs, dx, g_oihw, db2, scratch, dw_c=c)
    assert_close(g_oihw[0], 2 * wt.grad[0], 2e-5 if prec == "fp32" else 1e-3, "dhead dw (OIHW, accumulated)", "chw")
    assert abs(float(db2) - 2 * float(gs.sum())) < 1e-5


def test_dhead_prepare_power_iteration(K):
    """vcg_dhead_prepare == one power iteration of torch's spectral_norm on a 1 x K matrix (spectral_norm.py:92-114),
    the (h, w, c) filter copy and {sigma, |W|}; without do_iter: sigma of the STALE u, v (eval mode, :125-130)."""
    ops, L = K
    c, h, w = 512, 16, 16
    wt = rnd(1, c, h, w, seed=51, scale=0.0867)
    g = torch.Generator().manual_seed(52)
    u0 = F.normalize(torch.randn(1, generator=g), dim=0, eps=1e-12).cuda()
    v0 = F.normalize(torch.randn(c * h * w, generator=g), dim=0, eps=1e-12).cuda()
    wm = wt.reshape(1, -1)
    v_ref = F.normalize(torch.mv(wm.t(), u0), dim=0, eps=1e-12)
    u_ref = F.normalize(torch.mv(wm, v_ref), dim=0, eps=1e-12)
    u, v = u0.clone(), v0.clone()
    w_hwc, aux = torch.empty(c * h * w, device="cuda"), torch.empty(2, device="cuda")
    ops.dhead_prepare(wt, u, v, w_hwc, aux, True)
    assert torch.equal(u, u_ref) and rel_l2(v, v_ref) < 1e-6
    assert torch.equal(w_hwc, wt[0].permute(1, 2, 0).contiguous().view(-1))
    assert abs(float(aux[0]) - float(torch.dot(u_ref, torch.mv(wm, v_ref)))) < 1e-5 * float(wt.norm())
    assert abs(float(aux[1]) - float(wt.norm())) < 1e-5 * float(wt.norm())
    # idempotent: a second iteration on the result reproduces the same bits
    u2, v2 = u.clone(), v.clone()
    ops.dhead_prepare(wt, u2, v2, None, None, True)
    assert torch.equal(u2, u) and torch.equal(v2, v)
    # eval mode: stale u, v against a perturbed weight
    w2 = wt * 1.01 + 0.001
    ops.dhead_prepare(w2, u, v, None, aux, False)
    assert abs(float(aux[0]) - float(torch.dot(u, torch.mv(w2.reshape(1, -1), v)))) < 1e-5 * float(w2.norm())


def test_adam_multi(K):
    """bit-level agreement with torch.optim.Adam(betas=(0.5,0.999)) over 3 steps on ragged tensors."""
    import ctypes as C
    import numpy as np
    ops, L = K
    shapes = [(64, 3, 7, 7), (64,), (130001,), (3,), (257, 33)]
    ps = [rnd(*s, seed=50 + i) for i, s in enumerate(shapes)]
    ref = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.Adam(ref, lr=2e-4, betas=(0.5, 0.999), foreach=False)
    ms = [torch.zeros_like(p) for p in ps]
    vs = [torch.zeros_like(p) for p in ps]
    gs = [torch.zeros_like(p) for p in ps]
    chunks = []
    for p, g, m, v in zip(ps, gs, ms, vs):
        for o in range(0, p.numel(), 65536):
            cnt = min(65536, p.numel() - o)
            chunks.append((p.data_ptr() + 4 * o, g.data_ptr() + 4 * o, m.data_ptr() + 4 * o, v.data_ptr() + 4 * o, cnt))
    arr = (L.AdamChunk * len(chunks))()
    for i, ck in enumerate(chunks):
        arr[i].p, arr[i].g, arr[i].m, arr[i].v, arr[i].numel = ck
    table = torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy()).cuda()
    state = torch.zeros(4, device="cuda")
    for step in range(1, 4):
        for i, (r, g) in enumerate(zip(ref, gs)):
            g.copy_(rnd(*r.shape, seed=100 * step + i))
            r.grad = g.clone()
        opt.step()
        ops.adam_multi(table, len(chunks), state, 2e-4, 0.5, 0.999, 1e-8)
        assert float(state[0]) == step
        for r, p in zip(ref, ps):
            assert rel_l2(p, r.detach()) < 1e-7, (step, rel_l2(p, r.detach()))


def test_error_reporting(K):
    ops, L = K
    x = torch.zeros(1, 4, 4, 12, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(L.VcgError):
        ops.xform_fwd(x, 12, torch.zeros(1, 6, 6, 12, dtype=torch.bfloat16, device="cuda"), 0, 1)
