"""Tensor-level wrappers over the C ABI (one function per exported kernel family).

Every function takes pre-allocated torch CUDA tensors and enqueues on the current stream; nothing here
allocates except where noted, and nothing falls back to PyTorch math."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import lib as L


def rup(x, m):
    return (x + m - 1) // m * m


class _Prof:
    """Optional CUDA-event instrumentation of the kernel launches (bench.py's roofline leg).  external=True creates
    the events with cudaEventRecordExternal semantics: recorded while a CUDA graph is being captured they become
    event-record NODES of the graph, re-recorded by every replay, so the kernels of a replayed step can be timed."""
    records = None
    external = False


def prof_begin(external=False):
    _Prof.records = []
    _Prof.external = bool(external)


def prof_end():
    r, _Prof.records = _Prof.records, None
    return r


def _rec(kind, flops, tag=""):
    if _Prof.records is None:
        return None
    s = torch.cuda.Event(enable_timing=True, external=_Prof.external)
    s.record()
    s.vcg_stream = torch.cuda.current_stream().cuda_stream      # which stream the launch went to (tools/step_timeline.py)
    return (kind, flops, tag, s)


def _rec_end(tok):
    if tok is not None:
        e = torch.cuda.Event(enable_timing=True, external=_Prof.external)
        e.record()
        _Prof.records.append((*tok, e))


def _profiled(kind):
    """event-time a non-GEMM wrapper when profiling is on (family = `kind`, flops field carries bytes = 0)"""
    def deco(fn):
        def wrapper(*a, **k):
            if _Prof.records is None:
                return fn(*a, **k)
            tok = _rec(kind, 0.0, "")
            r = fn(*a, **k)
            _rec_end(tok)
            return r
        wrapper.__name__ = fn.__name__
        wrapper.__doc__ = fn.__doc__
        return wrapper
    return deco


@dataclass(frozen=True)
class ConvSpec:
    """Geometry of one reference convolution as executed by the kernels.

    co/ci/kh/kw are the reference OIHW dims (e.g. Networks.py:87 -> (128, 256, 3, 3)); wmap says how the
    executed input channels map onto them (PixelUnshuffle order / space-to-depth for stride 2)."""
    co: int
    ci: int
    kh: int
    kw: int
    wmap: int = L.WMAP_PLAIN
    c_phys: int = 0          # physical input channels as executed (0 -> derived)

    @property
    def pkh(self):
        return self.kh // 2 if self.wmap == L.WMAP_S2D else self.kh

    @property
    def pkw(self):
        return self.kw // 2 if self.wmap == L.WMAP_S2D else self.kw

    @property
    def cin_phys(self):
        if self.c_phys:
            return self.c_phys
        base = self.ci * 4 if self.wmap == L.WMAP_S2D else self.ci
        return rup(base, 8)

    @property
    def kwc_pad(self):
        return rup(self.pkw * self.cin_phys, 64)

    @property
    def cout_pad(self):
        return rup(self.co, 16)

    @property
    def out_c(self):
        return rup(self.co, 8)

    # data-gradient view: rows = physical input channels, K = (taps, physical output channels)
    @property
    def d_rows_pad(self):
        return rup(self.cin_phys, 16)

    @property
    def d_kwc_pad(self):
        return rup(self.pkw * self.out_c, 64)

    def wpack_desc(self, dtype, transpose_flip):
        return L.WpackDesc(dtype=L.dtype_code(dtype), co=self.co, ci=self.ci, kh=self.kh, kw=self.kw, wmap=self.wmap,
                           c_phys=self.cin_phys, co_phys=self.out_c,
                           rows_pad=self.d_rows_pad if transpose_flip else self.cout_pad,
                           pkh=self.pkh, pkw=self.pkw,
                           kwc_pad=self.d_kwc_pad if transpose_flip else self.kwc_pad,
                           transpose_flip=1 if transpose_flip else 0)

    def flops(self, n, ho, wo):
        """algorithmic FLOPs of one pass (forward, data-gradient or weight-gradient): 2*M*N*K with the
        reference's logical dims"""
        return 2.0 * n * ho * wo * self.co * self.ci * self.kh * self.kw

    def packed_shape(self, transpose_flip=False):
        if transpose_flip:
            return (self.d_rows_pad, self.pkh, self.d_kwc_pad)
        return (self.cout_pad, self.pkh, self.kwc_pad)


def _thin_kernel(dtype, c, cout, cout_pad, kh, kw, wo, has_stats):
    """mirror of vcg_conv_fold_supported (csrc/conv_tc_fold.cu): which kernel a bf16 conv launch runs on
    ('fold' or '' = conv_tc_kernel); used only to tag profiling records"""
    if dtype != torch.bfloat16 or has_stats or c % 64:
        return ""
    co8 = rup(cout, 8)
    bn = rup(kw * co8, 16)
    if cout <= 8 and wo >= 64 and kw in (2, 3, 7) and co8 <= cout_pad and bn <= 256 and min(8, 512 // bn) >= kh + 1 and \
            kh * (c // 64) * bn * 128 + 2 * kw * ((cout + 3) // 4) * 2048 + 2048 + 3 * 16384 <= 227 * 1024:
        return "fold"
    return ""


@_profiled("wpack")
def wpack(spec: ConvSpec, w_oihw, out, transpose_flip=False):
    d = spec.wpack_desc(out.dtype, transpose_flip)
    assert w_oihw.dtype == torch.float32 and w_oihw.is_contiguous()
    assert tuple(out.shape) == spec.packed_shape(transpose_flip)
    L.check(L.load().vcg_wpack(C.byref(d), L.ptr(w_oihw), L.ptr(out), L.stream_ptr()), "vcg_wpack")
    return out


def _to_device_bytes(carray, device):
    import numpy as np
    return torch.from_numpy(np.frombuffer(bytes(carray), dtype=np.uint8).copy()).to(device)


def wjob_table(entries, device, accumulate=True):
    """entries: list of (spec, oihw fp32 tensor, packed fwd tensor, packed data-gradient tensor or None).
    Returns (device byte tensor holding the vcg_wjob array, njobs, total_tiles) for vcg_wpack_multi /
    vcg_wunpack_multi; build once and cache (the pointers must stay valid)."""
    arr = (L.WJob * len(entries))()
    for j, (spec, oihw, packed, packed_t) in enumerate(entries):
        d = spec.wpack_desc(packed.dtype, False)
        for f in ("co", "ci", "kh", "kw", "wmap", "c_phys", "co_phys", "rows_pad", "pkh", "pkw", "kwc_pad"):
            setattr(arr[j], f, getattr(d, f))
        assert oihw.dtype == torch.float32 and oihw.is_contiguous()
        assert tuple(packed.shape) == spec.packed_shape(False)
        arr[j].oihw, arr[j].packed = oihw.data_ptr(), packed.data_ptr()
        if packed_t is not None:
            assert tuple(packed_t.shape) == spec.packed_shape(True) and packed_t.dtype == packed.dtype
            arr[j].packed_t, arr[j].t_kwc_pad = packed_t.data_ptr(), spec.d_kwc_pad
        arr[j].accumulate = 1 if accumulate else 0
    total = C.c_int32(0)
    L.check(L.load().vcg_wjob_plan(arr, len(entries), C.byref(total)), "vcg_wjob_plan")
    return _to_device_bytes(arr, device), len(entries), int(total.value)


@_profiled("wpack")
def wpack_multi(table, dtype):
    """OIHW fp32 masters -> both kernel layouts of every filter in the table, one launch"""
    dev, n, tiles = table
    L.check(L.load().vcg_wpack_multi(L.dtype_code(dtype), L.ptr(dev), n, tiles, L.stream_ptr()), "vcg_wpack_multi")


@_profiled("wunpack_grad")
def wunpack_multi(table):
    """grad_oihw += unpack(dw); dw = 0 for every filter in the table, one launch"""
    dev, n, tiles = table
    L.check(L.load().vcg_wunpack_multi(L.ptr(dev), n, tiles, L.stream_ptr()), "vcg_wunpack_multi")


def vecjob_table(entries, device):
    """entries: list of (src fp32 tensor, dst fp32 tensor, n)"""
    arr = (L.VecJob * len(entries))()
    for j, (src, dst, n) in enumerate(entries):
        arr[j].src, arr[j].dst, arr[j].n = src.data_ptr(), dst.data_ptr(), n
    return _to_device_bytes(arr, device), len(entries)


@_profiled("wunpack_grad")
def vecflush_multi(table):
    dev, n = table
    L.check(L.load().vcg_vecflush_multi(L.ptr(dev), n, L.stream_ptr()), "vcg_vecflush_multi")


@_profiled("wunpack_grad")
def wunpack_grad(spec: ConvSpec, dw_packed, grad_oihw, accumulate=False):
    d = spec.wpack_desc(torch.float32, False)
    assert dw_packed.dtype == torch.float32 and grad_oihw.dtype == torch.float32 and grad_oihw.is_contiguous()
    L.check(L.load().vcg_wunpack_grad(C.byref(d), L.ptr(dw_packed), L.ptr(grad_oihw), 1 if accumulate else 0,
                                      L.stream_ptr()), "vcg_wunpack_grad")
    return grad_oihw


def conv_fwd(spec: ConvSpec, x_pad, w_packed, bias, y, stats_acc=None, act=L.ACT_NONE):
    """y[n,ho,wo,out_c] = act(conv_valid(x_pad, w) + bias); x_pad: [n,hp,wp,cin_phys]."""
    n, hp, wp, c = x_pad.shape
    assert c == spec.cin_phys, (c, spec.cin_phys)
    d = L.ConvDesc(dtype=L.dtype_code(x_pad.dtype), n=n, hp=hp, wp=wp, c=c, kh=spec.pkh, kw=spec.pkw,
                   kwc_pad=spec.kwc_pad, cout=spec.co, cout_pad=spec.cout_pad, out_c=y.shape[-1], act=act,
                   stats=1 if stats_acc is not None else 0, flat=0,
                   out_f32=1 if (y.dtype == torch.float32 and x_pad.dtype != torch.float32) else 0)
    assert tuple(y.shape[:3]) == (n, hp - spec.pkh + 1, wp - spec.pkw + 1), (y.shape, x_pad.shape)
    thin = _thin_kernel(x_pad.dtype, c, spec.co, spec.cout_pad, spec.pkh, spec.pkw, y.shape[2], stats_acc is not None)
    tok = _rec(f"conv_{thin}_fwd" if thin else "conv_fwd", spec.flops(n, y.shape[1], y.shape[2]),
               f"{spec.ci}->{spec.co} k{spec.kh} @{y.shape[1]}")
    L.check(L.load().vcg_conv_fwd(C.byref(d), L.ptr(x_pad), L.ptr(w_packed), L.ptr(bias), L.ptr(y), L.ptr(stats_acc),
                                  L.stream_ptr()), "vcg_conv_fwd")
    _rec_end(tok)
    return y


def conv_dgrad(spec: ConvSpec, dy_pad, w_dgrad, dxp):
    """dxp[n,hp,wp,cin_phys] = full correlation of the zero-haloed dy with the flipped filter."""
    n, hd, wd, c = dy_pad.shape
    assert c == spec.out_c
    # logical rows of the data-gradient GEMM: the padding channels of a 3-channel image input have all-zero filter
    # rows, so their gradient is written as literal zeros instead of being computed
    d_cout = spec.ci if (spec.wmap == L.WMAP_PLAIN and spec.ci < spec.cin_phys) else spec.cin_phys
    d = L.ConvDesc(dtype=L.dtype_code(dy_pad.dtype), n=n, hp=hd, wp=wd, c=c, kh=spec.pkh, kw=spec.pkw,
                   kwc_pad=spec.d_kwc_pad, cout=d_cout, cout_pad=spec.d_rows_pad, out_c=dxp.shape[-1],
                   act=L.ACT_NONE, stats=0, flat=2, out_f32=0)   # 2: dy_pad's outer (k-1) border is zero
    assert tuple(dxp.shape[:3]) == (n, hd - spec.pkh + 1, wd - spec.pkw + 1), (dxp.shape, dy_pad.shape)
    thin = _thin_kernel(dy_pad.dtype, c, d_cout, spec.d_rows_pad, spec.pkh, spec.pkw, wd - spec.pkw + 1, False)
    tok = _rec(f"conv_{thin}_dgrad" if thin else "conv_dgrad", spec.flops(n, hd - 2 * (spec.pkh - 1), wd - 2 * (spec.pkw - 1)),
               f"{spec.ci}->{spec.co} k{spec.kh} @{hd - 2 * (spec.pkh - 1)}")
    L.check(L.load().vcg_conv_fwd(C.byref(d), L.ptr(dy_pad), L.ptr(w_dgrad), None, L.ptr(dxp), None, L.stream_ptr()),
            "vcg_conv_fwd(dgrad)")
    _rec_end(tok)
    return dxp


def conv_wgrad(spec: ConvSpec, x_pad, dy_pad, dw_packed, allow_simt=False):
    """dw_packed[cout_pad,kh,kwc_pad] += wgrad(x_pad, dy) ; dy_pad carries a zero halo of (k-1).
    allow_simt: shapes no tensor-core kernel takes may run on the (slow) SIMT kernel instead of raising."""
    n, hp, wp, c = x_pad.shape
    halo_h, halo_w = spec.pkh - 1, spec.pkw - 1
    assert halo_h == halo_w
    d = L.ConvDesc(dtype=L.dtype_code(x_pad.dtype), n=n, hp=hp, wp=wp, c=c, kh=spec.pkh, kw=spec.pkw,
                   kwc_pad=spec.kwc_pad, cout=spec.co, cout_pad=spec.cout_pad, out_c=dy_pad.shape[-1], act=0, stats=0,
                   flat=L.WGRAD_ALLOW_SIMT if allow_simt else 0, out_f32=0)
    assert dy_pad.shape[1] == hp - spec.pkh + 1 + 2 * halo_h
    tok = _rec("conv_wgrad", spec.flops(n, hp - spec.pkh + 1, wp - spec.pkw + 1), f"{spec.ci}->{spec.co} k{spec.kh} @{hp - spec.pkh + 1}")
    L.check(L.load().vcg_conv_wgrad(C.byref(d), L.ptr(x_pad), L.ptr(dy_pad), halo_h, dy_pad.shape[-1], L.ptr(dw_packed),
                                    L.stream_ptr()), "vcg_conv_wgrad")
    _rec_end(tok)
    return dw_packed


@_profiled("in_stats")
def in_stats(y, c, mean_rstd):
    """mean/rstd [n,c,2] of a dense NHWC tensor (fp64 accumulation).  mean_rstd: float32 [n*c*2*3]."""
    n, h, w, cp = y.shape
    assert mean_rstd.numel() >= n * c * 6
    L.check(L.load().vcg_in_stats(L.dtype_code(y.dtype), L.ptr(y), n, h * w, c, cp, L.ptr(mean_rstd), L.stream_ptr()),
            "vcg_in_stats")
    return mean_rstd


@_profiled("in_finalize")
def in_finalize(sums, nc, hw, mean_rstd):
    L.check(L.load().vcg_in_finalize(L.ptr(sums), nc, hw, L.ptr(mean_rstd), L.stream_ptr()), "vcg_in_finalize")
    return mean_rstd


def xform_dst_shape(n, h, w, c, mode, pad, dst_c=None):
    if mode == L.MODE_PLAIN:
        hd, wd, cd = h + 2 * pad, w + 2 * pad, c
    elif mode == L.MODE_SHUFFLE:
        hd, wd, cd = 2 * h + 2 * pad, 2 * w + 2 * pad, c // 4
    elif mode == L.MODE_UNSHUFFLE:
        hd, wd, cd = h // 2 + 2 * pad, w // 2 + 2 * pad, c * 4
    else:
        hd, wd, cd = (h + 2 * pad) // 2, (w + 2 * pad) // 2, c * 4
    return (n, hd, wd, dst_c or rup(cd, 8))


def xform_fwd(src, c, dst, mode, pad, mean_rstd=None, act=L.ACT_NONE, residual=None, res_off=0, stats_hw=0):
    n, h, w, src_c = src.shape
    tok = _rec("xform_fwd", float(src[..., :c].numel() * src.element_size() + dst.numel() * dst.element_size()),
               f"c{c} {h}x{w} mode{mode} pad{pad}{' norm' if mean_rstd is not None else ''}{' res' if residual is not None else ''}")
    d = L.XformDesc(dtype=L.dtype_code(src.dtype), n=n, h=h, w=w, c=c, src_c=src_c,
                    norm=1 if mean_rstd is not None else 0, act=act, mode=mode, pad=pad, dst_c=dst.shape[-1],
                    res_hp=residual.shape[1] if residual is not None else 0,
                    res_wp=residual.shape[2] if residual is not None else 0,
                    res_c=residual.shape[3] if residual is not None else 0, res_off=res_off, stats_hw=stats_hw)
    assert tuple(dst.shape) == xform_dst_shape(n, h, w, c, mode, pad, dst.shape[-1]), (dst.shape, src.shape, mode, pad)
    L.check(L.load().vcg_xform_fwd(C.byref(d), L.ptr(src), L.ptr(mean_rstd), L.ptr(residual), L.ptr(dst), L.stream_ptr()),
            "vcg_xform_fwd")
    _rec_end(tok)
    return dst


def _xb_desc(y_like_dtype, n, h, w, c, y_c, norm, act, pre_act, dy, dy_halo, nsrc, stats_hw=0, clear_halo=0):
    return L.XbwdDesc(dtype=L.dtype_code(y_like_dtype), n=n, h=h, w=w, c=c, y_c=y_c, norm=1 if norm else 0, act=act,
                      pre_act=pre_act, dy_halo=dy_halo, dy_c=dy.shape[-1], nsrc=nsrc, stats_hw=stats_hw,
                      clear_halo=1 if clear_halo else 0)


def xform_bwd_gather(srcs, y, n, h, w, c, dy, dy_halo, mean_rstd=None, act=L.ACT_NONE, pre_act=L.ACT_NONE,
                     gsums=None, dbias=None, stats_hw=0, clear_halo=False):
    """srcs: list of (dxp tensor, mode, pad[, folded]).  Writes g into the interior of dy (+ sums for phase 2).
    folded=True: fold_halo_ already added the reflect halo of dxp into its interior (fast kernel)."""
    tok = _rec("xform_bwd_gather", float(sum(s[0].numel() * s[0].element_size() for s in srcs) + 2 * n * h * w * c * dy.element_size()),
               f"c{c} {h}x{w} modes{[s[1] for s in srcs]}{' norm' if mean_rstd is not None else ''}")
    arr = (L.GSrc * max(1, len(srcs)))()
    for i, src in enumerate(srcs):
        t, mode, pad = src[:3]
        arr[i].dxp = t.data_ptr()
        arr[i].mode, arr[i].pad, arr[i].c_pitch = mode, pad, t.shape[-1]
        arr[i].folded = 1 if (len(src) > 3 and src[3]) else 0
    norm = mean_rstd is not None
    d = _xb_desc(dy.dtype, n, h, w, c, y.shape[-1] if y is not None else 8, norm, act, pre_act, dy, dy_halo, len(srcs),
                 stats_hw, clear_halo)
    L.check(L.load().vcg_xform_bwd_gather(C.byref(d), arr, L.ptr(y), L.ptr(mean_rstd), L.ptr(dy), L.ptr(gsums),
                                          L.ptr(dbias), L.stream_ptr()), "vcg_xform_bwd_gather")
    _rec_end(tok)
    return dy


@_profiled("fold_halo")
def fold_halo_(dxp, mode, pad, h, w, c):
    """in place: add the reflect halo of a consumer's padded-input gradient into its interior.
    (h, w, c) = dims of the activation the consumer reads (before its shuffle / unshuffle / s2d)."""
    L.check(L.load().vcg_fold_halo(L.dtype_code(dxp.dtype), L.ptr(dxp), dxp.shape[0], h, w, c, mode, pad, dxp.shape[-1],
                                   L.stream_ptr()), "vcg_fold_halo")
    return dxp


def xform_bwd_norm(y, n, h, w, c, dy, dy_halo, mean_rstd, gsums, pre_act=L.ACT_NONE, dbias=None, stats_hw=0):
    tok = _rec("xform_bwd_norm", float(3 * n * h * w * c * dy.element_size()), f"c{c} {h}x{w}")
    d = _xb_desc(dy.dtype, n, h, w, c, y.shape[-1], True, L.ACT_NONE, pre_act, dy, dy_halo, 0, stats_hw)
    L.check(L.load().vcg_xform_bwd_norm(C.byref(d), L.ptr(y), L.ptr(mean_rstd), L.ptr(gsums), L.ptr(dy), L.ptr(dbias),
                                        L.stream_ptr()), "vcg_xform_bwd_norm")
    _rec_end(tok)
    return dy


@_profiled("pack_nchw")
def pack_nchw(src, dst, halo=0):
    n, c, h, w = src.shape
    assert src.dtype == torch.float32 and src.is_contiguous()
    assert tuple(dst.shape[:3]) == (n, h + 2 * halo, w + 2 * halo)
    L.check(L.load().vcg_pack_nchw(L.dtype_code(dst.dtype), L.ptr(src), n, c, h, w, L.ptr(dst), dst.shape[-1], halo,
                                   L.stream_ptr()), "vcg_pack_nchw")
    return dst


@_profiled("unpack_nchw")
def unpack_nchw(src, c, dst, c_off=0):
    n, h, w, src_c = src.shape
    assert dst.dtype == torch.float32 and dst.is_contiguous() and tuple(dst.shape) == (n, c, h, w)
    base = C.c_void_p(src.data_ptr() + c_off * src.element_size())
    L.check(L.load().vcg_unpack_nchw(L.dtype_code(src.dtype), base, src_c, n, c, h, w, L.ptr(dst), L.stream_ptr()),
            "vcg_unpack_nchw")
    return dst


@_profiled("zero_")
def zero_(t):
    L.check(L.load().vcg_zero(L.ptr(t), t.numel() * t.element_size(), L.stream_ptr()), "vcg_zero")
    return t


@_profiled("zero_halo")
def zero_halo_(t, halo):
    """clear only the halo ring of an NHWC buffer [n, h+2*halo, w+2*halo, c]"""
    n, hp, wp, c = t.shape
    L.check(L.load().vcg_zero_halo(L.dtype_code(t.dtype), L.ptr(t), n, hp - 2 * halo, wp - 2 * halo, c, halo,
                                   L.stream_ptr()), "vcg_zero_halo")
    return t


@_profiled("l1_fwd_bwd")
def l1_fwd_bwd(a, b, out_sum, grad=None, scale=0.0):
    L.check(L.load().vcg_l1_fwd_bwd(L.ptr(a), L.ptr(b), a.numel(), scale, L.ptr(out_sum), L.ptr(grad), L.stream_ptr()),
            "vcg_l1_fwd_bwd")


@_profiled("mse_const_fwd_bwd")
def mse_const_fwd_bwd(d, target, out_sum, grad=None, scale=0.0):
    L.check(L.load().vcg_mse_const_fwd_bwd(L.ptr(d), d.numel(), float(target), scale, L.ptr(out_sum), L.ptr(grad),
                                           L.stream_ptr()), "vcg_mse_const_fwd_bwd")


@_profiled("kl_fwd_bwd")
def kl_fwd_bwd(mu, lv, out_sum, gmu=None, glv=None, scale=0.0):
    L.check(L.load().vcg_kl_fwd_bwd(L.ptr(mu), L.ptr(lv), mu.numel(), scale, L.ptr(out_sum), L.ptr(gmu), L.ptr(glv),
                                    L.stream_ptr()), "vcg_kl_fwd_bwd")


def _off(t, c_off):
    return C.c_void_p(t.data_ptr() + c_off * t.element_size())


@_profiled("reparam_fwd")
def reparam_fwd(mu_src, mu_off, lv_src, lv_off, eps, c, z, mu_out, lv_out, kl_sum=None):
    n, h, w, _ = mu_src.shape
    assert mu_src.dtype == torch.float32 and lv_src.dtype == torch.float32
    L.check(L.load().vcg_reparam_fwd(L.dtype_code(z.dtype), _off(mu_src, mu_off), mu_src.shape[-1],
                                     _off(lv_src, lv_off), lv_src.shape[-1], L.ptr(eps), n, h * w, c, L.ptr(z),
                                     L.ptr(mu_out), L.ptr(lv_out), L.ptr(kl_sum), L.stream_ptr()), "vcg_reparam_fwd")


@_profiled("reparam_bwd")
def reparam_bwd(mu_src, mu_off, lv_src, lv_off, eps, dz, c, dmu, dlv, gmu_ext=None, glv_ext=None, kl_scale=0.0):
    n, h, w, _ = mu_src.shape
    assert mu_src.dtype == torch.float32 and lv_src.dtype == torch.float32
    L.check(L.load().vcg_reparam_bwd(L.dtype_code(dz.dtype), _off(mu_src, mu_off), mu_src.shape[-1],
                                     _off(lv_src, lv_off), lv_src.shape[-1], L.ptr(eps), L.ptr(dz), dz.shape[-1],
                                     L.ptr(gmu_ext), L.ptr(glv_ext), kl_scale, n, h * w, c, L.ptr(dmu), dmu.shape[-1],
                                     L.ptr(dlv), dlv.shape[-1], L.stream_ptr()), "vcg_reparam_bwd")


@_profiled("dhead_fwd")
def dhead_fwd(x, w_khwc, bias, score, wnorm2):
    n = x.shape[0]
    k = x[0].numel()
    L.check(L.load().vcg_dhead_fwd(L.dtype_code(x.dtype), L.ptr(x), L.ptr(w_khwc), L.ptr(bias), n, k, L.ptr(score),
                                   L.ptr(wnorm2), L.stream_ptr()), "vcg_dhead_fwd")


@_profiled("dhead_bwd")
def dhead_bwd(x, w_khwc, wnorm2, gscore, dx, dw, dbias, scratch, dw_c=0):
    """dw_c = 0: dw is a (h, w, c)-ordered vector; dw_c = C: dw is the OIHW gradient tensor of the [1, C, kh, kw] filter
    (accumulated in place, e.g. a view of the optimiser's flat gradient buffer)."""
    n = x.shape[0]
    k = x[0].numel()
    if dw is not None:
        assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == k
    L.check(L.load().vcg_dhead_bwd(L.dtype_code(x.dtype), L.ptr(x), L.ptr(w_khwc), L.ptr(wnorm2), L.ptr(gscore), n, k,
                                   L.ptr(dx), L.ptr(dw), L.ptr(dbias), L.ptr(scratch), dw_c, L.stream_ptr()), "vcg_dhead_bwd")


@_profiled("dhead_prepare")
def dhead_prepare(w_oihw, u, v, w_hwc, aux, do_iter, scratch=None):
    """spectral-norm power iteration (in place on u, v when do_iter), (h, w, c) filter copy, aux = {sigma, |W|};
    scratch: 3 doubles of caller-owned device memory (allocated here when omitted)"""
    co, c, kh, kw = w_oihw.shape
    assert co == 1 and w_oihw.dtype == torch.float32 and w_oihw.is_contiguous()
    if scratch is None:
        scratch = torch.empty(3, dtype=torch.float64, device=w_oihw.device)
    L.check(L.load().vcg_dhead_prepare(L.ptr(w_oihw), c, kh * kw, L.ptr(u), L.ptr(v), L.ptr(w_hwc), L.ptr(aux),
                                       L.ptr(scratch), 1 if do_iter else 0, L.stream_ptr()), "vcg_dhead_prepare")


@_profiled("adam_multi")
def adam_multi(chunks_dev, nchunks, state_dev, lr, beta1, beta2, eps, grad_scale=1.0, flags=L.ADAM_TICK):
    """state_dev: float32[4] device tensor {step, lr/bc1, sqrt(bc2), -}; ADAM_TICK advances step by one."""
    L.check(L.load().vcg_adam_multi(L.ptr(chunks_dev), nchunks, L.ptr(state_dev), lr, beta1, beta2, eps, grad_scale,
                                    flags, L.stream_ptr()), "vcg_adam_multi")


@_profiled("cast_bf16")
def cast_bf16(src, dst, zero_src=False):
    """dst (bf16) = src (fp32), optionally src = 0: the gradient wire buffer of the data-parallel all-reduce"""
    assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and src.numel() == dst.numel()
    assert src.is_contiguous() and dst.is_contiguous()
    L.check(L.load().vcg_cast_bf16(L.ptr(src), L.ptr(dst), src.numel(), 1 if zero_src else 0, L.stream_ptr()), "vcg_cast_bf16")
    return dst


def probe_tmap(base_tensor, dims, strides_bytes, box):
    r = len(dims)
    return L.load().vcg_probe_tmap(L.ptr(base_tensor), r, (C.c_uint64 * r)(*dims), (C.c_uint64 * max(1, r - 1))(*strides_bytes),
                                   (C.c_uint32 * r)(*box))
