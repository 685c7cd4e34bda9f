"""Shared helpers for the GPU parity tests (torch fp32/fp64 references + error reporting)."""
import torch
import torch.nn.functional as F


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def describe_mismatch(got, ref, dims="nhwc"):
    """Human-readable error signature: where are the wrong entries? (for one-shot GPU debugging)."""
    got = got.double()
    ref = ref.double()
    err = (got - ref).abs()
    tol = 1e-2 * ref.abs().max().clamp_min(1e-6)
    bad = err > tol
    lines = [f"shape={tuple(ref.shape)} rel_l2={rel_l2(got, ref):.3e} max_abs={float(err.max()):.3e} "
             f"ref_absmax={float(ref.abs().max()):.3e} bad={int(bad.sum())}/{bad.numel()} "
             f"nan={int(torch.isnan(got).sum())}"]
    if bad.any() and ref.dim() == len(dims):
        for d, name in enumerate(dims):
            other = [i for i in range(ref.dim()) if i != d]
            frac = bad.float().mean(dim=other)
            idx = torch.nonzero(frac > 0).flatten().tolist()
            lines.append(f"  bad along {name}: {len(idx)}/{ref.shape[d]} indices, first {idx[:12]}, "
                         f"max frac {float(frac.max()):.2f}")
        i = torch.nonzero(bad)[0].tolist()
        lines.append(f"  first bad {i}: got {float(got[tuple(i)]):.5f} ref {float(ref[tuple(i)]):.5f}")
    return "\n".join(lines)


def assert_close(got, ref, tol, what, dims="nhwc"):
    r = rel_l2(got, ref)
    assert r <= tol and not torch.isnan(got).any(), f"{what}: rel_l2 {r:.3e} > {tol:.1e}\n" + describe_mismatch(got, ref, dims)
    return r


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def ref_unshuffle_phys(x):
    """PixelUnshuffle(2) in the kernels' physical channel order (i,j,c) instead of c*4+i*2+j."""
    n, c, h, w = x.shape
    u = F.pixel_unshuffle(x, 2).view(n, c, 4, h // 2, w // 2)
    return u.permute(0, 2, 1, 3, 4).reshape(n, 4 * c, h // 2, w // 2)


def ref_xform(x, mode, pad, norm=False, act=0, residual=None):
    """torch reference of vcg_xform_fwd on NCHW tensors; returns NCHW in the kernels' channel order."""
    v = x
    if norm:
        v = F.instance_norm(v, eps=1e-5)
    if act == 1:
        v = F.relu(v)
    elif act == 2:
        v = F.leaky_relu(v, 0.2)
    elif act == 3:
        v = torch.tanh(v)
    elif act == 4:
        v = torch.sigmoid(v)
    if residual is not None:
        v = v + residual
    if mode == 0:
        return F.pad(v, (pad,) * 4, mode="reflect") if pad else v
    if mode == 1:
        v = F.pixel_shuffle(v, 2)
        return F.pad(v, (pad,) * 4, mode="reflect") if pad else v
    if mode == 2:
        v = ref_unshuffle_phys(v)
        return F.pad(v, (pad,) * 4, mode="reflect") if pad else v
    v = F.pad(v, (pad,) * 4, mode="reflect") if pad else v
    return ref_unshuffle_phys(v)
