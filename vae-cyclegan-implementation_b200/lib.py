"""ctypes binding of libvcg_b200.so (include/vcg.h).  Fails loudly when the library is missing:
there is no eager/PyTorch fallback for any kernel."""
from __future__ import annotations

import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvcg_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
MODE_PLAIN, MODE_SHUFFLE, MODE_UNSHUFFLE, MODE_PAD_S2D = 0, 1, 2, 3
WMAP_PLAIN, WMAP_UNSHUFFLE, WMAP_S2D = 0, 1, 2
ADAM_TICK, ADAM_GRAD_BF16, ADAM_ZERO_GRAD = 1, 2, 4
WGRAD_ALLOW_SIMT = 4

i32 = C.c_int32


def _fields(*names):
    return [(n, i32) for n in names]


class ConvDesc(C.Structure):
    _fields_ = _fields("dtype", "n", "hp", "wp", "c", "kh", "kw", "kwc_pad", "cout", "cout_pad", "out_c",
                       "act", "stats", "flat", "out_f32")


class WpackDesc(C.Structure):
    _fields_ = _fields("dtype", "co", "ci", "kh", "kw", "wmap", "c_phys", "co_phys", "rows_pad", "pkh", "pkw",
                       "kwc_pad", "transpose_flip")


class XformDesc(C.Structure):
    _fields_ = _fields("dtype", "n", "h", "w", "c", "src_c", "norm", "act", "mode", "pad", "dst_c",
                       "res_hp", "res_wp", "res_c", "res_off", "stats_hw")


class GSrc(C.Structure):
    _fields_ = [("dxp", C.c_void_p), ("mode", i32), ("pad", i32), ("c_pitch", i32), ("folded", i32)]


class XbwdDesc(C.Structure):
    _fields_ = _fields("dtype", "n", "h", "w", "c", "y_c", "norm", "act", "pre_act", "dy_halo", "dy_c", "nsrc", "stats_hw", "clear_halo")


class WJob(C.Structure):
    _fields_ = _fields("co", "ci", "kh", "kw", "wmap", "c_phys", "co_phys", "rows_pad", "pkh", "pkw", "kwc_pad",
                       "transpose_flip") + [("oihw", C.c_void_p), ("packed", C.c_void_p), ("packed_t", C.c_void_p)] + \
        _fields("t_kwc_pad", "accumulate", "co_t", "tiles_ci", "tile0", "ntiles", "vec", "reserved")


class VecJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("n", i32), ("pad", i32)]


class AdamChunk(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("numel", i32)]


ADAM_CHUNK_BYTES = C.sizeof(AdamChunk)

_lib = None

_SIGS = {
    "vcg_version": (C.c_int, []),
    "vcg_last_error": (C.c_char_p, []),
    "vcg_launch_count": (C.c_longlong, []),
    "vcg_set_sm_budget": (C.c_int, [i32]),
    "vcg_set_l2_prefetch": (C.c_int, [i32]),
    "vcg_conv_fwd": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vcg_conv_wgrad": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, i32, i32, C.c_void_p, C.c_void_p]),
    "vcg_wpack": (C.c_int, [C.POINTER(WpackDesc), C.c_void_p, C.c_void_p, C.c_void_p]),
    "vcg_wunpack_grad": (C.c_int, [C.POINTER(WpackDesc), C.c_void_p, C.c_void_p, i32, C.c_void_p]),
    "vcg_wjob_plan": (C.c_int, [C.POINTER(WJob), i32, C.POINTER(i32)]),
    "vcg_wpack_multi": (C.c_int, [i32, C.c_void_p, i32, i32, C.c_void_p]),
    "vcg_wunpack_multi": (C.c_int, [C.c_void_p, i32, i32, C.c_void_p]),
    "vcg_vecflush_multi": (C.c_int, [C.c_void_p, i32, C.c_void_p]),
    "vcg_in_stats": (C.c_int, [i32, C.c_void_p, i32, i32, i32, i32, C.c_void_p, C.c_void_p]),
    "vcg_in_finalize": (C.c_int, [C.c_void_p, i32, i32, C.c_void_p, C.c_void_p]),
    "vcg_xform_fwd": (C.c_int, [C.POINTER(XformDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vcg_xform_bwd_gather": (C.c_int, [C.POINTER(XbwdDesc), C.POINTER(GSrc), C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "vcg_fold_halo": (C.c_int, [i32, C.c_void_p, i32, i32, i32, i32, i32, i32, i32, C.c_void_p]),
    "vcg_xform_bwd_norm": (C.c_int, [C.POINTER(XbwdDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "vcg_pack_nchw": (C.c_int, [i32, C.c_void_p, i32, i32, i32, i32, C.c_void_p, i32, i32, C.c_void_p]),
    "vcg_unpack_nchw": (C.c_int, [i32, C.c_void_p, i32, i32, i32, i32, i32, C.c_void_p, C.c_void_p]),
    "vcg_reparam_fwd": (C.c_int, [i32, C.c_void_p, i32, C.c_void_p, i32, C.c_void_p, i32, i32, i32, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vcg_reparam_bwd": (C.c_int, [i32, C.c_void_p, i32, C.c_void_p, i32, C.c_void_p, C.c_void_p, i32, C.c_void_p,
                                  C.c_void_p, C.c_float, i32, i32, i32, C.c_void_p, i32, C.c_void_p, i32, C.c_void_p]),
    "vcg_l1_fwd_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vcg_mse_const_fwd_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vcg_kl_fwd_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "vcg_dhead_fwd": (C.c_int, [i32, C.c_void_p, C.c_void_p, C.c_void_p, i32, i32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vcg_dhead_bwd": (C.c_int, [i32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, i32, i32, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, i32, C.c_void_p]),
    "vcg_dhead_prepare": (C.c_int, [C.c_void_p, i32, i32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, i32,
                                    C.c_void_p]),
    "vcg_adam_multi": (C.c_int, [C.c_void_p, i32, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                 i32, C.c_void_p]),
    "vcg_cast_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, i32, C.c_void_p]),
    "vcg_probe_tmap": (C.c_int, [C.c_void_p, i32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "vcg_zero": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "vcg_zero_halo": (C.c_int, [i32, C.c_void_p, i32, i32, i32, i32, i32, C.c_void_p]),
}
EXPORTS = tuple(_SIGS)


def load():
    """dlopen the library and type every entry point.  Raises if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python vae-cyclegan-implementation_b200/build.py` "
                "(or __graft_entry__.build()).  There is no fallback path.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.vcg_version() != 2:
            raise RuntimeError("libvcg_b200.so ABI version mismatch")
        _lib = lib
    return _lib


class VcgError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = load().vcg_last_error().decode(errors="replace")
        raise VcgError(f"{what} failed (code {rc}): {msg}")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def dtype_code(dt):
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise VcgError(f"unsupported dtype {dt}")


def launch_count():
    return int(load().vcg_launch_count())
