"""ctypes binding of libsrcgan_b200.so (C ABI declared in include/srcgan_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is
raised.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C srcgan_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsrcgan_b200.so")

DT_F32, DT_BF16 = 0, 1
WL_RSCK, WL_RSKC, WL_TC, WL_TC_S2, WL_TC_DGRAD_S2 = 0, 1, 2, 3, 4
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC = 0, 1, 2


class ConvParams(C.Structure):
    """Mirror of ``srcgan_conv_params`` (include/srcgan_b200.h)."""
    _fields_ = [
        ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("cin", C.c_int32), ("cout", C.c_int32),
        ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
        ("upsample", C.c_int32),
        ("ho", C.c_int32), ("wo", C.c_int32),
        ("dtype", C.c_int32), ("engine", C.c_int32),
        ("x", C.c_void_p), ("x_ld", C.c_int32),
        ("wgt", C.c_void_p),
        ("bias", C.c_void_p),
        ("y", C.c_void_p), ("y_ld", C.c_int32),
        ("act", C.c_int32), ("act_slope", C.c_float), ("alpha", C.c_float),
        ("r1", C.c_void_p), ("r1_ld", C.c_int32), ("beta1", C.c_float),
        ("r2", C.c_void_p), ("r2_ld", C.c_int32), ("beta2", C.c_float),
        ("mask", C.c_void_p), ("mask_ld", C.c_int32), ("mask_slope", C.c_float),
        ("signbits", C.c_void_p), ("maskbits", C.c_void_p),
        ("zero_row_period", C.c_int32), ("flags", C.c_int32),
        ("x_group_stride", C.c_int64),
    ]


class PackBlock(C.Structure):
    """Mirror of ``srcgan_pack_block`` (include/srcgan_b200.h): one source block of a batched weight re-pack."""
    _fields_ = [
        ("src", C.c_void_p), ("out", C.c_void_p),
        ("nstride", C.c_int64), ("kstride", C.c_int64), ("src_off", C.c_int64), ("elem0", C.c_int64),
        ("n_count", C.c_int32), ("k0", C.c_int32), ("k_len", C.c_int32),
        ("bn", C.c_int32), ("nchunks", C.c_int32), ("total_slots", C.c_int32),
        ("taps", C.c_int32), ("flip", C.c_int32), ("scale", C.c_float),
        ("slot_off", C.c_int32 * 16),
    ]


_P = C.c_void_p
_I = C.c_int
_L = C.c_int64
_F = C.c_float
_Z = C.c_size_t

# name -> (restype, argtypes); every symbol include/srcgan_b200.h declares
SIGNATURES = {
    "srcgan_version": (C.c_char_p, []),
    "srcgan_last_error": (C.c_char_p, []),
    "srcgan_launch_count": (_L, []),
    "srcgan_last_kernel": (C.c_char_p, []),
    "srcgan_pack_weights": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "srcgan_packed_weight_bytes": (_Z, [_I, _I, _I, _I, _I, _I]),
    "srcgan_pack_slots": (_I, [_I, _I, _I, _I, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "srcgan_pack_weights_batch": (_I, [_P, _I, _L, _P]),
    "srcgan_conv_fprop": (_I, [C.POINTER(ConvParams), _P]),
    "srcgan_conv_fprop_pair": (_I, [C.POINTER(ConvParams), C.POINTER(ConvParams), _P]),
    "srcgan_conv_fprop_pair_supported": (_I, [C.POINTER(ConvParams), C.POINTER(ConvParams)]),
    "srcgan_conv_dgrad": (_I, [C.POINTER(ConvParams), _P]),
    "srcgan_conv_wgrad_workspace_bytes": (_Z, [C.POINTER(ConvParams)]),
    "srcgan_conv_wgrad": (_I, [C.POINTER(ConvParams), _P, _P, _I, _P, _Z, _P]),
    "srcgan_conv_wgrad_split": (_I, [C.POINTER(ConvParams), _P, _I, _I, _P, _P, _I, _I, _P, _I, _I, _P, _Z, _P]),
    "srcgan_nchw_to_nhwc": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _P]),
    "srcgan_nhwc_to_nchw": (_I, [_P, _I, _I, _P, _I, _I, _I, _I, _P]),
    "srcgan_act_backward": (_I, [_P, _I, _P, _I, _P, _I, _L, _I, _F, _I, _P]),
    "srcgan_bias_act": (_I, [_P, _I, _L, _I, _P, _I, _F, _I, _P]),
    "srcgan_add": (_I, [_P, _I, _P, _I, _P, _I, _L, _I, _I, _P]),
    "srcgan_colsum_workspace_bytes": (_Z, [_L, _I]),
    "srcgan_colsum": (_I, [_P, _I, _I, _L, _I, _P, _F, _I, _P, _Z, _P]),
    "srcgan_depth_to_space": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _I, _P]),
    "srcgan_space_to_depth": (_I, [_P, _I, _P, _I, _P, _I, _F, _I, _I, _I, _I, _I, _P]),
    "srcgan_pixel_shuffle": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "srcgan_upsample2x": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _I, _P]),
    "srcgan_upsample2x_adjoint": (_I, [_P, _I, _P, _I, _P, _I, _F, _I, _I, _I, _I, _I, _P]),
    "srcgan_bn_workspace_bytes": (_Z, [_L, _I]),
    "srcgan_bn_forward": (_I, [_P, _I, _P, _I, _L, _I, _I, _P, _P, _P, _P, _P, _P, _I, _F, _F, _F, _P, _Z, _P]),
    "srcgan_bn_backward": (_I, [_P, _I, _P, _I, _P, _I, _P, _I, _L, _I, _I, _P, _P, _P, _P, _F, _I, _P, _P, _I, _P, _Z, _P]),
    "srcgan_gn_workspace_bytes": (_Z, [_I, _I]),
    "srcgan_gn_forward": (_I, [_P, _I, _P, _I, _I, _L, _I, _I, _I, _P, _P, _P, _P, _F, _P, _I, _I, _F, _P, _Z, _P]),
    "srcgan_gn_backward": (_I, [_P, _I, _P, _I, _P, _I, _I, _L, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P, _Z, _P]),
    "srcgan_loss_workspace_bytes": (_Z, [_L]),
    "srcgan_loss_fwd_bwd": (_I, [_I, _P, _P, _F, _L, _P, _P, _P, _Z, _P]),
    "srcgan_metrics_sqerr": (_I, [_P, _P, _L, _P, _P, _Z, _P]),
    "srcgan_metrics_ae": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "srcgan_ssim_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "srcgan_ssim": (_I, [_P, _P, _I, _I, _I, _I, _F, _P, _P, _Z, _P]),
    "srcgan_minmax": (_I, [_P, _L, _P, _P]),
    "srcgan_eval_metrics_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "srcgan_eval_metrics": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _Z, _P]),
    "srcgan_ssim_backward_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "srcgan_ssim_backward": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _F, _P, _P, _Z, _P]),
    "srcgan_rgb2lab": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "srcgan_lab2rgb": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "srcgan_rgb2lab_u8": (_I, [_P, _P, _I, _I, _I, _P]),
    "srcgan_lab2rgb_u8": (_I, [_P, _P, _I, _I, _I, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the shared library once; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "srcgan_b200: %s is missing - the CUDA extension has not been built "
            "(run __graft_entry__.build() or `make -C srcgan_b200/csrc`). There is no CPU fallback." % LIB_PATH)
    try:
        import torch  # noqa: F401  (makes sure libcudart is already mapped)
    except Exception:
        pass
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().srcgan_last_error().decode("utf-8", "replace")
        raise RuntimeError("srcgan_b200 %s failed (code %d): %s" % (what, rc, msg))


def launch_count() -> int:
    return int(load().srcgan_launch_count())


def last_kernel() -> str:
    """Name of the kernel this thread's most recent library call launched last."""
    return load().srcgan_last_kernel().decode()
