"""ctypes binding of the C ABI declared in include/ivid_b200.h.

The product path has NO CPU fallback: if libivid_b200.so is missing or fails to load this module raises, and every
entry point converts a non-zero status into the Python exception type the reference would have raised
(AssertionError for `assert`s, NotImplementedError, RuntimeError).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_int64, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libivid_b200.so")

IVID_OK = 0
IVID_ERR_INVALID_ARGUMENT = 1
IVID_ERR_NOT_IMPLEMENTED = 2
IVID_ERR_CUDA = 3
IVID_ERR_STATE = 4


class CondT(Structure):
    _fields_ = [
        ("kind", c_int),
        ("y_dev", c_void_p),
        ("mask_dev", c_void_p),
        ("mask_rgb_dev", c_void_p),
        ("noise_dev", c_void_p),
        ("seed", c_uint64),
        ("stream_id", c_uint32),
    ]


class StepArgsT(Structure):
    _fields_ = [
        ("kind", c_int),
        ("use_cfg", c_int),
        ("strength", c_float),
        ("clip_denoised", c_int),
        ("eta", c_float),
        ("classes_dev", c_void_p),
        ("cond", CondT),
        ("replace_rgb_dev", c_void_p),
        ("replace_rgb_mask_dev", c_void_p),
        ("replace_rgb_weight", c_double),
        ("replace_depth_dev", c_void_p),
        ("replace_depth_mask_dev", c_void_p),
        ("replace_depth_weight", c_double),
        ("constrain_depth_dev", c_void_p),
        ("constrain_depth_weight", c_double),
        ("step_noise_dev", c_void_p),
        ("seed", c_uint64),
    ]


class WarpParamsT(Structure):
    _fields_ = [("fov_deg", c_double), ("near", c_double), ("far", c_double), ("atol", c_double), ("rtol", c_double),
                ("erode_rgb", c_int), ("padding", c_double)]


# name -> (restype, argtypes); also the list tests/test_abi.py checks against the header
SIGNATURES = {
    "ivid_last_error": (c_char_p, []),
    "ivid_version": (c_int, []),
    "ivid_device_info": (c_int, [c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "ivid_unet_create": (c_int, [c_char_p, POINTER(c_void_p)]),
    "ivid_unet_destroy": (c_int, [c_void_p]),
    "ivid_unet_num_params": (c_int, [c_void_p, POINTER(c_int)]),
    "ivid_unet_param_info": (c_int, [c_void_p, c_int, POINTER(c_char_p), POINTER(c_int64), POINTER(c_int), POINTER(c_int)]),
    "ivid_unet_set_param": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int]),
    "ivid_unet_finalize": (c_int, [c_void_p, c_int]),
    "ivid_unet_weight_arena": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_uint64)]),
    "ivid_unet_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "ivid_unet_forward_cond": (c_int, [c_void_p, c_void_p, c_int, POINTER(CondT), c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "ivid_unet_debug_tap": (c_int, [c_void_p, c_int, c_char_p, c_void_p, c_uint64, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "ivid_unet_profile_begin": (c_int, [c_void_p]),
    "ivid_unet_profile_end": (c_int, [c_void_p, c_char_p, c_int]),
    "ivid_sampler_create": (c_int, [POINTER(c_double), c_int, POINTER(c_void_p)]),
    "ivid_sampler_destroy": (c_int, [c_void_p]),
    "ivid_sampler_table": (c_int, [c_void_p, c_int, POINTER(c_double), c_int]),
    "ivid_sampler_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, POINTER(StepArgsT), c_void_p]),
    "ivid_sampler_step_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, POINTER(StepArgsT), c_void_p]),
    "ivid_cfg_mix": (c_int, [c_void_p, c_float, c_void_p, c_uint64, c_void_p]),
    "ivid_sampler_run": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, POINTER(StepArgsT), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ivid_op_conv2d": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "ivid_op_group_norm": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                   c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ivid_op_attention": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ivid_warp_create": (c_int, [c_int, c_int, c_int, c_int, c_double, c_double, c_int, POINTER(c_void_p)]),
    "ivid_warp_destroy": (c_int, [c_void_p]),
    "ivid_warp_reset": (c_int, [c_void_p]),
    "ivid_warp_num_views": (c_int, [c_void_p, POINTER(c_int)]),
    "ivid_warp_add_view": (c_int, [c_void_p, c_void_p, c_void_p, c_int, POINTER(WarpParamsT), c_void_p]),
    "ivid_warp_mesh_from_depth": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(WarpParamsT), c_void_p, c_void_p, c_void_p]),
    "ivid_warp_set_mesh": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ivid_warp_get_mesh": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ivid_warp_render": (c_int, [c_void_p, c_void_p, c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ivid_warp_aggregate": (c_int, [c_void_p, c_void_p, c_int, POINTER(WarpParamsT), c_void_p, c_void_p]),
    "ivid_warp_resolve_frame": (c_int, [c_void_p, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ivid_warp_render_simple": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ivid_warp_forward_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, POINTER(WarpParamsT), c_void_p, c_void_p]),
    "ivid_warp_postfilter": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(WarpParamsT), c_void_p, c_void_p]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load (once) and return the native library; raises if it is missing — there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m ivid_b200.build` (ivid_b200 has no CPU fallback)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(l, name)
            except AttributeError:
                continue   # symbol check is done by tests/test_abi.py; optional groups may be absent in old builds
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def last_error() -> str:
    m = lib().ivid_last_error()
    return m.decode() if m else ""


def check(status: int) -> None:
    if status == IVID_OK:
        return
    msg = last_error()
    if status == IVID_ERR_INVALID_ARGUMENT:
        raise AssertionError(msg)
    if status == IVID_ERR_NOT_IMPLEMENTED:
        raise NotImplementedError(msg)
    raise RuntimeError(f"ivid_b200 native error {status}: {msg}")


def ptr(t) -> c_void_p:
    """Raw data pointer of a torch tensor (or None)."""
    if t is None:
        return c_void_p(None)
    return c_void_p(t.data_ptr())


def cur_stream(device=None) -> c_void_p:
    import torch
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)
