"""ctypes binding of libmip360_b200.so (the C ABI declared in include/mip360_b200.h).

The library is the only implementation of the hot path: there is no CPU fallback.  Importing this
module never builds anything; `load()` raises if the shared object is missing (run
`python -m mipnerf360_b200.build` or `__graft_entry__.build()` first).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_longlong, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmip360_b200.so")

P = c_void_p  # every tensor argument is passed as a raw device pointer

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/mip360_b200.h one to one
SIGNATURES = {
    "mip360_last_error": [],
    "mip360_version": [],
    "mip360_launch_count": [],
    "mip360_reset_launch_count": [],
    "mip360_sm_count": [],
    "mip360_set_option": [c_int, c_int],
    "mip360_level0_t_vals": [P, P, P, P, P, c_int, c_int, P],
    "mip360_level0_sample": [P, P, P, P, c_int, ctypes.c_ulonglong, ctypes.c_uint, P, P, P, P, c_int, c_int, P],
    "mip360_frustum_norm_sq": [P, P, c_int, P, c_int, c_int, P, P],
    "mip360_cast_ipe": [P, P, c_int, P, P, P, P, P, c_int, c_int, c_int, c_int, P, P, P, P, P],
    "mip360_cast_ipe_x": [P, P, c_int, P, P, P, c_int, P, P, c_int, c_int, c_int, c_int, P, P, P, P, c_int, P],
    "mip360_gaussian_to_xyz": [P, P, P, P, c_int, c_int, P, P, P],
    "mip360_gaussian_to_xyz_diag": [P, P, P, P, c_int, c_int, P, P, P],
    "mip360_sum_sq": [P, c_longlong, P, P],
    "mip360_contract": [P, c_longlong, P, P, P],
    "mip360_gaussian_contract": [P, P, P, c_longlong, P, P, P],
    "mip360_ipe": [P, P, c_longlong, P, P],
    "mip360_viewdir_enc": [P, c_int, c_int, c_int, P, P],
    "mip360_blur_weights": [P, c_int, c_int, c_float, P, P],
    "mip360_resample_cdf": [P, c_int, c_int, P, P],
    "mip360_resample_invert": [P, P, P, c_int, c_int, c_int, c_int, P, P, P],
    "mip360_resample": [P, P, P, P, c_int, c_int, c_float, c_int, P, P],
    "mip360_resample_sample": [P, P, P, P, c_int, ctypes.c_ulonglong, ctypes.c_uint, P, c_float, P, P, c_int, c_int, c_float,
                               c_int, P, P],
    "mip360_composite_fwd": [P, P, P, P, c_int, c_int, c_int, c_float, c_float, c_int, P, P, P, P, P],
    "mip360_composite_fwd_s": [P, P, P, P, c_int, c_int, c_int, c_float, c_float, c_int, P, P, P, P, P, P, P, P, P, P],
    "mip360_composite_bwd": [P, P, P, P, c_int, c_int, c_int, c_float, c_float, c_int, P, P, P, P, P, P, P, P, P],
    "mip360_density_to_weight_fwd": [P, P, P, c_int, c_int, c_int, c_float, P, P],
    "mip360_density_to_weight_bwd": [P, P, P, c_int, c_int, c_int, c_float, P, P, P],
    "mip360_head_grad_pack": [P, P, c_longlong, c_int, c_int, P, P],
    "mip360_t_to_s": [P, P, P, c_int, c_int, P, P, P],
    "mip360_s_to_t": [P, P, P, c_int, c_int, P, P],
    "mip360_partials_len": [c_int],
    "mip360_distortion_fwd": [P, P, c_int, c_int, P, P, P, P],
    "mip360_distortion_bwd": [P, P, c_int, c_int, P, P, P],
    "mip360_bounds_per_ray": [P, P, P, c_int, c_int, P, P],
    "mip360_bounds": [P, P, P, c_int, c_int, P, P, P],
    "mip360_bounds_reduce": [P, c_int, c_int, P, P],
    "mip360_interlevel_fwd": [P, P, P, c_int, c_int, c_int, c_float, P, P, P],
    "mip360_interlevel_bwd": [P, P, P, c_int, c_int, c_int, c_float, P, P, P],
    "mip360_linear_fwd": [P, P, P, c_int, c_int, c_int, c_int, P, P, c_int, P],
    "mip360_linear_fwd_head": [P, P, P, c_int, c_int, c_int, c_int, P, P, P, P],
    "mip360_head_bwd": [P, P, P, c_int, c_int, c_int, P, P, c_int, P, P],
    "mip360_linear_dgrad": [P, P, P, c_int, c_int, c_int, c_int, P, P],
    "mip360_linear_wgrad": [P, P, c_int, c_int, c_int, P, P, P],
    "mip360_cast_weight": [P, c_int, c_int, c_int, c_int, P, P, P],
    "mip360_mlp_fwd": [P, c_int, P, c_int, P, c_int, P, c_int, P, P],
    "mip360_mlp_fwd_fused_narrow": [P, c_int, P, c_int, P, c_int, P, P, P],
    "mip360_mlp_fwd_fused_head": [P, c_int, P, c_int, P, P, c_int, P, P],
    "mip360_mlp_bwd_fused_head": [P, P, c_int, P, c_int, P, P, P, P, P, P, P],
    "mip360_mlp_bwd": [P, P, P, c_int, P, c_int, P, c_int, P, P, P, P, P, P, P],
    "mip360_generate_rays": [P, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_int, c_float, P, P, P, P, P, P, P],
    "mip360_generate_rays_range": [P, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_int, c_float, c_longlong,
                                   c_longlong, P, P, P, P, P, P, P],
    "mip360_to8b": [P, c_longlong, P, P],
    "mip360_vis_partials_len": [],
    "mip360_vis_work_len": [],
    "mip360_normals_scaling": [P, c_int, c_int, P, P, P],
    "mip360_visualize_normals": [P, P, P, c_int, c_int, P, P, P],
    "mip360_depth_range": [P, P, c_longlong, c_double, c_float, c_float, c_int, c_int, P, P, P],
    "mip360_visualize_depth": [P, P, P, c_int, c_float, P, c_int, c_longlong, P, P, P],
    "mip360_adamw": [P, P, P, P, c_longlong, c_float, c_float, c_float, c_float, c_float, c_int, P],
    "mip360_adamw_pack": [P, c_int, c_int, P, P, P, P, c_float, c_float, c_float, c_float, c_float, c_int, P, c_int, c_int, P],
}
_RESTYPES = {
    "mip360_last_error": c_char_p,
    "mip360_launch_count": c_longlong,
    "mip360_reset_launch_count": None,
}



class Layer(ctypes.Structure):
    """struct mip360_layer of include/mip360_b200.h."""
    _fields_ = [("W", c_void_p), ("Wt", c_void_p), ("bias", c_void_p), ("n_pad", c_int), ("k_pad", c_int), ("act", c_int)]


class PackEntry(ctypes.Structure):
    """struct mip360_pack_entry of include/mip360_b200.h."""
    _fields_ = [("w_src", c_void_p * 2), ("b_src", c_void_p * 2), ("rows", c_int * 2), ("K", c_int), ("n_pad", c_int),
                ("k_pad", c_int), ("tile_begin", c_int), ("Wb", c_void_p), ("Wt", c_void_p), ("bias", c_void_p),
                ("w4", c_void_p)]


_lib = None


class Mip360Error(RuntimeError):
    pass


def load():
    """dlopen the CUDA library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Mip360Error(
                f"{LIB_PATH} is missing: the sm_100a kernel library has not been built. "
                "Run `python -m mipnerf360_b200.build`; there is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header and library out of sync
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        _lib = lib
    return _lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


# When set to a list, every call is bracketed by CUDA events on the launching stream and
# (name, int args, start, end) is appended: bench.py reads per-kernel durations from it.
PROFILE = None


def call(name, *args):
    lib = load()
    if PROFILE is None:
        rc = getattr(lib, name)(*args, stream())
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args, stream())
        e1.record()
        PROFILE.append((name, tuple(a for a in args if isinstance(a, int) and not isinstance(a, bool) and a < (1 << 31)),
                        e0, e1))
    if rc != 0:
        raise Mip360Error(f"{name} failed ({rc}): {lib.mip360_last_error().decode()}")


ERR_UNSUPPORTED = -3


def call_rc(name, *args):
    """Like call(), but a MIP360_ERR_UNSUPPORTED result is returned (True = ran, False = shape not covered by this entry
    point) instead of raised: for optional fast paths with a general fallback."""
    lib = load()
    if PROFILE is None:
        rc = getattr(lib, name)(*args, stream())
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args, stream())
        e1.record()
        if rc == 0:
            PROFILE.append((name, tuple(a for a in args if isinstance(a, int) and not isinstance(a, bool) and a < (1 << 31)),
                            e0, e1))
    if rc == ERR_UNSUPPORTED:
        return False
    if rc != 0:
        raise Mip360Error(f"{name} failed ({rc}): {lib.mip360_last_error().decode()}")
    return True


def check_cuda(*tensors, dtype=torch.float32):
    """The product path runs on the GPU only: reject CPU tensors loudly instead of falling back."""
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise Mip360Error("mipnerf360_b200 ops need CUDA tensors (sm_100a); there is no CPU fallback")
        if dtype is not None and t.dtype != dtype:
            raise Mip360Error(f"expected {dtype}, got {t.dtype}")


def f32c(t):
    """contiguous fp32 view/copy of a borrowed tensor (callers pass views such as t_vals[..., :-1])."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def launch_count():
    return int(load().mip360_launch_count())


def reset_launch_count():
    load().mip360_reset_launch_count()


OPT_RAY_GROUP, OPT_CTA_PAIR, OPT_SHORT_K, OPT_PACKED_EPILOGUE, OPT_FUSED_NARROW = 0, 1, 2, 3, 4


def set_option(key, value):
    """Switch a kernel variant on/off (all variants compute the same function; see mip360_set_option)."""
    if load().mip360_set_option(int(key), int(bool(value))) != 0:
        raise Mip360Error(load().mip360_last_error().decode())


def sm_count():
    return int(load().mip360_sm_count())
