"""ctypes binding of libscn_b200.so (include/scn_b200.h).  There is no fallback: if the CUDA library is
missing or an entry fails, the caller gets an exception -- never a CPU or PyTorch substitute."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libscn_b200.so")

FP32, TF32, BF16 = 0, 1, 2

_i64p = C.POINTER(C.c_int64)
_vp = C.c_void_p

# name -> (restype, argtypes); kept in one table so tests can check it against include/scn_b200.h
PROTOTYPES = {
    "scn_version": (C.c_int, []),
    "scn_last_error": (C.c_char_p, []),
    "scn_launch_count": (C.c_int64, []),
    "scn_profile": (None, [C.c_int]),
    "scn_profile_read": (C.c_int, [C.POINTER(C.c_double), C.c_int]),
    "scn_profile_kind_name": (C.c_char_p, [C.c_int]),
    "scn_meta_create": (_vp, [C.c_int]),
    "scn_meta_destroy": (None, [_vp]),
    "scn_pool_trim": (C.c_int, [C.c_int, C.c_int64]),
    "scn_tile_sort": (C.c_int, [C.c_int]),
    "scn_deterministic": (C.c_int, [C.c_int]),
    "scn_input_normals": (C.c_int, [_vp, _vp, C.c_int]),
    "scn_guided": (C.c_int, [_vp, _i64p]),
    "scn_subm_guided_table": (C.c_int, [_vp, _i64p, _vp, _vp, _vp]),
    "scn_normals": (C.c_int, [_vp, _i64p, _vp, _vp]),
    "scn_input_layer_build": (C.c_int, [_vp, _i64p, _vp, C.c_int, C.c_int64, C.c_int, C.c_int, _vp, _i64p]),
    "scn_input_layer_fwd": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "scn_input_layer_bwd": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "scn_output_layer_fwd": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "scn_output_layer_bwd": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "scn_float_coords": (C.c_int, [_vp, C.c_int64, C.POINTER(C.c_float), C.c_int, C.c_float, _vp, _vp, _vp]),
    "scn_n_points": (C.c_int64, [_vp]),
    "scn_nactive": (C.c_int64, [_vp, _i64p]),
    "scn_spatial_locations": (C.c_int, [_vp, _i64p, _vp]),
    "scn_resolution_scatter": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, C.c_int, _vp, _vp]),
    "scn_subm_dilation": (C.c_int, [_vp, C.c_int]),
    "scn_subm_rulebook": (C.c_int, [_vp, _i64p, _vp, _i64p]),
    "scn_subm_neighbour_table": (C.c_int, [_vp, _i64p, _vp]),
    "scn_strided_rulebook": (C.c_int, [_vp, _i64p, _i64p, _vp, _i64p]),
    "scn_strided_table": (C.c_int, [_vp, _i64p, _vp, _vp]),
    "scn_subm_fwd": (C.c_int, [_vp, _i64p, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.POINTER(C.c_double)]),
    "scn_fuses_residual": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "scn_subm_fwd_bn": (C.c_int, [_vp, _i64p, _vp, _vp, _vp, _vp, _vp, _vp, C.c_float, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp,
                                  C.POINTER(C.c_double)]),
    "scn_bn_eval_coeffs": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_float, _vp, _vp, _vp]),
    "scn_subm_bwd": (C.c_int, [_vp, _i64p, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "scn_conv_fwd": (C.c_int, [_vp, _i64p, _i64p, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.POINTER(C.c_double)]),
    "scn_conv_bwd": (C.c_int, [_vp, _i64p, _i64p, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "scn_deconv_fwd": (C.c_int, [_vp, _i64p, _i64p, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.POINTER(C.c_double)]),
    "scn_deconv_bwd": (C.c_int, [_vp, _i64p, _i64p, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "scn_bn_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_int, C.c_float, _vp]),
    "scn_bn_bwd_fusion": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_float, _vp]),
    "scn_bn_bwd_fusable": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "scn_bn_bwd_apply": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int64, C.c_int, _vp]),
    "scn_grad_bf16": (C.c_int, [_vp, _vp, _vp]),
    "scn_grad_stride": (C.c_int, [_vp, C.c_int64]),
    "scn_out_stats": (C.c_int, [_vp, _vp]),
    "scn_bf16_operand": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "scn_bf16_plan": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "scn_bn_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int, C.c_float, _vp]),
}

_lib = None


class ScnError(RuntimeError):
    pass


def lib():
    """Load (once) and return the CDLL.  Raises if the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ScnError(
                f"{LIB_PATH} not found: build it with `python occuseg_b200/csrc/build.py` "
                "(there is no CPU fallback for this path)")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(status):
    if status != 0:
        raise ScnError(lib().scn_last_error().decode("utf-8", "replace"))


def size3(v):
    """int / sequence / LongTensor -> ctypes int64[3]"""
    if hasattr(v, "tolist"):
        v = v.tolist()
    if isinstance(v, (int,)):
        v = [v, v, v]
    v = [int(x) for x in v]
    assert len(v) == 3, "only 3-D grids are on this path"
    return (C.c_int64 * 3)(*v)


def trim_memory(device=None, keep_bytes=0):
    """Return the free part of the library's cached scratch memory to the driver (see scn_pool_trim)."""
    if device is None:
        import torch
        device = torch.cuda.current_device()
    check(lib().scn_pool_trim(int(device), int(keep_bytes)))


def tile_sort(block_rows):
    """Rows per pattern-sort block of the tensor-core tile order (0 = natural order); returns the previous setting."""
    return int(lib().scn_tile_sort(int(block_rows)))


def deterministic(on):
    """Deterministic mode of the library (scn_deterministic): reductions merged in a fixed order instead of with floating-point
    atomics; returns the previous setting."""
    return bool(lib().scn_deterministic(1 if on else 0))


def launch_count():
    return int(lib().scn_launch_count())


def profile(enable):
    lib().scn_profile(int(bool(enable)))


def profile_read():
    """{kind: dict(launches, ms, bytes, flops)} accumulated since profile(True)."""
    buf = (C.c_double * 64)()
    n = lib().scn_profile_read(buf, 16)
    out = {}
    for k in range(n):
        name = lib().scn_profile_kind_name(k).decode()
        out[name] = dict(launches=int(buf[4 * k]), ms=buf[4 * k + 1], bytes=buf[4 * k + 2], flops=buf[4 * k + 3])
    return out
