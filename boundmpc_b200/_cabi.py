"""ctypes binding of include/boundmpc_b200.h.

The CUDA library is the only compute path of this package: if it cannot be loaded, or no
CUDA device is usable, every entry point raises — there is no CPU fallback.
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BMPC_LIB") or os.path.join(_HERE, "libboundmpc_b200.so")   # BMPC_LIB: development builds
_lib = None

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)


class BmpcConfig(ctypes.Structure):
    """struct bmpc_config (include/boundmpc_b200.h)"""
    _fields_ = [("N", ctypes.c_int32), ("nr_segs", ctypes.c_int32), ("dt", ctypes.c_double),
                ("u_min", ctypes.c_double), ("u_max", ctypes.c_double),
                ("ut_min", ctypes.c_double), ("ut_max", ctypes.c_double),
                ("q_lim_lower", ctypes.c_double * 7), ("q_lim_upper", ctypes.c_double * 7),
                ("dq_lim_lower", ctypes.c_double * 7), ("dq_lim_upper", ctypes.c_double * 7),
                ("tol", ctypes.c_double), ("max_iter", ctypes.c_int32), ("mu_init", ctypes.c_double),
                ("bound_push", ctypes.c_double), ("device", ctypes.c_int32), ("threads", ctypes.c_int32)]


EXPORTS = ["bmpc_create", "bmpc_destroy", "bmpc_dims", "bmpc_bounds", "bmpc_workspace_bytes", "bmpc_solve_batch",
           "bmpc_solve_batch_host", "bmpc_prepare_batch", "bmpc_prepare_batch_host", "bmpc_post_batch", "bmpc_post_batch_host", "bmpc_post_log_batch", "bmpc_update_batch", "bmpc_finish_batch", "bmpc_mpc_step_batch_host", "bmpc_eval_batch_host", "bmpc_kkt_step_batch_host", "bmpc_launch_count", "bmpc_launch_shape", "bmpc_fp64_peak", "bmpc_last_error"]


class BmpcError(RuntimeError):
    pass


def lib():
    """Load the CUDA library; raise loudly if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BmpcError(f"{LIB_PATH} not found: build it with `python -m boundmpc_b200.build` "
                        "(boundmpc_b200 has no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_size_t
    L.bmpc_create.argtypes = [ctypes.POINTER(BmpcConfig), ctypes.POINTER(vp)]
    L.bmpc_create.restype = ctypes.c_int
    L.bmpc_destroy.argtypes = [vp]
    L.bmpc_destroy.restype = None
    L.bmpc_dims.argtypes = [vp, c_int32_p, c_int32_p, c_int32_p]
    L.bmpc_bounds.argtypes = [vp, c_double_p, c_double_p, c_double_p, c_double_p]
    L.bmpc_workspace_bytes.argtypes = [vp, i32, ctypes.POINTER(sz)]
    L.bmpc_solve_batch.argtypes = [vp, i32] + [vp] * 12
    L.bmpc_solve_batch_host.argtypes = [vp, i32] + [vp] * 10
    L.bmpc_eval_batch_host.argtypes = [vp, i32] + [vp] * 9
    L.bmpc_kkt_step_batch_host.argtypes = [vp, i32] + [vp] * 7
    L.bmpc_prepare_batch.argtypes = [vp, i32, vp, i32, i32] + [vp] * 7
    L.bmpc_prepare_batch_host.argtypes = [vp, i32, vp, i32, i32] + [vp] * 6
    L.bmpc_post_batch.argtypes = [vp, i32, vp, i32, i32] + [vp] * 8
    L.bmpc_post_batch_host.argtypes = [vp, i32, vp, i32, i32] + [vp] * 7
    L.bmpc_post_log_batch.argtypes = [vp, i32, vp, i32, i32] + [vp] * 11
    L.bmpc_update_batch.argtypes = [vp, i32, vp, i32, i32] + [vp] * 7
    L.bmpc_finish_batch.argtypes = [vp, i32, vp, i32, i32] + [vp] * 10 + [i32, vp]
    L.bmpc_mpc_step_batch_host.argtypes = [vp, i32, vp, i32, i32] + [vp] * 12
    L.bmpc_launch_count.argtypes = [vp]
    L.bmpc_launch_count.restype = ctypes.c_int64
    L.bmpc_launch_shape.argtypes = [vp, c_int32_p, c_int32_p, c_int32_p, c_int32_p]
    L.bmpc_fp64_peak.argtypes = [vp, i32, c_double_p]
    L.bmpc_last_error.restype = ctypes.c_char_p
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        raise BmpcError(f"{what} failed ({rc}): {lib().bmpc_last_error().decode()}")


def ptr(a):
    """Host pointer of a C-contiguous float64/int32 numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(a.ctypes.data)
