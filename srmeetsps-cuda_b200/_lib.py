"""ctypes binding of include/srps_c_api.h.  Fails loudly when the CUDA library is missing:
there is no CPU or PyTorch fallback for any operator."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SRPS_LIB") or os.path.join(HERE, "libsrps_b200.so")     # SRPS_LIB: A/B builds of the same ABI

SRPS_ALBEDO_CLOSED_FORM = 0
SRPS_ALBEDO_REFERENCE_CG = 1
BUF_S, BUF_RHO, BUF_Z, BUF_N, BUF_DZ, BUF_Z0S = range(6)
BUF_W, BUF_G, BUF_E0, BUF_R = 16, 17, 18, 19


class Problem(C.Structure):
    _fields_ = [("h", C.c_int), ("w", C.c_int), ("n_images", C.c_int), ("n_channels", C.c_int), ("sf", C.c_int),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("mask", C.c_void_p), ("device", C.c_int), ("albedo_mode", C.c_int),
                ("cg_max_iter", C.c_int), ("cg_tol", C.c_float),
                ("strip_j0", C.c_int), ("strip_j1", C.c_int), ("rank", C.c_int), ("world", C.c_int)]


class Timings(C.Structure):
    _fields_ = [("ms_lighting", C.c_float), ("ms_albedo", C.c_float), ("ms_depth", C.c_float),
                ("ms_normals", C.c_float), ("ms_total", C.c_float), ("ms_depth_cg", C.c_float),
                ("cg_iters", C.c_int), ("albedo_cg_iters", C.c_int * 3), ("launches", C.c_longlong),
                ("cg_deferred", C.c_int), ("cg_zskip", C.c_int)]


EXPORTS = {
    "srps_ctx_create": (C.c_int, [C.POINTER(Problem), C.POINTER(C.c_void_p)]),
    "srps_ctx_destroy": (None, [C.c_void_p]),
    "srps_last_error": (C.c_char_p, [C.c_void_p]),
    "srps_npix": (C.c_int, [C.c_void_p]),
    "srps_npixs": (C.c_int, [C.c_void_p]),
    "srps_upload_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "srps_upload_images_u8": (C.c_int, [C.c_void_p, C.c_void_p]),
    "srps_upload_images_u8_strided": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong]),
    "srps_set_state": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "srps_download": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "srps_lighting": (C.c_int, [C.c_void_p]),
    "srps_albedo": (C.c_int, [C.c_void_p]),
    "srps_depth": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "srps_normals": (C.c_int, [C.c_void_p]),
    "srps_outer_iteration": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "srps_run": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "srps_get_timings": (C.c_int, [C.c_void_p, C.POINTER(Timings)]),
    "srps_synchronize": (C.c_int, [C.c_void_p]),
    "srps_dist_blob_size": (C.c_int, []),
    "srps_dist_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "srps_dist_connect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "srps_pixel_range": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong),
                                   C.POINTER(C.c_longlong)]),
    "srps_upload_state_strided": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
    "srps_timer_start": (C.c_int, [C.c_void_p]),
    "srps_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "srps_profile_kernels": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "srps_apply_depth_operator": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "srps_dev_lighting": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "srps_dev_albedo": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "srps_dev_depth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "srps_dev_normals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "srps_init_depth_mean": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "srps_init_depth_smooth_upsample": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                                  C.c_void_p, C.c_void_p]),
    "srps_build_info": (C.c_char_p, []),
}

_lib = None


def load():
    """Load libsrps_b200.so and bind every symbol include/srps_c_api.h declares."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m srmeetsps_cuda_b200.build` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
