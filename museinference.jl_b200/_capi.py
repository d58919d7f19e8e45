"""ctypes binding of include/muse_b200.h — the only way Python reaches the kernels.

The library must exist (``__graft_entry__.build()`` / ``_build.build_library()``); there is no
fallback: a missing or unloadable ``libmuse_b200.so`` raises ``MuseBackendError``.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

ABI_VERSION = 1

FAMILY_FUNNEL, FAMILY_HIERGAUSS, FAMILY_CORRGAUSS, FAMILY_TWOLAYER = 1, 2, 3, 4
START_ZEROS, START_PREV, START_TRUTH, START_USER = 0, 1, 2, 3
STATUS_G_CONVERGED, STATUS_XF_CONVERGED, STATUS_MAXITER, STATUS_LS_FAILED, STATUS_NONFINITE = 0, 1, 2, 3, 4
E_NAMES = {0: "OK", -1: "EINVAL", -2: "ENODEVICE", -3: "ECUDA", -4: "ENOMEM", -5: "EUNSUPPORTED", -6: "ESTATE"}

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class MuseBackendError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libmuse_b200: {E_NAMES.get(code, code)}: {message}")
        self.code = code


class muse_cfg(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("family", C.c_int32), ("d", C.c_int32), ("ntheta", C.c_int32),
        ("nsims", C.c_int32), ("device", C.c_int32), ("sim_offset", C.c_int64),
        ("nsims_h", C.c_int32), ("kernel", C.c_int32), ("h_sim_offset", C.c_int64),
        ("lbfgs_m", C.c_int32), ("max_iters", C.c_int32), ("group", C.c_int32), ("cluster", C.c_int32),
        ("P", c_double_p), ("L", c_double_p), ("stream", C.c_void_p),
    ]


class muse_profile(C.Structure):
    _fields_ = [
        ("launches", C.c_int64), ("solve_launches", C.c_int64), ("solve_ms", C.c_double),
        ("solve_units", C.c_double), ("solve_bytes", C.c_double), ("draw_launches", C.c_int64),
        ("draw_ms", C.c_double), ("other_launches", C.c_int64), ("other_ms", C.c_double),
        ("redo_units", C.c_int64), ("solve_flops", C.c_double),
    ]


PASS_KINDS = ("cold", "warm", "truth", "fiducial", "fd")      # MUSE_PASS_* of include/muse_b200.h


class muse_pass_profile(C.Structure):
    _fields_ = [("launches", C.c_int64 * 5), ("ms", C.c_double * 5), ("units", C.c_double * 5), ("bytes", C.c_double * 5)]


class muse_iterate_out(C.Structure):
    _fields_ = [
        ("n_iter", C.c_int32), ("theta_final", c_double_p), ("theta_hist", c_double_p), ("g_dat_hist", c_double_p),
        ("g_sims_hist", c_double_p), ("g_like_hist", c_double_p), ("g_prior_hist", c_double_p),
        ("h_inv_like_hist", c_double_p), ("h_prior_hist", c_double_p), ("h_inv_post_hist", c_double_p),
        ("seconds_hist", c_double_p), ("iters_hist", c_int32_p), ("fg_hist", c_int32_p), ("gnorm_hist", c_double_p),
        ("status_hist", c_int32_p),
    ]


class muse_cov_out(C.Structure):
    _fields_ = [("J", c_double_p), ("step", c_double_p), ("Hs", c_double_p), ("H", c_double_p),
                ("Sigma_inv", c_double_p), ("Sigma", c_double_p)]


# name → (restype, argtypes); must list every symbol include/muse_b200.h declares
SIGNATURES = {
    "muse_b200_abi_version": (C.c_int, []),
    "muse_b200_create": (C.c_int, [C.POINTER(muse_cfg), C.POINTER(C.c_void_p)]),
    "muse_b200_destroy": (C.c_int, [C.c_void_p]),
    "muse_b200_last_error": (C.c_char_p, [C.c_void_p]),
    "muse_b200_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "muse_b200_set_data": (C.c_int, [C.c_void_p, c_double_p]),
    "muse_b200_set_draws": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "muse_b200_set_draws_h": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "muse_b200_seed_draws": (C.c_int, [C.c_void_p, C.c_uint64]),
    "muse_b200_get_draws": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_double_p, c_double_p]),
    "muse_b200_set_z0": (C.c_int, [C.c_void_p, c_double_p]),
    "muse_b200_map_score": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_double, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, c_double_p, c_int32_p, c_int32_p, c_double_p, c_int32_p]),
    "muse_b200_map_score_async": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_double, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_int32]),
    "muse_b200_fetch": (C.c_int, [C.c_void_p, C.c_int32, c_double_p, c_int32_p, c_int32_p, c_double_p, c_int32_p]),
    "muse_b200_device_scores": (C.c_int, [C.c_void_p, C.POINTER(c_double_p), c_int32_p]),
    "muse_b200_comm_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "muse_b200_comm_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]),
    "muse_b200_comm_destroy": (C.c_int, [C.c_void_p]),
    "muse_b200_allgather_scores": (C.c_int, [C.c_void_p, C.c_int32, c_int32_p, c_double_p]),
    "muse_b200_allgather_rows": (C.c_int, [C.c_void_p, c_double_p, C.c_int32, c_int32_p, c_double_p]),
    "muse_b200_p2p_alloc": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_uint8)]),
    "muse_b200_p2p_connect": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint8)]),
    "muse_b200_p2p_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), c_int32_p]),
    "muse_b200_muse_iterate": (C.c_int, [C.c_void_p, c_double_p, C.c_int32, c_int32_p, C.c_int32, C.c_double, C.c_double,
                                         C.c_double, C.c_int32, c_double_p, c_double_p, C.POINTER(muse_iterate_out)]),
    "muse_b200_muse_covariance": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int32, C.c_int32, c_int32_p, C.c_double,
                                            c_double_p, C.POINTER(muse_cov_out)]),
    "muse_b200_muse_solve": (C.c_int, [C.c_void_p, c_double_p, C.c_int32, c_int32_p, C.c_int32, C.c_double, C.c_double, C.c_double,
                                       C.c_int32, c_double_p, c_double_p, C.c_int32, C.c_int32, c_int32_p,
                                       C.POINTER(muse_iterate_out), C.POINTER(muse_cov_out)]),
    "muse_b200_fd_jacobian": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int32, C.c_double, c_double_p, c_int32_p]),
    "muse_b200_implicit_h": (C.c_int, [C.c_void_p, c_double_p, C.c_int32, C.c_int32, C.c_int32, c_double_p, c_int32_p, c_int32_p]),
    "muse_b200_fd_start": (C.c_int, [C.c_void_p, C.c_int32]),
    "muse_b200_fd_scores": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int32, C.c_double, c_double_p, c_int32_p]),
    "muse_b200_get_maps": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_double_p]),
    "muse_b200_dgemm_host": (C.c_int, [c_double_p, c_double_p, c_double_p, C.c_int32, C.c_int32, C.c_int32]),
    "muse_b200_dgemm_time": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_double_p]),
    "muse_b200_profile_reset": (C.c_int, [C.c_void_p, C.c_int32]),
    "muse_b200_profile_get": (C.c_int, [C.c_void_p, C.POINTER(muse_profile)]),
    "muse_b200_profile_passes": (C.c_int, [C.c_void_p, C.POINTER(muse_pass_profile)]),
    "muse_b200_debug_timeline": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]),
    "muse_b200_geometry": (C.c_int, [C.c_void_p, c_int32_p, c_int32_p, c_int32_p]),
}

_lib = None


def library_path() -> str:
    return _build.LIB_PATH


def load_library() -> C.CDLL:
    """dlopen libmuse_b200.so and bind every declared symbol (raises if any is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise MuseBackendError(-2, f"{path} not built; run __graft_entry__.build() (nvcc, sm_100a). "
                                   "There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError ⇒ symbol missing
        fn.restype = res
        fn.argtypes = args
    if lib.muse_b200_abi_version() != ABI_VERSION:
        raise MuseBackendError(-1, "ABI version mismatch between _capi.py and libmuse_b200.so")
    _lib = lib
    return lib
