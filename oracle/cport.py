"""ctypes loader for oracle/csrc/muse_oracle.c (TEST INFRASTRUCTURE ONLY).

The C port is the timed CPU baseline (``bench.py`` cpu_baseline / ``--impl reference``) and a fast
checker for large shapes.  PARITY UNPINNED, like the rest of the oracle (oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_DIR, "_build", "libmuse_oracle.so")
_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build():
    subprocess.run(["make", "-s", "-C", _DIR], check=True)
    return _LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        lib = C.CDLL(_LIB)
        lib.muse_oracle_map_score.restype = C.c_int
        lib.muse_oracle_map_score.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_int,
                                              C.c_int, _dp, _dp, _ip, _ip, _dp, _ip, C.c_int]
        lib.muse_oracle_map_score_consts.restype = C.c_int
        lib.muse_oracle_map_score_consts.argtypes = lib.muse_oracle_map_score.argtypes + [_dp, _dp]
        lib.muse_oracle_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def max_threads() -> int:
    return load().muse_oracle_max_threads()


def map_score(family_id: int, xi, nu, xdat, theta_sim, theta_eval, atol, include_data, start_mode,
              z_start=None, want_z=False, nthreads=0, P=None, L=None):
    """Body of the mapped block (src/muse.jl:169-176 / 508-514) for a batch, on the CPU."""
    lib = load()
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    nu = np.ascontiguousarray(nu, dtype=np.float64)
    nsims, d = xi.shape
    ntheta = 2 if family_id == 2 else 1
    units = nsims + (1 if include_data else 0)
    xdat = np.ascontiguousarray(xdat if xdat is not None else np.zeros(d), dtype=np.float64)
    ts = np.ascontiguousarray(np.atleast_1d(theta_sim), dtype=np.float64)
    te = np.ascontiguousarray(np.atleast_1d(theta_eval), dtype=np.float64)
    z = None
    if start_mode == 1:
        z = np.array(z_start, dtype=np.float64, copy=True).reshape(units, d)
    elif want_z:
        z = np.zeros((units, d))
    g = np.empty((units, ntheta))
    iters = np.empty(units, dtype=np.int32)
    fg = np.empty(units, dtype=np.int32)
    gnorm = np.empty(units)
    status = np.empty(units, dtype=np.int32)
    Pc = np.ascontiguousarray(P, dtype=np.float64) if P is not None else None
    Lc = np.ascontiguousarray(L, dtype=np.float64) if L is not None else None
    rc = lib.muse_oracle_map_score_consts(family_id, d, nsims, xi.ctypes.data_as(_dp), nu.ctypes.data_as(_dp),
                                   xdat.ctypes.data_as(_dp), ts.ctypes.data_as(_dp), te.ctypes.data_as(_dp),
                                   float(atol), int(bool(include_data)), int(start_mode),
                                   z.ctypes.data_as(_dp) if z is not None else None, g.ctypes.data_as(_dp),
                                   iters.ctypes.data_as(_ip), fg.ctypes.data_as(_ip), gnorm.ctypes.data_as(_dp),
                                   status.ctypes.data_as(_ip), int(nthreads),
                                   Pc.ctypes.data_as(_dp) if Pc is not None else None,
                                   Lc.ctypes.data_as(_dp) if Lc is not None else None)
    if rc != 0:
        raise RuntimeError(f"muse_oracle_map_score failed: {rc}")
    return dict(g=g, iters=iters, fg_evals=fg, gnorm=gnorm, status=status, z=z)
