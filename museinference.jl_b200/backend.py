"""NumPy-facing wrapper of one libmuse_b200 handle (one GPU, one shard of simulations).

Thin by design: every method is one C-ABI call (include/muse_b200.h) with host NumPy buffers in
and out.  The arithmetic happens in the CUDA kernels; nothing here computes on the N×d batch.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import MuseBackendError

FAMILY_IDS = {"funnel": _capi.FAMILY_FUNNEL, "hiergauss": _capi.FAMILY_HIERGAUSS, "corrgauss": _capi.FAMILY_CORRGAUSS,
              "twolayer": _capi.FAMILY_TWOLAYER}
FAMILY_NTHETA = {"funnel": 1, "hiergauss": 2, "corrgauss": 1, "twolayer": 1}


def _f64(a, shape=None):
    arr = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and arr.shape != shape:
        raise ValueError(f"expected array of shape {shape}, got {arr.shape}")
    return arr


def _dp(arr):
    return arr.ctypes.data_as(_capi.c_double_p) if arr is not None else None


def _ip(arr):
    return arr.ctypes.data_as(_capi.c_int32_p) if arr is not None else None


class B200Backend:
    def __init__(self, family: str, d: int, nsims: int, *, sim_offset: int = 0, nsims_h: int = 0,
                 h_sim_offset: int = 0, device: int = 0, group: int = 0, cluster: int = 0, stream=None,
                 lbfgs_m: int = 0, max_iters: int = 0, kernel: int = 0, P=None, L=None):
        if family not in FAMILY_IDS:
            raise MuseBackendError(-5, f"model family {family!r} is outside the registered families "
                                       f"{sorted(FAMILY_IDS)}; Turing/Soss-defined models are not supported")
        self._lib = _capi.load_library()
        self.family, self.d, self.nsims = family, int(d), int(nsims)
        self.ntheta = FAMILY_NTHETA[family]
        self.nsims_h = int(nsims_h)
        self._P = _f64(P, (d, d)) if P is not None else None
        self._L = _f64(L, (d, d)) if L is not None else None
        cfg = _capi.muse_cfg(
            abi_version=_capi.ABI_VERSION, family=FAMILY_IDS[family], d=self.d, ntheta=self.ntheta,
            nsims=self.nsims, device=int(device), sim_offset=int(sim_offset), nsims_h=self.nsims_h, kernel=int(kernel),
            h_sim_offset=int(h_sim_offset), lbfgs_m=int(lbfgs_m), max_iters=int(max_iters), group=int(group),
            cluster=int(cluster), P=_dp(self._P), L=_dp(self._L),
            stream=None if stream is None else C.c_void_p(int(stream) if int(stream) != 0 else 1))
        self.comm = None
        self._h = C.c_void_p()
        rc = self._lib.muse_b200_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            msg = self._lib.muse_b200_last_error(None).decode()
            self._h = None
            raise MuseBackendError(rc, msg)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc != 0:
            raise MuseBackendError(rc, self._lib.muse_b200_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.muse_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_stream(self, stream):
        """``None`` → library-owned stream; an integer → that cudaStream_t (0, the legacy default stream, is passed
        as cudaStreamLegacy so that it is not mistaken for "none")."""
        if stream is None:
            arg = None
        else:
            arg = C.c_void_p(int(stream) if int(stream) != 0 else 1)
        self._check(self._lib.muse_b200_set_stream(self._h, arg))

    # ------------------------------------------------------------------ inputs
    def set_data(self, x):
        x = _f64(x, (self.d,))
        self._check(self._lib.muse_b200_set_data(self._h, _dp(x)))

    def set_draws(self, xi, nu, xi_master, nu_master):
        xi = _f64(xi, (self.nsims, self.d))
        nu = _f64(nu, (self.nsims, self.d))
        xm, nm = _f64(xi_master, (self.d,)), _f64(nu_master, (self.d,))
        self._check(self._lib.muse_b200_set_draws(self._h, _dp(xi), _dp(nu), _dp(xm), _dp(nm)))

    def set_draws_h(self, xi_h, nu_h):
        xi_h = _f64(xi_h, (self.nsims_h, self.d))
        nu_h = _f64(nu_h, (self.nsims_h, self.d))
        self._check(self._lib.muse_b200_set_draws_h(self._h, _dp(xi_h), _dp(nu_h)))

    def seed_draws(self, seed: int):
        self._check(self._lib.muse_b200_seed_draws(self._h, C.c_uint64(int(seed))))

    def get_draws(self, first: int, count: int):
        xi = np.empty((count, self.d))
        nu = np.empty((count, self.d))
        self._check(self._lib.muse_b200_get_draws(self._h, first, count, _dp(xi), _dp(nu)))
        return xi, nu

    def set_z0(self, z0):
        z0 = _f64(z0, (self.d,))
        self._check(self._lib.muse_b200_set_z0(self._h, _dp(z0)))

    # ------------------------------------------------------------------ hot path
    def _theta(self, th):
        th = np.ascontiguousarray(np.atleast_1d(np.asarray(th, dtype=np.float64)))
        if th.shape != (self.ntheta,):
            raise ValueError(f"θ must have {self.ntheta} component(s)")
        return th

    def map_score(self, theta_sim, theta_eval, atol, *, include_data: bool, warm_start: int,
                  first_sim: int = 0, count: int | None = None):
        count = self.nsims - first_sim if count is None else count
        units = count + (1 if include_data else 0)
        ts, te = self._theta(theta_sim), self._theta(theta_eval)
        g = np.empty((units, self.ntheta))
        iters = np.empty(units, dtype=np.int32)
        fg = np.empty(units, dtype=np.int32)
        gnorm = np.empty(units)
        status = np.empty(units, dtype=np.int32)
        self._check(self._lib.muse_b200_map_score(self._h, _dp(ts), _dp(te), float(atol), int(bool(include_data)),
                                                  int(warm_start), int(first_sim), int(count), _dp(g), _ip(iters),
                                                  _ip(fg), _dp(gnorm), _ip(status)))
        return dict(g=g, iters=iters, fg_evals=fg, gnorm=gnorm, status=status)

    def map_score_async(self, theta_sim, theta_eval, atol, *, include_data: bool, warm_start: int,
                        first_sim: int = 0, count: int | None = None):
        count = self.nsims - first_sim if count is None else count
        ts, te = self._theta(theta_sim), self._theta(theta_eval)
        self._check(self._lib.muse_b200_map_score_async(self._h, _dp(ts), _dp(te), float(atol),
                                                        int(bool(include_data)), int(warm_start), int(first_sim),
                                                        int(count)))
        return count + (1 if include_data else 0)

    def fetch(self, units: int):
        g = np.empty((units, self.ntheta))
        iters = np.empty(units, dtype=np.int32)
        fg = np.empty(units, dtype=np.int32)
        gnorm = np.empty(units)
        status = np.empty(units, dtype=np.int32)
        self._check(self._lib.muse_b200_fetch(self._h, int(units), _dp(g), _ip(iters), _ip(fg), _dp(gnorm), _ip(status)))
        return dict(g=g, iters=iters, fg_evals=fg, gnorm=gnorm, status=status)

    def device_scores(self):
        """(device pointer, capacity in units) of the score matrix of the last map_score[_async]."""
        ptr = _capi.c_double_p()
        cap = C.c_int32()
        self._check(self._lib.muse_b200_device_scores(self._h, C.byref(ptr), C.byref(cap)))
        return C.cast(ptr, C.c_void_p).value, cap.value

    # ------------------------------------------------------------------ exchange step (multi-GPU)
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = _capi.load_library().muse_b200_comm_unique_id(buf)
        if rc != 0:
            raise MuseBackendError(rc, "ncclGetUniqueId failed (is libnccl.so.2 loadable?)")
        return bytes(buf)

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self._lib.muse_b200_comm_init(self._h, int(nranks), int(rank), buf))
        self.comm = (int(nranks), int(rank))

    def p2p_alloc(self, nranks: int, rank: int, block_doubles: int) -> bytes:
        """This rank's peer-exchange region (include/muse_b200.h: muse_b200_p2p_alloc); returns its 64-byte IPC handle."""
        buf = (C.c_uint8 * 64)()
        self._check(self._lib.muse_b200_p2p_alloc(self._h, int(nranks), int(rank), int(block_doubles), buf))
        return bytes(buf)

    def p2p_connect(self, handles):
        """Map the peers' regions: ``handles`` = the 64-byte handles of all ranks in rank order."""
        blob = b"".join(handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self._check(self._lib.muse_b200_p2p_connect(self._h, buf))

    def p2p_info(self):
        blk, ready = C.c_int64(), C.c_int32()
        self._check(self._lib.muse_b200_p2p_info(self._h, C.byref(blk), C.byref(ready)))
        return int(blk.value), bool(ready.value)

    def allgather_scores(self, first_row: int, counts):
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        out = np.empty((int(counts.sum()), self.ntheta))
        self._check(self._lib.muse_b200_allgather_scores(self._h, int(first_row), _ip(counts), _dp(out)))
        return out

    def allgather_rows(self, local, counts):
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        local = np.ascontiguousarray(local, dtype=np.float64)
        ncol = local.shape[1]
        out = np.empty((int(counts.sum()), ncol))
        self._check(self._lib.muse_b200_allgather_rows(self._h, _dp(local), int(ncol), _ip(counts), _dp(out)))
        return out

    def _iterate_buffers(self, K: int, N: int):
        """History buffers of the in-library loops, allocated once per (maxsteps, nsims) and reused: the caller copies
        what it keeps."""
        nt, units = self.ntheta, self.nsims + 1
        cache = self.__dict__.setdefault("_iterate_bufs", {})
        if (K, N) not in cache:
            f = lambda *shape: np.zeros(shape)
            e32 = lambda: np.empty((K, units), dtype=np.int32)          # only rows < n_iter are written and read
            res = dict(theta_final=f(nt), theta_hist=f(K, nt), g_dat_hist=f(K, nt), g_sims_hist=np.empty((K, N, nt)),
                       g_like_hist=f(K, nt), g_prior_hist=f(K, nt), h_inv_like_hist=f(K, nt), h_prior_hist=f(K, nt),
                       h_inv_post_hist=f(K, nt), seconds_hist=f(K), iters_hist=e32(), fg_hist=e32(),
                       gnorm_hist=np.empty((K, units)), status_hist=e32())
            o = _capi.muse_iterate_out()
            for name, arr in res.items():
                setattr(o, name, _ip(arr) if arr.dtype == np.int32 else _dp(arr))
            cache.clear()
            cache[(K, N)] = (res, o)
        return cache[(K, N)]

    def muse_iterate(self, theta0, nsims_total: int, counts, maxsteps: int, theta_rtol, atol, alpha, first_start: int,
                     prior_mean=None, prior_sigma=None):
        """The outer θ loop of muse! inside the library (include/muse_b200.h: muse_b200_muse_iterate)."""
        N, K = int(nsims_total), int(maxsteps)
        t0 = self._theta(theta0)
        res, o = self._iterate_buffers(K, N)
        cnt = np.ascontiguousarray(counts, dtype=np.int32) if counts is not None else None
        pm = self._theta(prior_mean) if prior_mean is not None else None
        ps = self._theta(prior_sigma) if prior_sigma is not None else None
        self._check(self._lib.muse_b200_muse_iterate(self._h, _dp(t0), N, _ip(cnt), K, float(theta_rtol), float(atol),
                                                     float(alpha), int(first_start), _dp(pm), _dp(ps), C.byref(o)))
        out = dict(res)
        out["n_iter"] = int(o.n_iter)
        return out

    def muse_solve(self, theta0, nsims_total: int, counts, maxsteps: int, theta_rtol, atol, alpha, first_start: int,
                   prior_mean=None, prior_sigma=None, get_covariance: bool = False, nsims_h_total: int = 0, counts_h=None):
        """Loop + covariance stage with the θ update on the device (include/muse_b200.h: muse_b200_muse_solve).
        Returns (iterate dict as muse_iterate, covariance dict as muse_covariance or None).  The ctypes argument objects and
        the small input / output buffers are built once per call shape and reused (they cost more than the call)."""
        nt, N, K = self.ntheta, int(nsims_total), int(maxsteps)
        res, o = self._iterate_buffers(K, N)
        key = (K, N, bool(get_covariance), int(nsims_h_total), None if counts is None else tuple(int(c) for c in counts),
               None if counts_h is None else tuple(int(c) for c in counts_h), prior_mean is None, prior_sigma is None)
        cache = self.__dict__.setdefault("_solve_args", {})
        a = cache.get(key)
        if a is None:
            cache.clear()
            a = dict(t0=np.zeros(nt), pm=np.zeros(nt), ps=np.ones(nt),
                     cnt=np.ascontiguousarray(counts, dtype=np.int32) if counts is not None else None,
                     cnth=np.ascontiguousarray(counts_h, dtype=np.int32) if counts_h is not None else None)
            a["cres"] = dict(J=np.zeros((nt, nt)), step=np.zeros(nt), Hs=np.zeros((int(nsims_h_total), nt, nt)), H=np.zeros((nt, nt)),
                             Sigma_inv=np.zeros((nt, nt)), Sigma=np.zeros((nt, nt))) if get_covariance else None
            a["co"] = _capi.muse_cov_out(**{k: _dp(v) for k, v in a["cres"].items()}) if get_covariance else None
            a["ptr"] = (_dp(a["t0"]), _ip(a["cnt"]), _dp(a["pm"]) if prior_mean is not None else None,
                        _dp(a["ps"]) if prior_sigma is not None else None, _ip(a["cnth"]), C.byref(o),
                        C.byref(a["co"]) if get_covariance else None)
            cache[key] = a
        a["t0"][:] = theta0
        if prior_mean is not None:
            a["pm"][:] = prior_mean
        if prior_sigma is not None:
            a["ps"][:] = prior_sigma
        t0p, cntp, pmp, psp, cnthp, op, cop = a["ptr"]
        self._check(self._lib.muse_b200_muse_solve(self._h, t0p, N, cntp, K, float(theta_rtol), float(atol), float(alpha),
                                                   int(first_start), pmp, psp, 1 if get_covariance else 0, int(nsims_h_total), cnthp, op, cop))
        out = dict(res)
        out["n_iter"] = int(o.n_iter)
        return out, ({k: v.copy() for k, v in a["cres"].items()} if get_covariance else None)

    def muse_covariance(self, theta, gs, nsims_h_total: int, counts_h, atol, prior_sigma=None):
        """J, FD Jacobians, H and Σ after the loop (include/muse_b200.h: muse_b200_muse_covariance)."""
        nt = self.ntheta
        th = self._theta(theta)
        gs = _f64(gs).reshape(-1, nt)
        res = dict(J=np.zeros((nt, nt)), step=np.zeros(nt), Hs=np.zeros((int(nsims_h_total), nt, nt)), H=np.zeros((nt, nt)),
                   Sigma_inv=np.zeros((nt, nt)), Sigma=np.zeros((nt, nt)))
        o = _capi.muse_cov_out(**{k: _dp(v) for k, v in res.items()})
        cnt = np.ascontiguousarray(counts_h, dtype=np.int32) if counts_h is not None else None
        ps = self._theta(prior_sigma) if prior_sigma is not None else None
        self._check(self._lib.muse_b200_muse_covariance(self._h, _dp(th), _dp(gs), gs.shape[0], int(nsims_h_total), _ip(cnt),
                                                        float(atol), _dp(ps), C.byref(o)))
        return res

    def fd_jacobian(self, theta0, step, nsims_H: int, atol):
        t0 = self._theta(theta0)
        st = self._theta(step)
        Hs = np.empty((nsims_H, self.ntheta, self.ntheta))
        status = np.empty((nsims_H, self.ntheta, 2), dtype=np.int32)
        self._check(self._lib.muse_b200_fd_jacobian(self._h, _dp(t0), _dp(st), int(nsims_H), float(atol), _dp(Hs),
                                                    _ip(status)))
        return Hs, status

    def implicit_h(self, theta0, nsims_H: int, start: int = 0, cg_maxiter: int = 100):
        """Per-sim H of get_H!'s implicit-diff branch (include/muse_b200.h: muse_b200_implicit_h); returns (Hs, CG iterations, status)."""
        t0 = self._theta(theta0)
        Hs = np.zeros((nsims_H, self.ntheta, self.ntheta))
        iters = np.zeros((nsims_H, self.ntheta), dtype=np.int32)
        status = np.zeros(nsims_H, dtype=np.int32)
        self._check(self._lib.muse_b200_implicit_h(self._h, _dp(t0), int(nsims_H), int(start), int(cg_maxiter), _dp(Hs), _ip(iters), _ip(status)))
        return Hs, iters, status

    def fd_start(self, start: int):
        """Start of get_H!'s fiducial solve: START_ZEROS (default) or START_USER (the vector given to set_z0)."""
        self._check(self._lib.muse_b200_fd_start(self._h, int(start)))

    def fd_scores(self, theta_eval, theta_sims, nsims_H: int, atol):
        """Raw scores of the get_H! virtual sims at arbitrary sample points (include/muse_b200.h: muse_b200_fd_scores):
        ``theta_sims[2n + s]`` is the −/+ point of Jacobian column n; returns g[k, 2n + s, :] and the statuses."""
        nt = self.ntheta
        te = self._theta(theta_eval)
        ts = _f64(theta_sims, (2 * nt, nt))
        g = np.empty((nsims_H, 2 * nt, nt))
        status = np.empty((nsims_H, 2 * nt), dtype=np.int32)
        self._check(self._lib.muse_b200_fd_scores(self._h, _dp(te), _dp(ts), int(nsims_H), float(atol), _dp(g), _ip(status)))
        return g, status

    def get_maps(self, first_unit: int, count: int):
        z = np.empty((count, self.d))
        self._check(self._lib.muse_b200_get_maps(self._h, int(first_unit), int(count), _dp(z)))
        return z

    # ------------------------------------------------------------------ diagnostics
    def profile_reset(self, enable: bool = True):
        self._check(self._lib.muse_b200_profile_reset(self._h, int(enable)))

    def profile(self) -> dict:
        p = _capi.muse_profile()
        self._check(self._lib.muse_b200_profile_get(self._h, C.byref(p)))
        return {name: getattr(p, name) for name, _ in _capi.muse_profile._fields_}

    def profile_passes(self) -> dict:
        """Solver time, units and algorithmic bytes split by pass kind (cold / warm / truth / fiducial / fd)."""
        p = _capi.muse_pass_profile()
        self._check(self._lib.muse_b200_profile_passes(self._h, C.byref(p)))
        return {name: dict(launches=int(p.launches[i]), ms=float(p.ms[i]), units=float(p.units[i]), bytes=float(p.bytes[i]))
                for i, name in enumerate(_capi.PASS_KINDS)}

    def debug_timeline(self, items: int, fetch: bool = False):
        """Arm (fetch=False) or read (fetch=True) the controller timeline: items × 16 SM-clock stamps."""
        if not fetch:
            self._check(self._lib.muse_b200_debug_timeline(self._h, int(items), None))
            return None
        out = np.zeros((items, 16), dtype=np.int64)
        self._check(self._lib.muse_b200_debug_timeline(self._h, int(items), out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out

    def geometry(self) -> dict:
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self._lib.muse_b200_geometry(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(group_threads=a.value, cluster=b.value, groups=c.value)
