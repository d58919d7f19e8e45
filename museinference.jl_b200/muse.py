"""``muse`` / ``muse!`` / ``get_J!`` / ``get_H!`` / ``MuseResult`` on the B200 backend.

Host-side mirror of /root/reference/src/muse.jl with the same names, keyword arguments, defaults
and result fields; Python spells the in-place variants ``muse_``, ``get_J_``, ``get_H_`` (also
reachable as ``getattr(module, "muse!")`` etc.).  The three mapped blocks of the reference
(src/muse.jl:169-176, 417-442, 508-525) — everything that touches the N×d batch — are single calls
into libmuse_b200 (``backend.map_score`` / ``backend.fd_jacobian``); what remains here is the
O(nθ²) outer-solver arithmetic of src/muse.jl:159-166, 183-232, 411-413, 446, 529, 535-549, kept on
the host exactly as in the reference.

Keyword names: ASCII spellings are canonical, the reference's Unicode spellings are accepted too
(``θ_rtol``, ``∇z_logLike_atol``, ``α``, ``z₀``, ``H⁻¹_like′``, ``H⁻¹_update``).
"""
from __future__ import annotations

import math
import os
import pickle
import time
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _capi
from .parallel import LocalPool, block_partition
from .problem import AbstractMuseProblem, BaseDraws, FlatPrior, NormalPrior, SimpleMuseProblem
from ._capi import MuseBackendError

# Which in-library outer loop the common configuration takes when the caller does not say (fused_driver=True):
# "device" — θ update on the device, one host synchronisation per solve (csrc/muse_outer.cu; isotropic families);
# "host"   — host arithmetic between two passes (csrc/muse_driver.cu).  fused_driver=False: the line-by-line loop below.
DEFAULT_FUSED_DRIVER = os.environ.get("MUSE_FUSED_DRIVER", "device")

_KW_ALIASES = {
    "θ_rtol": "theta_rtol", "∇z_logLike_atol": "gradz_logLike_atol", "α": "alpha", "z₀": "z0",
    "H⁻¹_like′": "H_inv_like", "H⁻¹_update": "H_inv_update", "θ₀": "theta0",
}


def _ascii_kwargs(kw: dict) -> dict:
    return {_KW_ALIASES.get(k, k): v for k, v in kw.items()}


def central_fdm(p: int, q: int = 1):
    """``FiniteDifferences.central_fdm(p, q)`` as get_H!'s ``fdm`` keyword takes it (src/muse.jl:300): the symmetric integer grid
    of p points and the coefficients c with Σᵢ cᵢ gᵢᵏ = q!·δ_{kq} (exact rationals, rounded to Float64).  Only first
    derivatives of odd-point methods (q = 1; p = 3, 5, 7, …) can be served by the backend's ± sample points.  The step is explicit
    (``step`` or 0.1 ./ std(result.gs), src/muse.jl:411-413) or, when get_H! is called with neither, the package's own adaptive
    estimate (``AdaptedFDM`` below)."""
    from fractions import Fraction
    from math import factorial
    if q != 1 or p < 3 or p % 2 == 0:
        raise MuseBackendError(-5, "fdm: only central_fdm(p, 1) with an odd number of points p ≥ 3 is provided")
    grid = list(range(-(p // 2), p // 2 + 1))
    A = [[Fraction(g) ** k for g in grid] + [Fraction(factorial(q) if k == q else 0)] for k in range(p)]
    for c in range(p):
        piv = next(r for r in range(c, p) if A[r][c] != 0)
        A[c], A[piv] = A[piv], A[c]
        A[c] = [v / A[c][c] for v in A[c]]
        for r in range(p):
            if r != c and A[r][c] != 0:
                A[r] = [vr - A[r][c] * vc for vr, vc in zip(A[r], A[c])]
    return tuple(float(g) for g in grid), tuple(float(A[k][p]) for k in range(p))


def _fdm_coefs(grid, q):
    """Coefficients c with Σᵢ cᵢ gᵢᵏ = q!·δ_{kq}, k = 0 … p−1, on an arbitrary integer grid (exact rationals → Float64)."""
    from fractions import Fraction
    from math import factorial
    p = len(grid)
    A = [[Fraction(g) ** k for g in grid] + [Fraction(factorial(q) if k == q else 0)] for k in range(p)]
    for c in range(p):
        piv = next(r for r in range(c, p) if A[r][c] != 0)
        A[c], A[piv] = A[piv], A[c]
        A[c] = [v / A[c][c] for v in A[c]]
        for r in range(p):
            if r != c and A[r][c] != 0:
                A[r] = [vr - A[r][c] * vc for vr, vc in zip(A[r], A[c])]
    return tuple(float(A[k][p]) for k in range(p))


class AdaptedFDM:
    """``FiniteDifferences.central_fdm(p, q; adapt = 1)`` when it is called WITHOUT a step — what ``pjacobian`` does for
    ``step === nothing`` (src/util.jl:13), i.e. get_H! before any scores exist (src/muse.jl:411-413).  The package picks
    step = (q/(p−q)·C₁/C₂)^(1/p) with C₁ = eps(|f|)·Σ|c| and C₂ = |∇ᵖf|·Σ|c·gᵖ|/p!, the two magnitudes taken from the unadapted
    ``central_fdm(p + 2, p)`` at its default step (|∇ᵖf| → 10, eps(|f|) → eps) — the largest estimate over x − h, x, x + h and
    over the components of f — and caps it at 1000 × the default step.  ``magnitudes`` / ``step_from_magnitudes`` are split so that
    the function values can come from batched backend calls."""

    def __init__(self, p, q=1, adapt=1, condition=10.0):
        from fractions import Fraction
        from math import factorial
        self.p, self.q = int(p), int(q)
        M = self.p // 2
        self.grid = tuple(float(g) for g in range(-M, M + 1))
        ig = [int(g) for g in self.grid]
        self.coefs = _fdm_coefs(ig, self.q)
        self.coefs_nbhd = (_fdm_coefs([g - 1 for g in ig], self.q), self.coefs, _fdm_coefs([g + 1 for g in ig], self.q))
        self.condition = float(condition)
        self.df_mult = float(sum(abs(Fraction(c) * Fraction(g) ** self.p) for c, g in zip(self.coefs, ig)) / factorial(self.p))
        self.ferr_mult = sum(abs(c) for c in self.coefs)
        self.bound = AdaptedFDM(self.p + 2, self.p, adapt - 1, condition) if adapt >= 1 else None

    def _step_acc(self, df_magnitude, f_error):
        P, Q = self.p, self.q
        c1 = f_error * self.ferr_mult
        c2 = df_magnitude * self.df_mult
        step = (Q / (P - Q) * (c1 / c2)) ** (1 / P)
        return step, c1 * step ** (-Q) + c2 * step ** (P - Q)

    def default_step(self):
        return self._step_acc(self.condition, float(np.finfo(np.float64).eps))[0]

    def _limit(self, step):
        return min(step, 1000 * self.default_step())           # max_range = Inf: only the cap at 1000 × the default step applies

    def estimate(self, fs, step, coefs=None):
        coefs = self.coefs if coefs is None else coefs
        acc = fs[0] * coefs[0]
        for fk, ck in zip(fs[1:], coefs[1:]):                  # sum(fs .* coefs), left to right
            acc = acc + fk * ck
        return acc / step ** self.q

    def magnitudes(self, fs, step):
        df = max(float(np.max(np.abs(self.estimate(fs, step, c)))) for c in self.coefs_nbhd)
        return df, max(float(np.max(np.abs(v))) for v in fs)

    def step_from_magnitudes(self, df_magnitude, f_magnitude):
        if df_magnitude == 0.0 or f_magnitude == 0.0:
            return self._limit(self.default_step())
        return self._limit(self._step_acc(df_magnitude, float(np.spacing(f_magnitude)))[0])


class SimpleCovariance:
    """``CovarianceEstimation.SimpleCovariance(corrected=…)`` — get_J!'s ``covariance_method`` (src/muse.jl:494, 529): the sample
    covariance of the scores normalised by n − 1 (corrected, the reference's default) or by n.  Any callable
    ``gs (N×nθ) → nθ×nθ matrix`` may be passed instead (the package's shrinkage estimators are not provided)."""

    def __init__(self, corrected: bool = False):
        self.corrected = bool(corrected)

    def __call__(self, gs):
        return np.atleast_2d(np.cov(np.asarray(gs, dtype=np.float64), rowvar=False, ddof=1 if self.corrected else 0))


class FusedHistory:
    """``result.history`` of a solve that ran inside the library: the rows (one dict per iteration, src/muse.jl:211-221) are
    built from the library's history arrays on first access — a solve whose caller only wants θ̂ ± σ never pays for them.
    Behaves like the list it stands for (len, indexing, iteration, append, truth value); pickles as a plain list."""

    __slots__ = ("_H", "_secs", "_theta0", "_theta_final", "_items")

    def __init__(self, H, secs, theta0, theta_final):
        self._H, self._secs, self._theta0, self._theta_final, self._items = H, secs, theta0, theta_final, None

    def _rows(self):
        if self._items is None:
            H, secs, n = self._H, self._secs, len(self._secs)
            items, th_unreg_prev = [], self._theta0
            for k in range(n):
                gsk = H["g_sims_hist"][k]                     # identity transform: g and g′ are the same array
                g_like, g_prior = H["g_like_hist"][k], H["g_prior_hist"][k]
                H_inv_like = _diagm(H["h_inv_like_hist"][k])
                it_k, fg_k, gn_k, st_k = H["iters_hist"][k], H["fg_hist"][k], H["gnorm_hist"][k], H["status_hist"][k]
                items.append(dict(
                    theta=H["theta_hist"][k], theta_unreg=th_unreg_prev, theta_t=H["theta_hist"][k], theta_unreg_t=th_unreg_prev,
                    g_like_sims=gsk, g_like_sims_t=gsk, g_like_dat=H["g_dat_hist"][k], g_like=g_like,
                    g_prior=g_prior, g_post=g_like + g_prior,
                    H_inv_post=_diagm(H["h_inv_post_hist"][k]), H_prior=_diagm(H["h_prior_hist"][k]), H_inv_like=H_inv_like,
                    H_inv_like_sims=H_inv_like,
                    z_history_dat=dict(iters=int(it_k[0]), fg_evals=int(fg_k[0]), gnorm=float(gn_k[0]), status=int(st_k[0])),
                    z_history_sims=dict(iters=it_k[1:], fg_evals=fg_k[1:], gnorm=gn_k[1:], status=st_k[1:]),
                    t=secs[k], z_dat=None, z_sims=None))
                th_unreg_prev = H["theta_hist"][k + 1] if k + 1 < n else self._theta_final
            self._items = items
        return self._items

    def __len__(self):
        return len(self._secs) if self._items is None else len(self._items)

    def __bool__(self):
        return len(self) > 0

    def __getitem__(self, i):
        return self._rows()[i]

    def __iter__(self):
        return iter(self._rows())

    def append(self, entry):
        self._rows().append(entry)

    def __eq__(self, other):
        return list(self) == list(other)

    def __reduce__(self):
        return (list, (self._rows(),))

    def __repr__(self):
        return repr(self._rows())


@dataclass
class MuseResult:
    """src/muse.jl:29-42.  ``theta`` ↔ θ, ``Sigma``/``Sigma_inv`` ↔ Σ/Σ⁻¹, ``dist`` is the
    (mean, covariance) of the Normal/MvNormal the reference builds (src/muse.jl:542-546)."""

    theta: Optional[np.ndarray] = None
    H: Optional[np.ndarray] = None
    J: Optional[np.ndarray] = None
    Sigma_inv: Optional[np.ndarray] = None
    Sigma: Optional[np.ndarray] = None
    dist: Optional[tuple] = None
    history: list = field(default_factory=list)
    gs: object = field(default_factory=list)      # N×nθ array once filled (one row per sim)
    Hs: object = field(default_factory=list)      # n_H×nθ×nθ array once filled
    metadata: dict = field(default_factory=dict)
    rng: object = None
    time: float = 0.0

    def __repr__(self):   # src/muse.jl:45-59
        if self.theta is None:
            return "MuseResult()"
        if self.Sigma is not None:
            sd = np.sqrt(np.diag(self.Sigma))
            return "MuseResult(" + ", ".join(f"{t:.4g}±{s:.3g}" for t, s in zip(self.theta, sd)) + ")"
        return "MuseResult(" + ", ".join(f"{t:.4g}" for t in self.theta) + ")"


def _prior_mean_sigma(prior, ntheta):
    """(mean, sigma) arrays of an independent-Normal prior, None for a flat prior, False for anything else."""
    if type(prior) is FlatPrior:
        return None
    if type(prior) is NormalPrior:
        mean, sigma = np.empty(ntheta), np.empty(ntheta)
        mean[:] = prior.mean
        sigma[:] = prior.sigma
        return mean, sigma
    return False


def _diagm(v):
    """np.diag(v) for a short 1-D v without its generic-path overhead (called 3× per iteration of every solve)."""
    n = v.size
    a = np.zeros((n, n))
    a.flat[::n + 1] = v
    return a


def _check_problem(prob):
    if not isinstance(prob, SimpleMuseProblem):
        raise MuseBackendError(-5, f"{type(prob).__name__} is not supported by the B200 backend: only "
                                   "SimpleMuseProblem over a registered family (funnel, hiergauss, corrgauss, twolayer); "
                                   "Turing/Soss-defined models raise, there is no CPU fallback")


def _check_status(out, what, skip_errors=False, pool=None):
    """src/interface.jl:168-171: non-convergence warns, a non-finite objective is an error.  With several ranks the
    decision is collective: every rank learns whether ANY rank saw a failure and raises (or not) together — a rank that
    raised alone would leave the others waiting in the next exchange step."""
    bad = np.flatnonzero(out["status"] == _capi.STATUS_NONFINITE)
    any_bad = bool(bad.size)
    if pool is not None and pool.world > 1 and not skip_errors:
        any_bad = bool(pool.any_flag(any_bad))
    if any_bad and not skip_errors:
        where = f"unit(s) {bad[:8].tolist()}" if bad.size else "a unit of another rank"
        raise FloatingPointError(f"{what}: MAP solution failed with a non-finite objective for {where}")
    return bad


# =============================================================================== muse
def muse(prob, theta0, **kwargs):
    """src/muse.jl:107."""
    return muse_(MuseResult(), prob, theta0, **kwargs)


def muse_(result: MuseResult, prob: AbstractMuseProblem, theta0=None, **kwargs):
    """``muse!`` — src/muse.jl:112-250."""
    kw = _ascii_kwargs(kwargs)
    rng = kw.pop("rng", None)
    z0 = kw.pop("z0", None)
    maxsteps = kw.pop("maxsteps", 50)
    theta_rtol = kw.pop("theta_rtol", 1e-1)
    atol = kw.pop("gradz_logLike_atol", 1e-2)
    nsims = kw.pop("nsims", 100)
    alpha = kw.pop("alpha", 0.7)
    kw.pop("progress", False)
    pool = kw.pop("pool", None) or LocalPool()
    regularize = kw.pop("regularize", None)
    regularize_is_identity = regularize is None
    regularize = regularize or (lambda t: t)
    fused_driver = kw.pop("fused_driver", True)
    H_inv_like = kw.pop("H_inv_like", None)
    H_inv_update = kw.pop("H_inv_update", "sims")
    broyden_memory = kw.pop("broyden_memory", math.inf)
    checkpoint_filename = kw.pop("checkpoint_filename", None)
    get_covariance = kw.pop("get_covariance", False)
    save_MAPs = kw.pop("save_MAPs", False)
    if kw:
        raise TypeError(f"muse!: unknown keyword argument(s) {sorted(kw)}")
    _check_problem(prob)

    # :134  rng: given, else the result's, else seed 0.  The reference falls back to copy(Random.default_rng()), which does
    # not advance the global stream either: repeated calls without an rng reuse the same sims there too, unless the
    # caller consumed the global RNG in between.  Pass distinct seeds for independent solves.
    if rng is None:
        rng = result.rng if result.rng is not None else 0
    result.rng = rng
    # :135-136
    theta = prob.standardize_theta(result.theta if result.theta is not None else theta0)
    theta_unreg = theta.copy()
    theta_t = prob.transform_theta(theta)                                          # θ′ — also what the kernels take
    theta_unreg_t = theta_t.copy()
    transformed = prob.has_transform
    history = result.history
    alpha_fn = alpha if callable(alpha) else (lambda i: alpha)                     # :145-149

    nh_total = max(1, nsims // 10) if get_covariance else 0
    be = prob.backend_for(nsims, rng, pool, nh_total)
    first_pass = True                                                              # :151 ẑs = zeros | z₀
    if z0 is not None:
        be.set_z0(z0)

    # Common configuration: the whole loop below runs inside the library (csrc/muse_driver.cu), same arithmetic,
    # no interpreter between two passes.  Anything else (callable α, regularize, Broyden, save_MAPs, resume,
    # checkpoints, a custom prior, a test double as backend) takes the line-by-line loop.
    prior_ms = _prior_mean_sigma(prob.prior, prob.ntheta)
    fused = (fused_driver and not history and not callable(alpha) and regularize_is_identity and H_inv_like is None
             and H_inv_update == "sims" and not save_MAPs and checkpoint_filename is None and prior_ms is not False
             and hasattr(be, "muse_iterate") and maxsteps >= 1 and nsims >= 2 and not transformed)
    if fused:
        counts = None
        if pool.world > 1:
            try:
                pool.bind(be, nsims, nh_total)
            except TypeError:                   # a pool without the peer-exchange set-up
                pool.bind(be)
            counts = block_partition(nsims, pool.world)[1]
        mode = DEFAULT_FUSED_DRIVER if fused_driver is True else fused_driver
        device_loop = (mode == "device" and hasattr(be, "muse_solve") and prob.family not in ("corrgauss", "twolayer") and maxsteps <= 64
                       and pool.world <= 16)      # the limits of csrc/muse_outer.cu (history rows, rank table); else the host loop
        cdev = None
        if device_loop:
            counts_h = block_partition(nh_total, pool.world)[1] if (pool.world > 1 and get_covariance) else None
            tc0 = time.perf_counter()
            r, cdev = be.muse_solve(theta, nsims, counts, maxsteps, theta_rtol, atol, alpha,
                                    _capi.START_USER if z0 is not None else _capi.START_ZEROS,
                                    *(prior_ms if prior_ms else (None, None)), get_covariance=bool(get_covariance),
                                    nsims_h_total=nh_total, counts_h=counts_h)
            t_solve = time.perf_counter() - tc0
        else:
            r = be.muse_iterate(theta, nsims, counts, maxsteps, theta_rtol, atol, alpha,
                                _capi.START_USER if z0 is not None else _capi.START_ZEROS,
                                *(prior_ms if prior_ms else (None, None)))
        # the library's history buffers are reused by the next call: take ONE private copy of the executed rows of each
        # array; the per-iteration dicts of result.history are built from those copies when somebody looks at them
        n = r["n_iter"]
        H = {key: r[key][:n].copy() for key in ("theta_hist", "g_dat_hist", "g_sims_hist", "g_like_hist", "g_prior_hist", "h_inv_like_hist",
                                              "h_prior_hist", "h_inv_post_hist", "iters_hist", "fg_hist", "gnorm_hist", "status_hist")}
        secs = r["seconds_hist"][:n].tolist()
        theta_final = r["theta_final"].copy()
        result.history = history = FusedHistory(H, secs, theta.copy(), theta_final)
        result.time += sum(secs)
        if n:
            result.theta = theta_final.copy()                                      # :230
            result.gs = H["g_sims_hist"][n - 1]                                    # :231 (a view of this solve's private copy)
        maxsteps = 0                                                               # the loop below has nothing left to do
        if cdev is not None and r["n_iter"]:                                       # :244-247, already done on the stream
            result.J, result.H, result.Hs = cdev["J"], cdev["H"], cdev["Hs"]
            result.Sigma_inv, result.Sigma = cdev["Sigma_inv"], cdev["Sigma"]
            result.dist = (result.theta.copy(), result.Sigma.copy())               # :542-546
            result.metadata["fd_step"] = cdev["step"]
            result.time = t_solve
            return result
        if get_covariance and r["n_iter"]:                                         # :244-247, same stage in the library
            tc = time.perf_counter()
            counts_h = block_partition(nh_total, pool.world)[1] if pool.world > 1 else None
            c = be.muse_covariance(result.theta, result.gs, nh_total, counts_h, atol, prior_ms[1] if prior_ms else None)
            result.J, result.H, result.Hs = c["J"], c["H"], c["Hs"]
            result.Sigma_inv, result.Sigma = c["Sigma_inv"], c["Sigma"]
            result.dist = (result.theta.copy(), result.Sigma.copy())               # :542-546
            result.metadata["fd_step"] = c["step"]
            result.time += time.perf_counter() - tc
            return result

    for i in range(len(history) + 1, maxsteps + 1):                                # :159
        t0 = time.perf_counter()
        if i > 2:                                                                  # :163-166 (Δθ′)
            dth = history[-1]["theta_t"] - history[-2]["theta_t"]
            q = -(dth @ history[-1]["H_inv_post"] @ dth)
            if q < 0:
                raise ValueError("DomainError: sqrt of a negative number in the θ convergence test (src/muse.jl:165)")
            if math.sqrt(q) < theta_rtol:
                break

        # MUSE gradient: the mapped block :169-176 is one backend call on this rank's shard.  The kernels take θ′ and
        # return g′ = ∇θ′ logLike (:173); g = ∇θ logLike (:172) is g′ ⊘ ∂θ/∂θ′ (identical arrays without a transform).
        warm = (_capi.START_USER if z0 is not None else _capi.START_ZEROS) if first_pass else _capi.START_PREV
        if getattr(pool, "uses_device_gather", lambda: False)():
            # multi-GPU: the scores are all-gathered with NCCL from device memory on the launch stream
            units = be.map_score_async(theta_t, theta_t, atol, include_data=True, warm_start=warm)
            g_like_sims_t = pool.allgather_device_scores(be, 1, nsims)             # the one exchange step
            out = be.fetch(units)
        else:
            out = be.map_score(theta_t, theta_t, atol, include_data=True, warm_start=warm)
            g_like_sims_t = pool.allgather_rows(out["g"][1:], nsims)               # the one exchange step
        first_pass = False
        _check_status(out, "muse!", pool=pool)
        g_like_dat = out["g"][0].copy()                                            # :177-178 (g_like_dat′)
        g_like_sims = g_like_sims_t / prob.dinv_transform(theta_t) if transformed else g_like_sims_t   # :177 (g)

        g_like = g_like_dat - np.mean(g_like_sims_t, axis=0)                       # :183
        g_prior = prob.prior_grad_t(theta_t) if transformed else np.asarray(prob.prior.grad(theta), dtype=np.float64)   # :184
        g_post = g_like + g_prior                                                  # :185

        h_inv_like_sims = -1.0 / np.var(g_like_sims_t, axis=0, ddof=1)             # :188
        H_inv_like_sims = np.diag(h_inv_like_sims)                                 # :189
        if H_inv_like is None or H_inv_update == "sims":                           # :190-191
            H_inv_like = H_inv_like_sims
        elif i > 2 and H_inv_update in ("broyden", "diagonal_broyden"):            # :192-205
            j0 = int(max(2, i - broyden_memory))
            H_inv_like = history[j0 - 2]["H_inv_like_sims"]
            for j in range(j0, i):
                d_th = history[j - 1]["theta_t"] - history[j - 2]["theta_t"]
                d_g = history[j - 1]["g_like"] - history[j - 2]["g_like"]
                H_inv_like = H_inv_like + np.outer((d_th - H_inv_like @ d_g) / (d_th @ H_inv_like @ d_g), d_th) @ H_inv_like
                if H_inv_update == "diagonal_broyden":
                    H_inv_like = np.diag(np.diag(H_inv_like))

        H_prior = prob.prior_hess_t(theta_t) if transformed else np.asarray(prob.prior.hess(theta), dtype=np.float64)   # :207
        H_inv_post = np.linalg.inv(np.linalg.inv(H_inv_like) + H_prior)            # :208

        t = time.perf_counter() - t0
        entry = dict(                                                              # :211-221
            theta=theta.copy(), theta_unreg=theta_unreg.copy(), theta_t=theta_t.copy(), theta_unreg_t=theta_unreg_t.copy(),
            g_like_sims=g_like_sims.copy(), g_like_sims_t=g_like_sims_t.copy(), g_like_dat=g_like_dat, g_like=g_like, g_prior=g_prior, g_post=g_post,
            H_inv_post=H_inv_post, H_prior=H_prior, H_inv_like=np.array(H_inv_like, copy=True),
            H_inv_like_sims=H_inv_like_sims,
            z_history_dat=dict(iters=int(out["iters"][0]), fg_evals=int(out["fg_evals"][0]),
                               gnorm=float(out["gnorm"][0]), status=int(out["status"][0])),
            z_history_sims=dict(iters=out["iters"][1:].copy(), fg_evals=out["fg_evals"][1:].copy(),
                                gnorm=out["gnorm"][1:].copy(), status=out["status"][1:].copy()),
            t=t, z_dat=None, z_sims=None,
        )
        if save_MAPs:                                                              # :139-143, 219
            keep = save_MAPs if callable(save_MAPs) else (lambda z: z)
            entry["z_dat"] = keep(be.get_maps(0, 1)[0])
            entry["z_sims"] = keep(be.get_maps(1, be.nsims))
        history.append(entry)

        theta_unreg_t = theta_t - alpha_fn(i) * (H_inv_post @ g_post)              # :224
        theta_unreg = prob.inv_transform_theta(theta_unreg_t)                      # :225
        theta_t = prob.standardize_theta(regularize(theta_unreg_t))                # :226
        theta = prob.inv_transform_theta(theta_t)                                  # :227

        result.theta = theta_unreg.copy()                                          # :230
        result.gs = g_like_sims.copy()                                             # :231 (N×nθ array, one row per sim)
        result.time += t                                                           # :232
        if checkpoint_filename is not None and pool.rank == 0:                     # :234
            with open(checkpoint_filename, "wb") as fh:
                pickle.dump(result, fh)

    if get_covariance:                                                             # :244-247
        get_J_(result, prob, rng=rng, nsims=nsims, gradz_logLike_atol=atol, pool=pool, _nh_total=nh_total)
        get_H_(result, prob, rng=rng, nsims=max(1, nsims // 10), gradz_logLike_atol=atol, pool=pool, _nsims_total=nsims)
    return result


# =============================================================================== get_J!
def get_J_(result: MuseResult, prob: AbstractMuseProblem, theta0=None, **kwargs):
    """``get_J!`` — src/muse.jl:484-532."""
    kw = _ascii_kwargs(kwargs)
    z0 = kw.pop("z0", None)
    atol = kw.pop("gradz_logLike_atol", 1e-2)
    rng = kw.pop("rng", None)
    nsims = kw.pop("nsims", 100)
    pool = kw.pop("pool", None) or LocalPool()
    kw.pop("progress", False)
    skip_errors = kw.pop("skip_errors", False)
    covariance_method = kw.pop("covariance_method", None) or SimpleCovariance(corrected=True)   # :494
    if not callable(covariance_method):
        raise TypeError("get_J!: covariance_method must be a SimpleCovariance or a callable gs → matrix")
    nh_total = kw.pop("_nh_total", 0)
    if kw:
        raise TypeError(f"get_J!: unknown keyword argument(s) {sorted(kw)}")
    _check_problem(prob)
    if rng is None:
        rng = result.rng if result.rng is not None else 0
    theta0 = prob.standardize_theta(theta0 if theta0 is not None else result.theta)   # :498
    nsims_existing = len(result.gs)
    nsims_remaining = nsims - nsims_existing                                       # :499-500
    if nsims_remaining > 0:
        # rngs = split_rng(rng, nsims)[nsims_existing+1:end]  (:506) → global sims [existing, nsims)
        be = prob.backend_for(nsims, rng, pool, nh_total)
        off, cnt = pool.shard(nsims)
        lo = min(max(nsims_existing, off) - off, cnt)      # a shard that lies wholly below nsims_existing has nothing to do
        n_local = cnt - lo
        if z0 is not None:
            be.set_z0(z0)
        warm = _capi.START_USER if z0 is not None else _capi.START_TRUTH           # :511
        theta0_t = prob.transform_theta(theta0)
        if n_local > 0:
            out = be.map_score(theta0_t, theta0_t, atol, include_data=False, warm_start=warm, first_sim=lo, count=n_local)
        else:
            out = dict(g=np.zeros((0, prob.ntheta)), status=np.zeros(0, dtype=np.int32))
        bad = _check_status(out, "get_J!", skip_errors, pool)
        g_local = out["g"] / prob.dinv_transform(theta0_t) if prob.has_transform else out["g"]   # :513 UnTransformedθ()
        if pool.world == 1:
            g_new = np.delete(g_local, bad, axis=0) if bad.size else g_local       # skipmissing (:508)
        else:
            # gather blocks of the *full* partition; ranks pad the part below nsims_existing with NaN rows
            full = np.full((cnt, prob.ntheta), np.nan)
            full[lo:] = g_local
            if bad.size:
                full[lo + bad] = np.nan
            allg = pool.allgather_rows(full, nsims)[nsims_existing:]
            g_new = allg[~np.isnan(allg).any(axis=1)]
        old = np.asarray(result.gs, dtype=np.float64).reshape(-1, prob.ntheta)
        result.gs = np.concatenate([old, np.asarray(g_new).reshape(-1, prob.ntheta)], axis=0)
    gs = np.asarray(result.gs, dtype=np.float64).reshape(-1, prob.ntheta)
    if theta0.size == 1:
        result.J = np.array([[np.var(gs[:, 0], ddof=1)]])                          # :529 var
    else:
        result.J = np.asarray(covariance_method(gs), dtype=np.float64)             # :529 cov(covariance_method, gs)
    finalize_result_(result, prob)
    return result


# =============================================================================== get_H!
def get_H_(result: MuseResult, prob: AbstractMuseProblem, theta0=None, **kwargs):
    """``get_H!`` finite-difference branch — src/muse.jl:296-333, 407-450."""
    kw = _ascii_kwargs(kwargs)
    atol = kw.pop("gradz_logLike_atol", 1e-2)
    rng = kw.pop("rng", None)
    nsims = kw.pop("nsims", 10)
    step = kw.pop("step", None)
    pool = kw.pop("pool", None) or LocalPool()
    kw.pop("pmap_over", None)
    kw.pop("progress", False)
    skip_errors = kw.pop("skip_errors", False)
    z0 = kw.pop("z0", None)
    nsims_total = kw.pop("_nsims_total", 0)
    implicit_diff = bool(kw.pop("implicit_diff", False))                           # :310, 335-405
    kw.pop("implicit_diff_H1_is_zero", False)          # H1 ≡ 0 for every registered family (their score does not depend on x)
    cg_kwargs = dict(kw.pop("implicit_diff_cg_kwargs", None) or {})
    cg_maxiter = int(cg_kwargs.pop("maxiter", 100))
    if cg_kwargs.pop("Pl", None) is not None or cg_kwargs:
        raise MuseBackendError(-5, "implicit_diff_cg_kwargs: only `maxiter` is provided (the reference's default preconditioner is the identity)")
    fdm = kw.pop("fdm", None)                                                      # :300; None = central_fdm(3,1)
    if fdm is not None:
        try:
            grid, coefs = fdm
            grid, coefs = tuple(float(g) for g in grid), tuple(float(c) for c in coefs)
            M = len(grid) // 2
            assert len(grid) == len(coefs) and len(grid) % 2 == 1 and grid == tuple(float(g) for g in range(-M, M + 1))
        except Exception:
            raise MuseBackendError(-5, "fdm must come from central_fdm(p, 1) (odd p): a symmetric integer grid and its coefficients")
        if grid == (-1.0, 0.0, 1.0) and coefs == (-0.5, 0.0, 0.5):
            fdm = None
    if kw:
        raise TypeError(f"get_H!: unknown keyword argument(s) {sorted(kw)}")
    _check_problem(prob)
    if rng is None:
        rng = result.rng if result.rng is not None else 0
    theta0 = prob.standardize_theta(theta0 if theta0 is not None else result.theta)   # :315
    nsims_existing = len(result.Hs)
    nsims_remaining = nsims - nsims_existing                                       # :317-319
    if nsims_remaining <= 0:
        return result
    t0 = time.perf_counter()
    if implicit_diff:
        # H = H1 + H2 per sim, H2 = −(∂θ∇z logLike)ᵀ A⁻¹ (∂θ_sim ∇z logLike) by conjugate gradients (:340-388); the nested-AD
        # derivatives of the reference are closed forms for the registered families (csrc/muse_implicit.cu)
        if pool.world > 1 or prob.has_transform:
            raise MuseBackendError(-5, "get_H!(implicit_diff=true) is provided on one GPU and for the identity θ-transform")
        be = prob.backend_for(max(nsims_total, nsims_remaining), rng, pool, 0)
        if not hasattr(be, "implicit_h"):
            raise MuseBackendError(-5, "get_H!(implicit_diff=true): not provided by this backend")
        if z0 is not None:
            be.set_z0(z0)
        Hs_new, iters, status = be.implicit_h(theta0, nsims_remaining, _capi.START_USER if z0 is not None else _capi.START_ZEROS, cg_maxiter)
        bad = np.flatnonzero(status == _capi.STATUS_NONFINITE)
        if bad.size and not skip_errors:
            raise FloatingPointError("get_H!: MAP solution failed with a non-finite objective")
        keep = np.setdiff1d(np.arange(nsims_remaining), bad)
        nt = prob.ntheta
        oldH = np.asarray(result.Hs, dtype=np.float64).reshape(-1, nt, nt)
        result.Hs = np.concatenate([oldH, Hs_new[keep]], axis=0)
        result.metadata.setdefault("implicit_diff_cg_hists", []).extend(iters[keep].tolist())   # :404 (iteration counts)
        result.H = np.mean(result.Hs, axis=0)                                      # :446
        result.time += time.perf_counter() - t0
        finalize_result_(result, prob)
        return result
    if step is None and len(result.gs) > 0:                                        # :411-413
        step = 0.1 / np.std(np.asarray(result.gs, dtype=np.float64).reshape(-1, prob.ntheta), axis=0, ddof=1)
    adaptive = step is None                    # src/util.jl:13: fdm(f, 0.0) — FiniteDifferences estimates a step per sim and component
    if adaptive and (pool.world > 1 or prob.has_transform):
        raise MuseBackendError(-5, "get_H!: without `step` and without scores from get_J!/muse! the finite-difference step is "
                                   "estimated per simulation, which is provided on one GPU and for the identity θ-transform")
    if not adaptive:
        step = prob.standardize_theta(step)

    # rngs = split_rng(rng, nsims_remaining)  (:323): the first nsims_remaining child streams
    n_total = max(nsims_total, nsims_remaining)
    be = prob.backend_for(n_total, rng, pool, nsims_remaining)
    _, hcnt = pool.shard(nsims_remaining)
    if z0 is not None:                                                             # :419  z_start = @something(z₀, …)
        if not hasattr(be, "fd_start"):
            raise MuseBackendError(-5, "get_H!: this backend cannot start the fiducial solve from a user z₀")
        be.set_z0(z0)
        be.fd_start(_capi.START_USER)
    try:
        if adaptive:
            Hs_local, status, steps = _fd_jacobian_adaptive(be, prob, theta0, hcnt, atol, fdm)
            result.metadata["fd_adaptive_steps"] = steps
        elif not prob.has_transform and fdm is None:
            Hs_local, status = be.fd_jacobian(theta0, step, hcnt, atol)            # :417-442 + src/util.jl:9-26
        else:
            # pjacobian perturbs the UNtransformed θ₀ (src/util.jl:15; sims at θ, MAP and score at θ₀, :430-432); the kernels
            # take θ′, so the sample points are mapped one by one and sum(fs .* coefs) / step and the change of variables
            # g = g′ ⊘ ∂θ/∂θ′ are formed here.  One fd_scores call serves the ± pair of one grid distance m (p-point central
            # methods: m = 1 … (p − 1)/2); the centre point, whose coefficient is 0, is not executed.
            nt = prob.ntheta
            grid, coefs = fdm if fdm is not None else ((-1.0, 0.0, 1.0), (-0.5, 0.0, 0.5))
            M = len(grid) // 2
            theta0_t = prob.transform_theta(theta0)
            dinv = prob.dinv_transform(theta0_t) if prob.has_transform else 1.0
            g_at = {}                                                              # grid value → scores (hcnt × nθ columns × nθ)
            status = np.zeros((hcnt, nt, 2), dtype=np.int32)
            for m_ in range(1, M + 1):
                pts = np.empty((2 * nt, nt))
                for n in range(nt):
                    for sgn in (0, 1):
                        th = theta0.copy()
                        th[n] = theta0[n] + (0.0 + step[n] * (float(m_) if sgn else -float(m_)))
                        pts[2 * n + sgn] = prob.transform_theta(th)
                g_t, st = be.fd_scores(theta0_t, pts, hcnt, atol)
                g_u = g_t / dinv
                g_at[-float(m_)] = g_u[:, 0::2, :]
                g_at[float(m_)] = g_u[:, 1::2, :]
                status = np.maximum(status, np.where(st.reshape(hcnt, nt, 2) == _capi.STATUS_NONFINITE, _capi.STATUS_NONFINITE, 0))
            Hs_local = np.empty((hcnt, nt, nt))
            for n in range(nt):
                acc = g_at[grid[0]][:, n, :] * coefs[0]
                for gk, ck in zip(grid[1:], coefs[1:]):
                    acc = acc + (g_at[gk][:, n, :] * ck if gk != 0.0 else 0.0)
                Hs_local[:, :, n] = acc / step[n]
    finally:
        if z0 is not None:
            be.fd_start(_capi.START_ZEROS)
    bad_local = np.flatnonzero((status.reshape(hcnt, -1) == _capi.STATUS_NONFINITE).any(axis=1))
    any_bad = bool(bad_local.size)
    if pool.world > 1 and not skip_errors:
        any_bad = bool(pool.any_flag(any_bad))           # collective decision: all ranks raise together
    if any_bad and not skip_errors:
        raise FloatingPointError("get_H!: MAP solution failed with a non-finite objective")
    nt = prob.ntheta
    flat = Hs_local.reshape(hcnt, nt * nt).copy()
    if bad_local.size:
        flat[bad_local] = np.nan
    if getattr(pool, "uses_device_gather", lambda: False)():
        allH = pool.allgather_host_rows(be, flat, nsims_remaining)
    else:
        allH = pool.allgather_rows(flat, nsims_remaining)
    allH = allH[~np.isnan(allH).any(axis=1)]
    oldH = np.asarray(result.Hs, dtype=np.float64).reshape(-1, nt, nt)
    result.Hs = np.concatenate([oldH, allH.reshape(-1, nt, nt)], axis=0)           # (n_H × nθ × nθ array)

    result.H = np.mean(result.Hs, axis=0)                                          # :446
    result.time += time.perf_counter() - t0
    finalize_result_(result, prob)
    return result


def _fd_jacobian_adaptive(be, prob, theta0, count, atol, fdm):
    """The finite-difference Jacobians of get_H! with FiniteDifferences' own step (``AdaptedFDM``).  Stage 1 — the bound estimator's
    p + 2 sample points are the same for every sim (its step is the default one): (p + 1)/2 ``fd_scores`` calls (one per grid
    distance, both signs) and one at θ₀ serve all sims and components, and give every sim its steps.  Stage 2 — the steps differ per
    sim, the backend takes ONE table of sample points per call: sim k is served by a call over sims 0 … k with its points, of which
    row k is kept (n_H(n_H + 1)/2 solves per grid distance: get_H!'s default is n_H = 10).  The centre value is the one of stage 1."""
    nt = prob.ntheta
    grid, coefs = fdm if fdm is not None else ((-1.0, 0.0, 1.0), (-0.5, 0.0, 0.5))
    adm = AdaptedFDM(len(grid), 1)
    bnd = adm.bound
    hb = bnd._limit(bnd.default_step())
    status = np.zeros((count, nt, 2), dtype=np.int32)

    def scores(offsets, n_units):
        """offsets[n] = (ε₋, ε₊) of component n → (g at ε₋, g at ε₊), each n_units × nθ (column) × nθ."""
        pts = np.empty((2 * nt, nt))
        for n in range(nt):
            for sgn in (0, 1):
                th = theta0.copy()
                th[n] = theta0[n] + offsets[n][sgn]
                pts[2 * n + sgn] = th
        g, st = be.fd_scores(theta0, pts, n_units, atol)
        bad = np.where(st.reshape(n_units, nt, 2) == _capi.STATUS_NONFINITE, _capi.STATUS_NONFINITE, 0)
        status[:n_units] = np.maximum(status[:n_units], bad)
        return g[:, 0::2, :], g[:, 1::2, :]

    g_b = {}
    for m_ in range(1, len(bnd.grid) // 2 + 1):
        g_b[-float(m_)], g_b[float(m_)] = scores([(0.0 + hb * -float(m_), 0.0 + hb * float(m_))] * nt, count)
    g_b[0.0], _ = scores([(0.0, 0.0)] * nt, count)
    steps = np.empty((count, nt))
    for k in range(count):
        for n in range(nt):
            steps[k, n] = adm.step_from_magnitudes(*bnd.magnitudes([g_b[gp][k, n, :] for gp in bnd.grid], hb))
    Hs = np.empty((count, nt, nt))
    M = len(grid) // 2
    for k in range(count):
        vals = {0.0: g_b[0.0][k]}
        for m_ in range(1, M + 1):
            lo, hi = scores([(0.0 + steps[k, n] * -float(m_), 0.0 + steps[k, n] * float(m_)) for n in range(nt)], k + 1)
            vals[-float(m_)], vals[float(m_)] = lo[k], hi[k]
        for n in range(nt):
            Hs[k, :, n] = adm.estimate([vals[gp][n, :] for gp in grid], steps[k, n], coefs)
    return Hs, status, steps


# =============================================================================== finalize_result!
def finalize_result_(result: MuseResult, prob: AbstractMuseProblem):
    """src/muse.jl:535-549."""
    H, J, theta = result.H, result.J, result.theta
    if H is not None and J is not None and theta is not None:
        H_prior = -np.asarray(prob.prior.hess(theta), dtype=np.float64)            # :539
        result.Sigma_inv = H.T @ np.linalg.inv(J) @ H + H_prior                    # :540
        result.Sigma = np.linalg.inv(result.Sigma_inv)                             # :541
        result.dist = (np.array(theta, copy=True), result.Sigma.copy())            # :542-546
    return result
