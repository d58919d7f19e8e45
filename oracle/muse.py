"""MUSE outer solver + covariance (oracle side; TEST INFRASTRUCTURE ONLY).

Line-by-line NumPy restatement of

    muse!             /root/reference/src/muse.jl:112-250
    get_H! (FD path)  /root/reference/src/muse.jl:296-333, 407-450
    get_J!            /root/reference/src/muse.jl:484-532
    finalize_result!  /root/reference/src/muse.jl:535-549
    pjacobian         /root/reference/src/util.jl:9-26
    split_rng         /root/reference/src/util.jl:85-92   (semantics: sim k ↔ fixed base normals)

Differences from the reference that are *representation only*:
  * θ is always a 1-D float64 array (length nθ); "θ isa Number" branches of the reference
    (src/muse.jl:189, 446, 529) are selected by ``nθ == 1`` and give identical numbers.
  * RNG objects are replaced by ``Draws`` (base normals ξ_k, ν_k per sim plus the master
    stream's own draw), see oracle/__init__.py.
  * Prior gradient / Hessian come from a prior object with analytic derivatives instead of
    ForwardDiff (src/muse.jl:184, 207, 539).
  * θ-transforms are the identity for the plain families (true of SimpleMuseProblem,
    src/interface.jl:20, 28), so primed and unprimed quantities coincide; a ``TransformedFamily``
    (oracle/families.py) supplies ``transform_θ`` / ``inv_transform_θ`` and the score in both spaces,
    and the primed (transformed-space) quantities of src/muse.jl:136, 164, 173, 183-227 become distinct:
    history keys ``theta_t``, ``theta_unreg_t``, ``g_like_sims_t`` hold θ′, θunreg′, g_like_sims′; the
    keys g_like_dat, g_like, g_prior, g_post, H_* hold the primed quantities (the reference stores only
    those), ``theta``, ``theta_unreg``, ``g_like_sims`` the unprimed ones.

Quirks reproduced on purpose:
  * convergence test ``sqrt(-(Δθ' H⁻¹_post Δθ)) < θ_rtol`` uses the *inverse* Hessian (:165);
  * ``result.gs`` are the scores at the θ *before* the last update (:231);
  * ``get_J!`` after ``muse!`` with the same nsims runs no new sims (:499-502);
  * ``get_H!``'s fiducial solves all use the master stream's own draw because the closure
    argument shadows nothing (``rngs`` vs ``rng``, :417-418), start from ``zero(z)`` (:419,
    src/interface.jl:184-186) and only serve as start points;
  * ``get_H!`` re-splits from the start on resume (:323), unlike ``get_J!`` (:506);
  * central_fdm(3,1) with an explicit step evaluates the centre point and multiplies it by 0.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

from .lbfgs import lbfgs_minimize
from .philox import philox_normals, MASTER_INDEX


# ----------------------------------------------------------------------------- priors
class FlatPrior:
    """logPriorθ ≡ 0 (reference default, src/interface.jl:121)."""

    def logp(self, theta):
        return 0.0

    def grad(self, theta):
        return np.zeros_like(np.asarray(theta, dtype=np.float64))

    def hess(self, theta):
        n = np.asarray(theta).size
        return np.zeros((n, n))


class NormalPrior:
    """Independent N(mean, sigma) per component; the funnel example uses N(0, 3)
    (src/simple.jl:69-71: ``-θ^2/(2*3^2)``)."""

    def __init__(self, mean=0.0, sigma=3.0):
        self.mean = mean
        self.sigma = sigma

    def logp(self, theta):
        t = np.asarray(theta, dtype=np.float64)
        return float(-np.sum((t - self.mean) ** 2 / (2 * np.asarray(self.sigma) ** 2)))

    def grad(self, theta):
        t = np.asarray(theta, dtype=np.float64)
        return -(t - self.mean) / np.asarray(self.sigma) ** 2

    def hess(self, theta):
        t = np.asarray(theta, dtype=np.float64)
        return np.diag(np.broadcast_to(-1.0 / np.asarray(self.sigma, dtype=np.float64) ** 2, t.shape).copy())


# ----------------------------------------------------------------------------- draws
@dataclass
class Draws:
    """Base normals: row k of ``xi``/``nu`` belongs to child stream k (0-based);
    ``xi_master``/``nu_master`` is the master stream's own draw."""

    xi: np.ndarray
    nu: np.ndarray
    xi_master: np.ndarray
    nu_master: np.ndarray

    @property
    def nsims(self):
        return self.xi.shape[0]

    @staticmethod
    def from_philox(seed: int, nsims: int, d: int, offset: int = 0) -> "Draws":
        xi = np.stack([philox_normals(seed, offset + k, 0, d) for k in range(nsims)]) if nsims else np.zeros((0, d))
        nu = np.stack([philox_normals(seed, offset + k, 1, d) for k in range(nsims)]) if nsims else np.zeros((0, d))
        return Draws(xi, nu, philox_normals(seed, MASTER_INDEX, 0, d), philox_normals(seed, MASTER_INDEX, 1, d))

    @staticmethod
    def from_numpy(seed: int, nsims: int, d: int) -> "Draws":
        rng = np.random.Generator(np.random.Philox(seed))
        xi = rng.standard_normal((nsims, d))
        nu = rng.standard_normal((nsims, d))
        return Draws(xi, nu, rng.standard_normal(d), rng.standard_normal(d))


# ----------------------------------------------------------------------------- problem
class OracleProblem:
    """Plays the role of ``SimpleMuseProblem`` (src/simple.jl:4-12) for a registered family."""

    def __init__(self, family, x, draws: Draws, prior=None):
        self.family = family
        self.x = np.asarray(x, dtype=np.float64)
        self.draws = draws
        self.prior = prior or FlatPrior()

    # sample_x_z(prob, rng_k, θ)  — src/simple.jl:95
    def sample_x_z(self, k, theta):
        if k == "master":
            return self.family.sample(theta, self.draws.xi_master, self.draws.nu_master)
        return self.family.sample(theta, self.draws.xi[k], self.draws.nu[k])

    # ẑ_at_θ  — src/interface.jl:162-166
    def z_at_theta(self, x, z0, theta, atol):
        soln = lbfgs_minimize(lambda z: self.family.neg_loglike_and_grad(x, z, theta), z0, g_tol=atol)
        return soln.minimizer, soln

    # ∇θ_logLike(prob, x, z, θ, UnTransformedθ())  — src/simple.jl:92, src/interface.jl:57-58
    def grad_theta(self, x, z, theta):
        return self.family.score(x, z, theta)

    # ∇θ_logLike(prob, x, z, θ′, Transformedθ())  — src/interface.jl:57-58 (identity transform: the same function)
    def grad_theta_t(self, x, z, theta_t):
        f = getattr(self.family, "score_t", None)
        return f(x, z, theta_t) if f else self.family.score(x, z, theta_t)

    # transform_θ / inv_transform_θ  — src/interface.jl:20, 28
    def transform_theta(self, theta):
        f = getattr(self.family, "transform_theta", None)
        return f(theta) if f else np.array(theta, dtype=np.float64, copy=True)

    def inv_transform_theta(self, theta_t):
        f = getattr(self.family, "inv_transform_theta", None)
        return f(theta_t) if f else np.array(theta_t, dtype=np.float64, copy=True)

    # logPriorθ(prob, θ′, Transformedθ()) = logPriorθ(prob, inv_transform_θ(θ′), UnTransformedθ())  — src/soss.jl:107-108;
    # its gradient and Hessian in θ′ (the reference: ForwardDiff, src/muse.jl:184, 207) by central differences of the
    # analytic untransformed gradient composed with the transform (identity transform: the analytic ones directly)
    def prior_grad_t(self, theta_t):
        if not hasattr(self.family, "transform_theta"):
            return self.prior.grad(theta_t)
        D = self.family.dinv_transform(theta_t)
        return D * self.prior.grad(self.inv_transform_theta(theta_t))

    def prior_hess_t(self, theta_t):
        if not hasattr(self.family, "transform_theta"):
            return self.prior.hess(theta_t)
        tt = np.asarray(theta_t, dtype=np.float64)
        n = tt.size
        H = np.empty((n, n))
        for j in range(n):                      # column j of the Hessian = ∂(∇logπ′)/∂θ′_j, Richardson-extrapolated
            def gcol(h):
                e = np.zeros(n); e[j] = h
                return (self.prior_grad_t(tt + e) - self.prior_grad_t(tt - e)) / (2 * h)
            H[:, j] = (4.0 * gcol(5e-4) - gcol(1e-3)) / 3.0
        return 0.5 * (H + H.T)

    # ẑ_guess_from_truth — src/interface.jl:184-186
    def z_guess_from_truth(self, x, z, theta):
        return np.zeros_like(z)


def map_score_unit(prob: OracleProblem, x, z0, theta, atol):
    """Body of the mapped block src/muse.jl:170-175 for one unit."""
    zhat, soln = prob.z_at_theta(x, z0, theta, atol)
    g = prob.grad_theta(x, zhat, theta)
    return zhat, g, soln


# ----------------------------------------------------------------------------- result
@dataclass
class MuseResult:
    """src/muse.jl:29-42."""

    theta: Optional[np.ndarray] = None
    H: Optional[np.ndarray] = None
    J: Optional[np.ndarray] = None
    Sigma_inv: Optional[np.ndarray] = None
    Sigma: Optional[np.ndarray] = None
    dist: Optional[tuple] = None        # (mean vector, covariance matrix)
    history: list = field(default_factory=list)
    gs: list = field(default_factory=list)
    Hs: list = field(default_factory=list)
    metadata: dict = field(default_factory=dict)
    time: float = 0.0


def _var_corrected(rows: np.ndarray) -> np.ndarray:
    """Elementwise ``var`` with N-1 (Statistics.var of a vector of vectors)."""
    return np.var(rows, axis=0, ddof=1)


# ----------------------------------------------------------------------------- muse!
def muse(prob, theta0, **kw):
    """src/muse.jl:107."""
    return muse_bang(MuseResult(), prob, theta0, **kw)


def muse_bang(result: MuseResult, prob: OracleProblem, theta0=None, *, z0=None, maxsteps=50,
              theta_rtol=1e-1, gradz_logLike_atol=1e-2, nsims=100, alpha=0.7,
              regularize: Callable = lambda t: t, H_inv_like=None, H_inv_update="sims",
              broyden_memory=math.inf, get_covariance=False, save_MAPs=False):
    # :135-136
    theta = np.atleast_1d(np.asarray(result.theta if result.theta is not None else theta0, dtype=np.float64)).copy()
    theta_unreg = theta.copy()
    theta_t = prob.transform_theta(theta)
    theta_unreg_t = theta_t.copy()
    history = result.history
    ntheta = theta.size
    alpha_fn = alpha if callable(alpha) else (lambda i, _a=alpha: _a)   # :145-149

    # :151  (one throw-away sample from the master stream, only for the shape)
    zshape = prob.sample_x_z("master", theta)[1]
    zs = [np.array(z0, dtype=np.float64, copy=True) if z0 is not None else np.zeros_like(zshape)
          for _ in range(nsims + 1)]

    for i in range(len(history) + 1, maxsteps + 1):                       # :159
        t0 = time.perf_counter()
        if i > 2:                                                         # :163-166
            dth = history[-1]["theta_t"] - history[-2]["theta_t"]
            q = -(dth @ history[-1]["H_inv_post"] @ dth)
            if q < 0:
                raise ValueError("sqrt of a negative number in the θ convergence test (DomainError in the reference)")
            if math.sqrt(q) < theta_rtol:
                break

        # MUSE gradient  :169-176  (unit 0 = data, units 1..nsims = sims)
        gs_all, gts_all, zs_new, hists = [], [], [], []
        for u in range(nsims + 1):
            x = prob.x if u == 0 else prob.sample_x_z(u - 1, theta)[0]
            zhat, g, soln = map_score_unit(prob, x, zs[u], theta, gradz_logLike_atol)      # :171-172
            gs_all.append(g)
            gts_all.append(prob.grad_theta_t(x, zhat, theta_t))           # :173
            zs_new.append(zhat)
            hists.append(soln)
        g_like_sims = np.array(gs_all[1:])                                # :177
        g_like_dat = gts_all[0]                                           # :178 (primed from here on)
        g_like_sims_t = np.array(gts_all[1:])
        zs = zs_new                                                       # :181

        g_like = g_like_dat - np.mean(g_like_sims_t, axis=0)              # :183
        g_prior = prob.prior_grad_t(theta_t)                              # :184
        g_post = g_like + g_prior                                         # :185

        h_inv_like_sims = -1.0 / _var_corrected(g_like_sims_t)            # :188
        H_inv_like_sims = np.diag(h_inv_like_sims)                        # :189
        if H_inv_like is None or H_inv_update == "sims":                  # :190-191
            H_inv_like = H_inv_like_sims
        elif i > 2 and H_inv_update in ("broyden", "diagonal_broyden"):   # :192-205
            j0 = int(max(2, i - broyden_memory))
            H_inv_like = history[j0 - 2]["H_inv_like_sims"]
            for j in range(j0, i):
                dth = history[j - 1]["theta_t"] - history[j - 2]["theta_t"]
                dgl = history[j - 1]["g_like"] - history[j - 2]["g_like"]
                H_inv_like = H_inv_like + np.outer((dth - H_inv_like @ dgl) / (dth @ H_inv_like @ dgl), dth) @ H_inv_like
                if H_inv_update == "diagonal_broyden":
                    H_inv_like = np.diag(np.diag(H_inv_like))

        H_prior = prob.prior_hess_t(theta_t)                              # :207
        H_inv_post = np.linalg.inv(np.linalg.inv(H_inv_like) + H_prior)   # :208

        t = time.perf_counter() - t0
        history.append(dict(                                              # :211-221
            theta=theta.copy(), theta_unreg=theta_unreg.copy(), theta_t=theta_t.copy(), theta_unreg_t=theta_unreg_t.copy(),
            g_like_sims=g_like_sims.copy(), g_like_sims_t=g_like_sims_t.copy(),
            g_like_dat=g_like_dat.copy(), g_like=g_like.copy(),
            g_prior=g_prior.copy(), g_post=g_post.copy(),
            H_inv_post=H_inv_post.copy(), H_prior=H_prior.copy(), H_inv_like=H_inv_like.copy(),
            H_inv_like_sims=H_inv_like_sims.copy(),
            z_history_dat=hists[0], z_history_sims=hists[1:], t=t,
            z_dat=zs[0].copy() if save_MAPs else None,
            z_sims=[z.copy() for z in zs[1:]] if save_MAPs else None,
        ))

        theta_unreg_t = theta_t - alpha_fn(i) * (H_inv_post @ g_post)     # :224
        theta_unreg = prob.inv_transform_theta(theta_unreg_t)             # :225
        theta_t = np.atleast_1d(np.asarray(regularize(theta_unreg_t), dtype=np.float64))   # :226
        theta = prob.inv_transform_theta(theta_t)                         # :227

        result.theta = theta_unreg.copy()                                 # :230
        result.gs = [g.copy() for g in g_like_sims]                       # :231
        result.time += t                                                  # :232

    if get_covariance:                                                    # :244-247
        get_J_bang(result, prob, nsims=nsims, gradz_logLike_atol=gradz_logLike_atol)
        get_H_bang(result, prob, nsims=max(1, nsims // 10), gradz_logLike_atol=gradz_logLike_atol)
    return result


# ----------------------------------------------------------------------------- get_J!
class SimpleCovariance:
    """[EXT CovarianceEstimation 0.2] ``SimpleCovariance(corrected=…)``: the sample covariance, normalised by n − 1 (corrected) or n."""

    def __init__(self, corrected=False):
        self.corrected = bool(corrected)

    def __call__(self, gs):
        return np.atleast_2d(np.cov(np.asarray(gs, dtype=np.float64), rowvar=False, ddof=1 if self.corrected else 0))


def get_J_bang(result: MuseResult, prob: OracleProblem, theta0=None, *, z0=None,
               gradz_logLike_atol=1e-2, nsims=100, covariance_method=None):
    theta0 = np.atleast_1d(np.asarray(theta0 if theta0 is not None else result.theta, dtype=np.float64))   # :498
    nsims_existing = len(result.gs)
    nsims_remaining = nsims - nsims_existing
    if nsims_remaining > 0:
        for k in range(nsims_existing, nsims):                            # :506 rngs[nsims_existing+1:end]
            x, z = prob.sample_x_z(k, theta0)                             # :510
            zstart = np.array(z0, dtype=np.float64) if z0 is not None else z   # :511
            _, g, _ = map_score_unit(prob, x, zstart, theta0, gradz_logLike_atol)   # :512-513
            result.gs.append(g)
    gs = np.array(result.gs)
    if theta0.size == 1:
        result.J = np.array([[np.var(gs[:, 0], ddof=1)]])                 # :529 var
    else:                                                                 # :529 cov(covariance_method, gs), default SimpleCovariance(corrected=true)
        result.J = (covariance_method or SimpleCovariance(corrected=True))(gs)
    finalize_result_bang(result, prob)
    return result


# ----------------------------------------------------------------------------- get_H!
def central_fdm(p, q=1):
    """[EXT FiniteDifferences 0.12] ``central_fdm(p, q)`` without adaptation: the symmetric integer grid of p points (odd p:
    −(p−1)/2 … (p−1)/2; even p: without 0) and the coefficients c with Σᵢ cᵢ gᵢᵏ = q!·δ_{kq}, k = 0 … p−1 — solved in exact
    rational arithmetic, as the package does, then rounded to Float64.  central_fdm(3,1) → grid [−1, 0, 1], coefs [−1/2, 0, 1/2];
    central_fdm(5,1) → [1/12, −2/3, 0, 2/3, −1/12]."""
    from fractions import Fraction
    from math import factorial
    if p < 2 or q != 1 and q >= p:
        raise ValueError("central_fdm: need p ≥ 2 points and q < p")
    grid = list(range(-(p // 2), p // 2 + 1)) if p % 2 else [g for g in range(-(p // 2), p // 2 + 1) if g != 0]
    A = [[Fraction(g) ** k for g in grid] + [Fraction(factorial(q) if k == q else 0)] for k in range(p)]
    for c in range(p):                                   # Gauss–Jordan over the rationals
        piv = next(r for r in range(c, p) if A[r][c] != 0)
        A[c], A[piv] = A[piv], A[c]
        A[c] = [v / A[c][c] for v in A[c]]
        for r in range(p):
            if r != c and A[r][c] != 0:
                A[r] = [vr - A[r][c] * vc for vr, vc in zip(A[r], A[c])]
    return tuple(float(g) for g in grid), tuple(float(A[k][p]) for k in range(p))


class AdaptedFDM:
    """[EXT FiniteDifferences 0.12] ``central_fdm(p, q; adapt = 1, condition = 10, factor = 1, max_range = Inf)`` called WITHOUT a step
    (what ``pjacobian`` does when ``step === nothing``, src/util.jl:13): the step is estimated per call by minimising a bound on
    round-off + truncation error,

        step = (q/(p−q) · C₁/C₂)^(1/p),   C₁ = eps(|f|)·Σ|c|·factor,   C₂ = |∇ᵖf|·Σ|c·gᵖ|/p!,

    where |∇ᵖf| and |f| come from the *bound estimator* — the unadapted ``central_fdm(p + 2, p)`` evaluated at ITS default step
    (the same formula with |∇ᵖf| → condition, eps(|f|) → eps(Float64)) — as the largest magnitude over the estimates at x − h, x, x + h
    (the same p + 2 function values with the grid shifted by ∓1) and over the components of a vector-valued f; the step is then capped
    at max_range / max|grid| and at 1000 × the default step.  Restated from the package source as recalled (src/methods.jl:
    ``estimate_step``, ``_estimate_magnitudes``, ``_compute_step_acc``, ``_limit_step``); the one published figure at hand —
    ``estimate_step(central_fdm(5, 1), sin, 1.0)`` = (0.001065235154086019, 1.9541865128909085e-13) in the package's documentation,
    quoted from memory — is reproduced to the last digit (tests/test_oracle.py); it is sensitive to the order of the floating-point
    sum Σ fᵢcᵢ (left to right, as Julia folds a static vector): a compensated sum already moves the sixth digit."""

    def __init__(self, p, q=1, adapt=1, condition=10.0, factor=1.0, max_range=math.inf):
        from fractions import Fraction
        self.p, self.q = int(p), int(q)
        self.grid, self.coefs = central_fdm(p, q)
        shift = lambda d: _fdm_coefs([g + d for g in self.grid], q)
        self.coefs_nbhd = (shift(-1), self.coefs, shift(+1))
        self.condition, self.factor, self.max_range = float(condition), float(factor), float(max_range)
        self.df_mult = float(sum(abs(Fraction(c) * Fraction(g) ** self.p) for c, g in zip(self.coefs, self.grid)) / math.factorial(self.p))
        self.ferr_mult = sum(abs(c) for c in self.coefs)
        self.bound = AdaptedFDM(p + 2, p, adapt - 1, condition, factor, max_range) if adapt >= 1 else None

    def _step_acc(self, df_magnitude, f_error):
        P, Q = self.p, self.q
        c1 = f_error * self.ferr_mult * self.factor
        c2 = df_magnitude * self.df_mult
        step = (Q / (P - Q) * (c1 / c2)) ** (1 / P)
        return step, c1 * step ** (-Q) + c2 * step ** (P - Q)

    def default_step(self):
        return self._step_acc(self.condition, float(np.finfo(np.float64).eps))

    def _limit(self, step, acc):
        step_max = self.max_range / max(abs(g) for g in self.grid)
        if step > step_max:
            step, acc = step_max, math.nan
        step_default, _ = self.default_step()
        if step > 1000 * step_default:
            step, acc = 1000 * step_default, math.nan
        return step, acc

    def evaluate(self, f, x, step):
        return [np.atleast_1d(np.asarray(f(x + step * g), dtype=np.float64)).copy() for g in self.grid]

    def estimate(self, fs, step, coefs=None):
        coefs = self.coefs if coefs is None else coefs
        acc = fs[0] * coefs[0]
        for fk, ck in zip(fs[1:], coefs[1:]):                             # sum(fs .* coefs): left to right
            acc = acc + fk * ck
        return acc / step ** self.q

    def magnitudes(self, fs, step):
        """(|∇^q f|, |f|) in a neighbourhood of x from the values ``fs`` at x + step·grid (the bound estimator's rôle)."""
        df = max(float(np.max(np.abs(self.estimate(fs, step, c)))) for c in self.coefs_nbhd)
        return df, max(float(np.max(np.abs(v))) for v in fs)

    def step_from_magnitudes(self, df_magnitude, f_magnitude):
        if df_magnitude == 0.0 or f_magnitude == 0.0:
            return self._limit(*self.default_step())
        return self._limit(*self._step_acc(df_magnitude, float(np.spacing(f_magnitude))))

    def estimate_step(self, f, x):
        if self.bound is None:
            return self._limit(*self.default_step())
        hb = self.bound.estimate_step(f, x)[0]
        return self.step_from_magnitudes(*self.bound.magnitudes(self.bound.evaluate(f, x, hb), hb))

    def __call__(self, f, x, step=None):
        step = self.estimate_step(f, x)[0] if step is None else step
        return self.estimate(self.evaluate(f, x, step), step)


def _fdm_coefs(grid, q):
    """Coefficients c with Σᵢ cᵢ gᵢᵏ = q!·δ_{kq}, k = 0 … p−1, for an arbitrary integer grid (exact rationals → Float64)."""
    from fractions import Fraction
    p = len(grid)
    A = [[Fraction(g) ** k for g in grid] + [Fraction(math.factorial(q) if k == q else 0)] for k in range(p)]
    for c in range(p):
        piv = next(r for r in range(c, p) if A[r][c] != 0)
        A[c], A[piv] = A[piv], A[c]
        A[c] = [v / A[c][c] for v in A[c]]
        for r in range(p):
            if r != c and A[r][c] != 0:
                A[r] = [vr - A[r][c] * vc for vr, vc in zip(A[r], A[c])]
    return tuple(float(A[k][p]) for k in range(p))


def pjacobian_adaptive(f, theta0, fdm=None):
    """src/util.jl:9-26 with ``step === nothing``: per component n, ``fdm(ε -> f(θ₀ + ε eₙ), 0.0)`` with FiniteDifferences' own
    step estimate.  Returns (Jacobian, steps)."""
    x = np.array(theta0, dtype=np.float64, copy=True)
    grid, _ = fdm if fdm is not None else ((-1.0, 0.0, 1.0), None)
    adm = AdaptedFDM(len(grid), 1)
    cols, steps = [], []
    for n in range(x.size):
        def fn(eps, _n=n):
            xx = x.copy()
            xx[_n] = x[_n] + eps
            return f(xx)
        h = adm.estimate_step(fn, 0.0)[0]
        cols.append(adm(fn, 0.0, h))
        steps.append(h)
    return np.stack(cols, axis=1), np.array(steps)


def pjacobian(f, theta0, step, fdm=None):
    """src/util.jl:9-26 with an explicit step per component; fdm = (grid, coefs), default central_fdm(3,1).
    [EXT FiniteDifferences 0.12] with an explicit step the estimate is ``sum(fs .* coefs) / step`` where
    fs = f.(0 .+ step .* grid) — every grid point is evaluated, also the centre one that central methods multiply by 0."""
    x = np.array(theta0, dtype=np.float64, copy=True)
    cols = []
    grid, coefs = fdm if fdm is not None else ((-1.0, 0.0, 1.0), (-0.5, 0.0, 0.5))
    for n in range(x.size):
        h = float(step[n])
        fs = []
        for gpt in grid:
            eps = 0.0 + h * gpt
            xn = x[n]
            x[n] = xn + eps
            fs.append(np.array(f(x.copy()), dtype=np.float64, copy=True))
            x[n] = xn
        acc = fs[0] * coefs[0]
        for fk, ck in zip(fs[1:], coefs[1:]):
            acc = acc + fk * ck
        cols.append(acc / h)
    return np.stack(cols, axis=1)


def cg(A, b, *, maxiter=100, reltol=None, abstol=0.0):
    """[EXT IterativeSolvers 0.9] ``cg(A, b; maxiter, Pl = I, log = true)`` from a zero start: plain conjugate gradients,
    stopped when ‖r‖ ≤ max(reltol·‖b‖, abstol), reltol = √eps.  Returns (x, iterations).  (A may be negative definite — the
    reference hands it the Hessian of logLike: the signs cancel in α = ρ / uᵀAu.)"""
    b = np.asarray(b, dtype=np.float64)
    reltol = np.sqrt(np.finfo(np.float64).eps) if reltol is None else reltol
    x = np.zeros_like(b)
    r = b.copy()
    u = np.zeros_like(b)
    rho_prev = 1.0
    residual = np.linalg.norm(r)
    tol = max(reltol * residual, abstol)
    it = 0
    while not residual <= tol and it < maxiter:
        rho = float(r @ r)
        beta = rho / rho_prev
        u = r + beta * u
        c = A(u)
        alpha = rho / float(u @ c)
        x = x + alpha * u
        r = r - alpha * c
        residual = np.linalg.norm(r)
        rho_prev = rho
        it += 1
    return x, it


def implicit_diff_H(prob: OracleProblem, k, theta0, z0=None, cg_maxiter=100):
    """One sim of the implicit-diff branch of get_H! (src/muse.jl:340-388): H = H1 + H2 with
    H1 = ∂θ_sim[∇θ′ logLike(x(θ_sim), ẑ, θ′)] (0 for the registered families: their score does not see x) and
    H2 = −(∂θ ∇z logLike)ᵀ · A⁻¹ · (∂θ_sim ∇z logLike), A = ∇²z logLike, solved column by column with conjugate gradients.
    ẑ is the MAP at θ₀ with ∇z_logLike_atol = 1e-1 (hard-coded in the reference, :346).  The nested-AD derivatives of the
    reference are the families' analytic ones (oracle/families.py; checked against finite differences in tests/test_oracle.py)."""
    fam = prob.family
    theta0 = np.atleast_1d(np.asarray(theta0, dtype=np.float64))
    xi, nu = (prob.draws.xi_master, prob.draws.nu_master) if k == "master" else (prob.draws.xi[k], prob.draws.nu[k])
    x, z = prob.sample_x_z(k, theta0)
    zstart = np.array(z0, dtype=np.float64) if z0 is not None else prob.z_guess_from_truth(x, z, theta0)
    zhat, _ = prob.z_at_theta(x, zstart, theta0, 1e-1)
    dF = fam.dgradz_dtheta(x, zhat, theta0)                       # d × nθ
    dF1 = fam.dx_dtheta_sim(theta0, xi, nu)                       # d × nθ
    cols, iters = [], []
    for n in range(theta0.size):
        u, it = cg(lambda w: fam.hess_z_apply(zhat, theta0, w), dF1[:, n], maxiter=cg_maxiter)
        cols.append(u)
        iters.append(it)
    H2 = -(dF.T @ np.stack(cols, axis=1))
    return H2, iters


def get_H_bang(result: MuseResult, prob: OracleProblem, theta0=None, *, gradz_logLike_atol=1e-2,
               nsims=10, step=None, z0=None, fdm=None, implicit_diff=False, implicit_diff_cg_kwargs=None):
    theta0 = np.atleast_1d(np.asarray(theta0 if theta0 is not None else result.theta, dtype=np.float64))   # :315
    nsims_existing = len(result.Hs)
    nsims_remaining = nsims - nsims_existing
    if nsims_remaining <= 0:
        return result
    t0 = time.perf_counter()
    ks = list(range(nsims_remaining))                                     # :323 split_rng(rng, nsims_remaining)

    if implicit_diff:                                                     # :335-405
        hists = result.metadata.setdefault("implicit_diff_cg_hists", [])
        for k in ks:
            H, iters = implicit_diff_H(prob, k, theta0, z0, **({"cg_maxiter": implicit_diff_cg_kwargs["maxiter"]} if implicit_diff_cg_kwargs and "maxiter" in implicit_diff_cg_kwargs else {}))
            result.Hs.append(H)
            hists.append(iters)
        result.H = np.mean(np.array(result.Hs), axis=0)                   # :446
        result.time += time.perf_counter() - t0
        finalize_result_bang(result, prob)
        return result

    if step is None and len(result.gs) > 0:                               # :411-413
        step = 0.1 / np.std(np.array(result.gs), axis=0, ddof=1)
    adaptive = step is None                                               # fdm(f, 0.0): FiniteDifferences estimates the step itself
    if not adaptive:
        step = np.atleast_1d(np.asarray(step, dtype=np.float64))

    # fiducial MAPs  :417-423  (every one is the MAP of the master stream's own draw)
    zfids = []
    for _ in ks:
        x, z = prob.sample_x_z("master", theta0)
        zstart = np.array(z0, dtype=np.float64) if z0 is not None else prob.z_guess_from_truth(x, z, theta0)
        zfid, _ = prob.z_at_theta(x, zstart, theta0, gradz_logLike_atol)
        zfids.append(zfid)

    # FD Jacobian per sim  :426-433
    for zfid, k in zip(zfids, ks):
        def f(theta, _k=k, _z=zfid):
            x, _ = prob.sample_x_z(_k, theta)                             # sim generated at θ
            zhat, _ = prob.z_at_theta(x, _z, theta0, gradz_logLike_atol)  # MAP at fiducial θ₀
            return prob.grad_theta(x, zhat, theta0)                       # score at fiducial θ₀
        if adaptive:
            Hk, steps_k = pjacobian_adaptive(f, theta0, fdm)
            result.metadata.setdefault("fd_adaptive_steps", []).append(steps_k)
            result.Hs.append(Hk)
        else:
            result.Hs.append(pjacobian(f, theta0, step, fdm))

    result.H = np.mean(np.array(result.Hs), axis=0)                       # :446
    result.time += time.perf_counter() - t0
    finalize_result_bang(result, prob)
    return result


# ----------------------------------------------------------------------------- finalize_result!
def finalize_result_bang(result: MuseResult, prob: OracleProblem):
    H, J, theta = result.H, result.J, result.theta
    if H is not None and J is not None and theta is not None:
        H_prior = -prob.prior.hess(theta)                                 # :539
        result.Sigma_inv = H.T @ np.linalg.inv(J) @ H + H_prior           # :540
        result.Sigma = np.linalg.inv(result.Sigma_inv)                    # :541
        result.dist = (theta.copy(), result.Sigma.copy())                 # :542-546
    return result
