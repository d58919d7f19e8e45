"""L-BFGS (oracle side; TEST INFRASTRUCTURE ONLY).

[EXT] Restates what the reference's default ``ẑ_at_θ`` runs
(/root/reference/src/interface.jl:162-166):

    Optim.optimize(Optim.only_fg(z -> .-logLike_and_∇z_logLike(prob, x, z, θ)), z₀,
                   Optim.LBFGS(), Optim.Options(g_tol = ∇z_logLike_atol))

Optim.jl is a dependency with compat "1.5" (Project.toml:45), not vendored under
/root/reference.  Target of this restatement: **Optim.jl v1.7.8** (with NLSolversBase v7.8.3 and
LineSearches v7.2.0) — any release from v1.5.0 to v1.9.x satisfies the compat range, and the L-BFGS path restated
here (``twoloop!`` with ``pseudo_iteration``, ``scaleinvH0``, ``reset_search_direction!`` when dφ₀ ≥ 0,
``update_h!`` skipping the pair when 1/(dx·dg) is infinite, ``assess_convergence`` with ``successive_f_tol = 1``,
``g_abstol`` tested at the initial point) is the same in all of them.  PARITY UNPINNED: restated from the published
source as recalled; the reference ships no vector that would pin it.  The one upstream known answer at hand — the run printed
in Optim's manual, ``optimize(f, [0.0, 0.0], LBFGS())`` on Rosenbrock: 24 iterations, 67 f and ∇f calls, final objective
5.3784…e-17 (quoted from memory) — is reproduced exactly (tests/test_oracle.py).  Restated from:

  * ``optimize`` main loop           src/multivariate/optimize/optimize.jl
  * ``LBFGS`` state / twoloop! / update_state! / update_h! / reset_search_direction!
                                     src/multivariate/solvers/first_order/l_bfgs.jl
  * ``perform_linesearch!``          src/utilities/perform_linesearch.jl
  * ``assess_convergence``           src/utilities/assess_convergence.jl
  * objective caching                NLSolversBase ``value_gradient!`` (re-evaluates only when
                                     the point differs from the last one evaluated)

Defaults that matter: ``LBFGS(m=10, alphaguess=InitialStatic(alpha=1), linesearch=HagerZhang(),
scaleinvH0=true)``; ``Options(x_abstol=0, x_reltol=0, f_abstol=0, f_reltol=0, g_abstol=g_tol,
iterations=1000, allow_f_increases=true, successive_f_tol=1)``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .hagerzhang import HagerZhang, LineSearchException


@dataclass
class OptimResult:
    minimizer: np.ndarray
    minimum: float
    iterations: int
    f_calls: int
    g_residual: float
    x_converged: bool
    f_converged: bool
    g_converged: bool
    ls_failed: bool = False
    alphas: list = field(default_factory=list)   # accepted step per iteration (diagnostic)

    @property
    def converged(self) -> bool:  # Optim.converged
        return self.x_converged or self.f_converged or self.g_converged


class _CachedObjective:
    """NLSolversBase-style objective: remembers the last point evaluated."""

    def __init__(self, fg):
        self._fg = fg
        self.x_last = None
        self.f = math.nan
        self.g = None
        self.f_calls = 0

    def value_gradient(self, x):
        if self.x_last is None or not np.array_equal(x, self.x_last):
            f, g = self._fg(x)
            self.f, self.g = float(f), np.array(g, dtype=np.float64, copy=True)
            self.x_last = np.array(x, copy=True)
            self.f_calls += 1
        return self.f, self.g


def _twoloop(gr, rho, dx_hist, dg_hist, m, pseudo_iteration, scaleinvH0=True):
    lower = pseudo_iteration - m
    upper = pseudo_iteration - 1
    q = gr.copy()
    alpha = np.zeros(m)
    for index in range(upper, lower - 1, -1):
        if index < 1:
            continue
        i = (index - 1) % m            # mod1(index, m) - 1
        alpha[i] = rho[i] * np.dot(dx_hist[i], q)
        q = q - alpha[i] * dg_hist[i]
    if scaleinvH0 and pseudo_iteration > 1:
        i = (upper - 1) % m
        scaling = np.dot(dx_hist[i], dg_hist[i]) / np.dot(dg_hist[i], dg_hist[i])
        s = scaling * q
    else:
        s = q.copy()
    for index in range(lower, upper + 1):
        if index < 1:
            continue
        i = (index - 1) % m
        beta = rho[i] * np.dot(dg_hist[i], s)
        s = s + dx_hist[i] * (alpha[i] - beta)
    return -s


def lbfgs_minimize(fg, x0, g_tol, m=10, iterations=1000, successive_f_tol=1,
                   linesearch: HagerZhang | None = None) -> OptimResult:
    """Minimise ``fg(x) -> (f, grad)`` from ``x0`` until ``max|grad| <= g_tol``."""
    ls = linesearch or HagerZhang()
    d = _CachedObjective(fg)
    x = np.array(x0, dtype=np.float64, copy=True)
    n = x.size

    # initial_state: value_gradient!!(d, initial_x)
    f_x, g = d.value_gradient(x)
    g = g.copy()
    x_previous = x.copy()
    g_previous = g.copy()
    rho = np.full(m, np.nan)
    dx_hist = [np.full(n, np.nan) for _ in range(m)]
    dg_hist = [np.full(n, np.nan) for _ in range(m)]
    f_x_previous = math.nan
    pseudo_iteration = 0
    alphas = []

    # initial_convergence
    stopped = (not math.isfinite(f_x)) or (not np.all(np.isfinite(g)))
    g_converged = (np.max(np.abs(g)) if n else 0.0) <= g_tol
    x_converged = f_converged = False
    counter_f_tol = 0
    converged = g_converged
    ls_failed = False
    iteration = 0

    while (not converged) and (not stopped) and iteration < iterations:
        iteration += 1

        # ---- update_state!
        pseudo_iteration += 1
        s = _twoloop(g, rho, dx_hist, dg_hist, m, pseudo_iteration)
        g_previous = g.copy()

        # perform_linesearch!
        dphi_0 = float(np.dot(g, s))
        if dphi_0 >= 0.0:
            pseudo_iteration = 1                  # reset_search_direction!
            s = -g
            dphi_0 = float(np.dot(g, s))
        phi_0 = f_x
        alpha = 1.0                               # InitialStatic(alpha=1.0, scaled=false)
        f_x_previous = phi_0
        x_previous = x.copy()

        def phidphi(a, _x=x, _s=s):
            fa, ga = d.value_gradient(_x + a * _s)
            return fa, float(np.dot(ga, _s))

        try:
            alpha, _ = ls(phidphi, alpha, phi_0, dphi_0)
            lssuccess = True
        except LineSearchException as ex:
            alpha = ex.alpha
            lssuccess = False
        alphas.append(alpha)
        dx = alpha * s
        x = x + dx
        if not lssuccess:
            ls_failed = True
            break

        # ---- update_g!
        f_x, g = d.value_gradient(x)
        g = g.copy()

        # ---- assess_convergence
        x_abschange = float(np.max(np.abs(x - x_previous))) if n else 0.0
        x_converged = x_abschange <= 0.0
        f_abschange = abs(f_x - f_x_previous)
        f_converged = f_abschange <= 0.0
        g_residual = float(np.max(np.abs(g))) if n else 0.0
        g_converged = g_residual <= g_tol
        counter_f_tol = counter_f_tol + 1 if f_converged else 0
        converged = x_converged or g_converged or (counter_f_tol > successive_f_tol)

        # ---- update_h!
        dg = g - g_previous
        dxdg = float(np.dot(dx, dg))
        rho_iteration = math.inf if dxdg == 0.0 else 1.0 / dxdg
        if math.isinf(rho_iteration):
            pseudo_iteration = 0
        else:
            idx = (pseudo_iteration - 1) % m
            dx_hist[idx] = dx.copy()
            dg_hist[idx] = dg.copy()
            rho[idx] = rho_iteration

        if (not math.isfinite(f_x)) or (not np.all(np.isfinite(g))):
            break

    return OptimResult(
        minimizer=x, minimum=f_x, iterations=iteration, f_calls=d.f_calls,
        g_residual=float(np.max(np.abs(g))) if n else 0.0,
        x_converged=x_converged, f_converged=f_converged, g_converged=g_converged,
        ls_failed=ls_failed, alphas=alphas,
    )
