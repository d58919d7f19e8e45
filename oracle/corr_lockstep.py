"""Model of the device algorithm for the correlated-Gaussian family (TEST INFRASTRUCTURE ONLY).

csrc/muse_corr.cu advances all units of a pass in lock-step and, because the objective is quadratic, evaluates
every Hager–Zhang trial of an iteration in closed form from ONE product P·s:
    φ(α) = f + α g·s + ½α² sᵀ(I + aP)s,   φ′(α) = g·s + α sᵀ(I + aP)s,   ∇f(z + αs) = g + α(s + a P s).
This module restates that scheme in NumPy so that a CPU test can hold it against the oracle's honest L-BFGS
(oracle/lbfgs.py: P·z evaluated at every trial point, as the reference's AD-based ``logLike_and_∇z_logLike`` would,
/root/reference/src/simple.jl:85): same iteration and evaluation counts, same ẑ to round-off."""
from __future__ import annotations

import numpy as np

from .hagerzhang import HagerZhang, LineSearchException
from .lbfgs import _twoloop


def lockstep(P, a, dth, x, z0, g_tol, m=10, iterations=1000):
    d = x.size
    z = z0.copy()
    Pz = P @ z if np.any(z) else np.zeros(d)
    g = (z - x) + a * Pz
    f = 0.5 * (np.dot(x - z, x - z) + a * np.dot(z, Pz) + dth)
    rho = np.full(m, np.nan); dxh = [None] * m; dgh = [None] * m
    pseudo = 0; it = 0; fcalls = 1; counter = 0
    converged = np.max(np.abs(g)) <= g_tol
    ls = HagerZhang()
    while not converged and it < iterations:
        it += 1; pseudo += 1
        s = _twoloop(g, rho, dxh, dgh, m, pseudo)
        dphi0 = float(np.dot(g, s))
        if dphi0 >= 0:
            pseudo = 1; s = -g; dphi0 = float(np.dot(g, s))
        q = P @ s                                    # the one product of the iteration
        sAs = float(np.dot(s, s) + a * np.dot(s, q))
        seen = {}
        def phidphi(al):
            if al not in seen:
                seen[al] = 1
            return f + al * dphi0 + 0.5 * al * al * sAs, dphi0 + al * sAs
        try:
            alpha, _ = ls(phidphi, 1.0, f, dphi0)
        except LineSearchException as ex:
            alpha = ex.alpha; break
        # count evaluations like NLSolversBase caching: distinct points, + the final point if not the last evaluated
        last = list(seen)[-1] if seen else None
        fcalls += len(seen) + (0 if last == alpha else 1)
        f_prev = f
        dx = alpha * s
        dg = alpha * (s + a * q)
        z = z + dx
        g = g + dg
        f = f + alpha * dphi0 + 0.5 * alpha * alpha * sAs
        xconv = np.max(np.abs(dx)) <= 0
        fconv = abs(f - f_prev) <= 0
        gconv = np.max(np.abs(g)) <= g_tol
        counter = counter + 1 if fconv else 0
        converged = xconv or gconv or counter > 1
        dxdg = float(np.dot(dx, dg))
        if dxdg == 0:
            pseudo = 0
        else:
            idx = (pseudo - 1) % m
            dxh[idx], dgh[idx], rho[idx] = dx, dg, 1.0 / dxdg
    return z, it, fcalls, float(np.max(np.abs(g)))


