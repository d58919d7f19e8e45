"""Model of the device algorithm for F3 (lock-step L-BFGS with ONE P·s product per iteration; line-search trials in
closed form along the line, quadratic objective) against the oracle's honest L-BFGS."""
import math, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
from oracle.lbfgs import _twoloop
from oracle.hagerzhang import HagerZhang, LineSearchException


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


def main():
    rng = np.random.default_rng(1)
    for d in (64, 256, 1024):
        A = rng.standard_normal((d, d)); S0 = A @ A.T / d + 0.1 * np.eye(d)
        P = np.linalg.inv(S0); L = np.linalg.cholesky(S0)
        fam = O.CorrGauss(d, P, L)
        for th in (1.0, 0.0, -1.0):
            a = math.exp(-th)
            worst = 0; its = []
            for k in range(6):
                x, ztrue = fam.sample([th], rng.standard_normal(d), rng.standard_normal(d))
                for z0 in (np.zeros(d), ztrue):
                    ref = O.lbfgs_minimize(lambda z: fam.neg_loglike_and_grad(x, z, [th]), z0, g_tol=1e-2)
                    z, it, fc, gres = lockstep(P, a, d * th, x, z0, 1e-2)
                    rel = np.max(np.abs(z - ref.minimizer)) / np.max(np.abs(ref.minimizer))
                    worst = max(worst, rel)
                    its.append((ref.iterations, it, ref.f_calls, fc))
            print(d, th, "worst rel diff %.2e" % worst, "iters/fcalls (ref, model):", its[:4])

main()
