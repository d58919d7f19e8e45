"""Hager–Zhang line search (oracle side; TEST INFRASTRUCTURE ONLY).

[EXT] Restates ``LineSearches.HagerZhang`` (LineSearches.jl ``src/hagerzhang.jl``, the
default line search of ``Optim.LBFGS()``), which the reference reaches through
``Optim.optimize(..., Optim.LBFGS(), ...)`` at /root/reference/src/interface.jl:163.
LineSearches.jl is not vendored under /root/reference (transitive dependency of
Optim "1.5", unpinned).  Target of this restatement: **LineSearches.jl v7.2.0** (what Optim v1.7–v1.9 —
the releases the compat range "1.5" resolved to while MuseInference v0.2.4 was current, Julia 1.7–1.10 — depend on
through `LineSearches = "7.0.1"`); ``hagerzhang.jl`` has not changed its arithmetic since v7.0 (the B0–B3 /
``secant2!`` / ``update!`` / ``bisect!`` structure, the ``nextfloat(values[ia]) >= values[ib]`` stagnation exit,
``psi3``, ``iterfinitemax = -log2(eps)``).  PARITY UNPINNED: restated from the published source as recalled, no
reference-held vector exists to check it against.  This is the published algorithm of

    W. W. Hager and H. Zhang, "Algorithm 851: CG_DESCENT", ACM TOMS 32 (2006),
    stages B0-B3 (bracket), S1-S4 (secant²), U0-U3 (update / bisect)

with LineSearches.jl's defaults: delta=0.1, sigma=0.9, alphamax=Inf, rho=5.0,
epsilon=1e-6, gamma=0.66, linesearchmax=50, psi3=0.1, mayterminate=false.

``mayterminate`` is only ever set by ``InitialHagerZhang``; Optim's default
``InitialStatic`` leaves it false, so the first trial point is never accepted without
entering the secant² stage.  Kept as a field to make that explicit.

The structure (growing ``alphas/values/slopes`` lists addressed by index) follows the
upstream routine one-to-one so the CUDA state machine, which keeps O(1) state, is
checked against an independently shaped implementation.
"""
from __future__ import annotations

import math

import numpy as np

EPS = float(np.finfo(np.float64).eps)


def _eps_of(x: float) -> float:
    """Julia ``eps(x::Float64)``: distance to the next float above ``abs(x)``."""
    ax = abs(x)
    return float(np.nextafter(ax, np.inf) - ax)


def _nextfloat(x: float) -> float:
    return float(np.nextafter(x, np.inf))


class LineSearchException(Exception):
    def __init__(self, message: str, alpha: float):
        super().__init__(message)
        self.alpha = alpha


class HagerZhang:
    def __init__(self, delta=0.1, sigma=0.9, alphamax=math.inf, rho=5.0, epsilon=1e-6,
                 gamma=0.66, linesearchmax=50, psi3=0.1):
        self.delta = delta
        self.sigma = sigma
        self.alphamax = alphamax
        self.rho = rho
        self.epsilon = epsilon
        self.gamma = gamma
        self.linesearchmax = linesearchmax
        self.psi3 = psi3
        self.mayterminate = False

    # ------------------------------------------------------------------ main entry
    def __call__(self, phidphi, c: float, phi_0: float, dphi_0: float):
        """Return ``(alpha, phi(alpha))``.  ``phidphi(alpha) -> (phi, dphi)``."""
        delta, sigma, alphamax, rho = self.delta, self.sigma, self.alphamax, self.rho
        epsilon, gamma, linesearchmax, psi3 = self.epsilon, self.gamma, self.linesearchmax, self.psi3

        if not (math.isfinite(phi_0) and math.isfinite(dphi_0)):
            raise LineSearchException("Value and slope at step length = 0 must be finite.", 0.0)
        if dphi_0 >= EPS * abs(phi_0):
            raise LineSearchException("Search direction is not a direction of descent.", 0.0)
        elif dphi_0 >= 0:
            return 0.0, phi_0

        iterfinitemax = int(math.ceil(-math.log2(EPS)))
        alphas = [0.0]
        values = [phi_0]
        slopes = [dphi_0]

        phi_lim = phi_0 + epsilon * abs(phi_0)
        assert c >= 0
        if c <= EPS:
            return 0.0, phi_0
        assert math.isfinite(c) and c <= alphamax
        phi_c, dphi_c = phidphi(c)
        iterfinite = 1
        while not (math.isfinite(phi_c) and math.isfinite(dphi_c)) and iterfinite < iterfinitemax:
            self.mayterminate = False
            iterfinite += 1
            c *= psi3
            phi_c, dphi_c = phidphi(c)
        if not (math.isfinite(phi_c) and math.isfinite(dphi_c)):
            self.mayterminate = False
            return 0.0, phi_0
        alphas.append(c)
        values.append(phi_c)
        slopes.append(dphi_c)

        if self.mayterminate and _satisfies_wolfe(c, phi_c, dphi_c, phi_0, dphi_0, phi_lim, delta, sigma):
            self.mayterminate = False
            return c, phi_c

        # ---- initial bracketing (HZ stages B0-B3); indices are 0-based here
        isbracketed = False
        ia, ib = 0, 1
        it = 1
        cold = -1.0
        phi_cold = math.nan
        while (not isbracketed) and it < linesearchmax:
            if dphi_c >= 0.0:
                ib = len(alphas) - 1
                for i in range(ib - 1, -1, -1):
                    if values[i] <= phi_lim:
                        ia = i
                        break
                isbracketed = True
            elif values[-1] > phi_lim:
                ib = len(alphas) - 1
                ia = 0
                ia, ib = _bisect(phidphi, alphas, values, slopes, ia, ib, phi_lim)
                isbracketed = True
            else:
                cold = c
                phi_cold = phi_c
                if _nextfloat(cold) >= alphamax:
                    self.mayterminate = False
                    return cold, phi_cold
                c *= rho
                if c > alphamax:
                    c = alphamax
                phi_c, dphi_c = phidphi(c)
                iterfinite = 1
                while (not (math.isfinite(phi_c) and math.isfinite(dphi_c))
                       and c > _nextfloat(cold) and iterfinite < iterfinitemax):
                    alphamax = c
                    iterfinite += 1
                    c = (cold + c) / 2
                    phi_c, dphi_c = phidphi(c)
                if not (math.isfinite(phi_c) and math.isfinite(dphi_c)):
                    return cold, phi_cold
                alphas.append(c)
                values.append(phi_c)
                slopes.append(dphi_c)
            it += 1

        # ---- secant² / bisection main loop
        while it < linesearchmax:
            a = alphas[ia]
            b = alphas[ib]
            assert b > a
            if b - a <= _eps_of(b):
                self.mayterminate = False
                return a, values[ia]
            iswolfe, iA, iB = _secant2(phidphi, alphas, values, slopes, ia, ib, phi_lim, delta, sigma)
            if iswolfe:
                self.mayterminate = False
                return alphas[iA], values[iA]
            A = alphas[iA]
            B = alphas[iB]
            assert B > A
            if B - A < gamma * (b - a):
                if _nextfloat(values[ia]) >= values[ib] and _nextfloat(values[iA]) >= values[iB]:
                    self.mayterminate = False
                    return A, values[iA]
                ia, ib = iA, iB
            else:
                c = (A + B) / 2.0
                phi_c, dphi_c = phidphi(c)
                assert math.isfinite(phi_c) and math.isfinite(dphi_c)
                alphas.append(c)
                values.append(phi_c)
                slopes.append(dphi_c)
                ia, ib = _update(phidphi, alphas, values, slopes, iA, iB, len(alphas) - 1, phi_lim)
            it += 1

        raise LineSearchException(
            f"Linesearch failed to converge, reached maximum iterations {linesearchmax}.", alphas[ia])


def _div(x: float, y: float) -> float:
    """IEEE-754 division (Julia and CUDA semantics): x/0 is ±Inf or NaN, never an exception."""
    if y == 0.0:
        if x == 0.0 or x != x:
            return math.nan
        return math.copysign(math.inf, x) * math.copysign(1.0, y)
    return x / y


def _satisfies_wolfe(c, phi_c, dphi_c, phi_0, dphi_0, phi_lim, delta, sigma) -> bool:
    wolfe1 = (delta * dphi_0 >= _div(phi_c - phi_0, c)) and (dphi_c >= sigma * dphi_0)
    wolfe2 = ((2 * delta - 1) * dphi_0 >= dphi_c >= sigma * dphi_0) and (phi_c <= phi_lim)
    return wolfe1 or wolfe2


def _secant(a, b, dphi_a, dphi_b):
    return _div(a * dphi_b - b * dphi_a, dphi_b - dphi_a)


def _secant2(phidphi, alphas, values, slopes, ia, ib, phi_lim, delta, sigma):
    phi_0 = values[0]
    dphi_0 = slopes[0]
    a, b = alphas[ia], alphas[ib]
    dphi_a, dphi_b = slopes[ia], slopes[ib]
    if not (dphi_a < 0.0 and dphi_b >= 0.0):
        raise RuntimeError("Search direction is not a direction of descent; "
                           f"(dphi_a = {dphi_a}; dphi_b = {dphi_b})")
    c = _secant(a, b, dphi_a, dphi_b)
    assert math.isfinite(c)
    phi_c, dphi_c = phidphi(c)
    assert math.isfinite(phi_c) and math.isfinite(dphi_c)
    alphas.append(c)
    values.append(phi_c)
    slopes.append(dphi_c)
    ic = len(alphas) - 1
    if _satisfies_wolfe(c, phi_c, dphi_c, phi_0, dphi_0, phi_lim, delta, sigma):
        return True, ic, ic
    iA, iB = _update(phidphi, alphas, values, slopes, ia, ib, ic, phi_lim)
    a = alphas[iA]
    b = alphas[iB]
    if iB == ic:
        c = _secant(alphas[ib], alphas[iB], slopes[ib], slopes[iB])
    elif iA == ic:
        c = _secant(alphas[ia], alphas[iA], slopes[ia], slopes[iA])
    if (iA == ic or iB == ic) and (a <= c <= b):
        phi_c, dphi_c = phidphi(c)
        assert math.isfinite(phi_c) and math.isfinite(dphi_c)
        alphas.append(c)
        values.append(phi_c)
        slopes.append(dphi_c)
        ic = len(alphas) - 1
        if _satisfies_wolfe(c, phi_c, dphi_c, phi_0, dphi_0, phi_lim, delta, sigma):
            return True, ic, ic
        iA, iB = _update(phidphi, alphas, values, slopes, iA, iB, ic, phi_lim)
    return False, iA, iB


def _update(phidphi, alphas, values, slopes, ia, ib, ic, phi_lim):
    a, b = alphas[ia], alphas[ib]
    assert slopes[ia] < 0.0
    assert values[ia] <= phi_lim
    assert slopes[ib] >= 0.0
    assert b > a
    c = alphas[ic]
    phi_c = values[ic]
    dphi_c = slopes[ic]
    if c < a or c > b:
        return ia, ib
    if dphi_c >= 0.0:
        return ia, ic
    if phi_c <= phi_lim:
        return ic, ib
    return _bisect(phidphi, alphas, values, slopes, ia, ic, phi_lim)


def _bisect(phidphi, alphas, values, slopes, ia, ib, phi_lim):
    a, b = alphas[ia], alphas[ib]
    assert slopes[ia] < 0.0
    assert values[ia] <= phi_lim
    assert slopes[ib] < 0.0
    assert values[ib] > phi_lim
    assert b > a
    while b - a > _eps_of(b):
        d = (a + b) / 2.0
        phi_d, gphi = phidphi(d)
        assert math.isfinite(phi_d) and math.isfinite(gphi)
        alphas.append(d)
        values.append(phi_d)
        slopes.append(gphi)
        idd = len(alphas) - 1
        if gphi >= 0.0:
            return ia, idd
        if phi_d <= phi_lim:
            a = d
            ia = idd
        else:
            b = d
            ib = idd
    return ia, ib
