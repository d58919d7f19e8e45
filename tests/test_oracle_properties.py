"""Property tests of the restated third-party algorithms (oracle/hagerzhang.py, oracle/lbfgs.py): the reference's own tests
hold no golden vectors for this path, so what can be checked independently of any reference output is checked
here — the guarantees the published algorithms give, on randomly generated problems (hypothesis), and agreement of the
minimisers with SciPy's unrelated L-BFGS-B implementation."""
import math

import numpy as np
import pytest
from hypothesis import assume, given, settings, strategies as st

import oracle as O

DELTA, SIGMA, EPSILON = 0.1, 0.9, 1e-6      # LineSearches.HagerZhang defaults


@settings(max_examples=150, deadline=None, derandomize=True)
@given(a=st.floats(0.05, 20.0), b=st.floats(-3.0, 3.0), q=st.floats(0.0, 5.0), w=st.floats(0.1, 6.0), amp=st.floats(0.0, 0.8),
       c0=st.sampled_from([1.0, 0.3, 4.0]))
def test_hagerzhang_returns_a_wolfe_point_on_random_line_functions(a, b, q, w, amp, c0):
    """φ(α) = ½a(α−b)² + q(α−b)⁴/4 + amp·a/w²·(1 − cos(w α)): bounded below, possibly non-convex.  Whenever α = 0 is a
    descent point the search must return α > 0 with either the Wolfe or the approximate Wolfe conditions of Hager & Zhang
    (T1/T2 of Algorithm 851), within linesearchmax evaluations."""
    phi = lambda x: 0.5 * a * (x - b) ** 2 + 0.25 * q * (x - b) ** 4 + amp * a / w ** 2 * (1.0 - math.cos(w * x))
    dphi = lambda x: a * (x - b) + q * (x - b) ** 3 + amp * a / w * math.sin(w * x)
    phi0, dphi0 = phi(0.0), dphi(0.0)
    assume(not (-1e-9 < dphi0 < 0.0))       # slopes in the denormal range are a degenerate case of their own (below)
    calls = []

    def phidphi(x):
        calls.append(x)
        return phi(x), dphi(x)

    ls = O.HagerZhang()
    if dphi0 >= 0:
        if dphi0 >= np.finfo(float).eps * abs(phi0):
            with pytest.raises(O.LineSearchException):
                ls(phidphi, c0, phi0, dphi0)
        return
    alpha, val = ls(phidphi, c0, phi0, dphi0)
    assert alpha > 0 and len(calls) <= 50 + 60
    assert val == pytest.approx(phi(alpha), rel=1e-12, abs=1e-12)
    wolfe = (DELTA * dphi0 >= (val - phi0) / alpha) and (dphi(alpha) >= SIGMA * dphi0)
    approx = ((2 * DELTA - 1) * dphi0 >= dphi(alpha) >= SIGMA * dphi0) and (val <= phi0 + EPSILON * abs(phi0))
    assert wolfe or approx, (alpha, val, phi0, dphi0, dphi(alpha))


@settings(max_examples=40, deadline=None, derandomize=True)
@given(n=st.integers(2, 40), logcond=st.floats(0.0, 4.0), seed=st.integers(0, 2 ** 31 - 1))
def test_lbfgs_solves_random_spd_quadratics(n, logcond, seed):
    """f(z) = ½ zᵀAz − bᵀz with a random SPD A of condition number 10^logcond: the iteration must stop with ‖∇f‖∞ ≤ g_tol
    at A⁻¹b, with a monotonically non-increasing objective (every accepted Hager–Zhang step satisfies a decrease condition)."""
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = (Q * np.logspace(0, logcond, n)) @ Q.T
    A = 0.5 * (A + A.T)
    b = rng.standard_normal(n)
    trace = []

    def fg(z):
        g = A @ z - b
        f = 0.5 * z @ (A @ z) - b @ z
        trace.append(f)
        return f, g

    soln = O.lbfgs_minimize(fg, np.zeros(n), g_tol=1e-8)
    zstar = np.linalg.solve(A, b)
    assert soln.g_converged and soln.g_residual <= 1e-8
    assert np.abs(A @ soln.minimizer - b).max() <= 1e-8
    np.testing.assert_allclose(soln.minimizer, zstar, rtol=0, atol=1e-7 * 10 ** logcond * max(1.0, np.abs(zstar).max()) / 10 ** logcond + 1e-6)
    assert soln.minimum <= trace[0] + 1e-12
    assert soln.iterations <= 1000 and soln.f_calls == len(trace)


@settings(max_examples=25, deadline=None, derandomize=True)
@given(n=st.integers(2, 30), seed=st.integers(0, 2 ** 31 - 1))
def test_lbfgs_minimiser_agrees_with_scipy_on_random_smooth_convex_functions(n, seed):
    """f(z) = Σ log cosh(Mz − c)_i + ½λ‖z‖²: smooth, strictly convex, not quadratic.  SciPy's L-BFGS-B (a different code
    base: Nocedal's Fortran) must find the same minimiser."""
    from scipy.optimize import minimize
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((2 * n, n))
    c = rng.standard_normal(2 * n)
    lam = 0.1

    def fg(z):
        r = M @ z - c
        f = np.sum(np.logaddexp(r, -r) - math.log(2.0)) + 0.5 * lam * z @ z
        g = M.T @ np.tanh(r) + lam * z
        return f, g

    soln = O.lbfgs_minimize(fg, np.zeros(n), g_tol=1e-6)
    ref = minimize(lambda z: fg(z), np.zeros(n), jac=True, method="L-BFGS-B", options=dict(gtol=1e-11, ftol=1e-15, maxiter=5000))
    assert soln.g_converged and soln.g_residual <= 1e-6
    np.testing.assert_allclose(soln.minimizer, ref.x, rtol=0, atol=1e-5)
    assert soln.minimum <= ref.fun + 1e-9


def test_hagerzhang_degenerate_slope_follows_ieee_semantics():
    """A descent slope in the denormal range drives the secant step to c = 0; the Julia original (like the CUDA controller)
    then divides by zero under IEEE rules — NaN, comparisons false — instead of raising.  The restatement must do the same."""
    a, b = 1.0, 5e-324
    out = O.HagerZhang()(lambda x: (0.5 * a * (x - b) ** 2, a * (x - b)), 0.3, 0.0, -5e-324)
    assert out[0] >= 0.0 and math.isfinite(out[1])
