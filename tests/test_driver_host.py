"""The in-library host loop and covariance stage (museinference.jl_b200/csrc/muse_driver.cu: muse_b200_muse_iterate,
muse_b200_muse_covariance) compiled for the host against stubs of the solver passes (tests/csrc/driver_host.cpp) and held
against a NumPy restatement of /root/reference/src/muse.jl:163-166, 183-224, 411-413, 446, 529, 535-541 on canned scores —
for every nθ from 1 to MUSE_MAX_NTHETA = 8, although the registered families only reach nθ ≤ 2 on the GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def drv(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("drv") / "libdrv.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-I", "/usr/local/cuda/include", "-o", out,
                    os.path.join(ROOT, "tests", "csrc", "driver_host.cpp")], check=True)
    return C.CDLL(out)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


@pytest.mark.parametrize("nt,prior", [(1, True), (2, False), (3, True), (5, False), (8, True)])
def test_host_loop_and_covariance_stage_match_numpy(drv, nt, prior):
    rng = np.random.default_rng(100 + nt)
    nsims, maxsteps, nh, alpha, rtol = 57, 9, 6, 0.7, 0.02
    theta0 = rng.normal(0, 1, nt)
    pm = rng.normal(0, 1, nt) if prior else None
    ps = rng.uniform(0.5, 3.0, nt) if prior else None
    # canned scores: sims scatter around a mean that shrinks from iteration to iteration (so that the loop converges)
    canned = np.empty((maxsteps, nsims + 1, nt))
    for i in range(maxsteps):
        canned[i, 1:] = rng.normal(0, 1, (nsims, nt)) * rng.uniform(2, 5, nt) + rng.normal(0, 3, nt)
        canned[i, 0] = canned[i, 1:].mean(axis=0) + rng.normal(0, 1, nt) * 0.4 ** i
    Hs = rng.normal(0, 1, (nh, nt, nt)) + 4 * np.eye(nt)
    n_iter = C.c_int(0)
    th_final, th_hist, hpost_hist, glike_hist = np.zeros(nt), np.zeros((maxsteps, nt)), np.zeros((maxsteps, nt)), np.zeros((maxsteps, nt))
    J, step, H, Sinv, S = np.zeros((nt, nt)), np.zeros(nt), np.zeros((nt, nt)), np.zeros((nt, nt)), np.zeros((nt, nt))
    rc = drv.host_driver_run(C.c_int(nt), C.c_int(nsims), C.c_int(maxsteps), C.c_double(rtol), C.c_double(alpha), _p(theta0), _p(pm), _p(ps),
                             _p(canned), _p(Hs), C.c_int(nh), C.byref(n_iter), _p(th_final), _p(th_hist), _p(hpost_hist), _p(glike_hist),
                             _p(J), _p(step), _p(H), _p(Sinv), _p(S))
    assert rc == 0
    # NumPy restatement
    theta, hist_t, hist_h, n = theta0.copy(), [], [], 0
    for i in range(1, maxsteps + 1):
        if i > 2:
            dth = hist_t[-1] - hist_t[-2]
            if np.sqrt(-(dth * hist_h[-1] * dth).sum()) < rtol:
                break
        g = canned[i - 1]
        g_like = g[0] - g[1:].mean(axis=0)
        g_prior = -(theta - pm) / ps ** 2 if prior else np.zeros(nt)
        h_like = -1.0 / g[1:].var(axis=0, ddof=1)
        h_prior = -1.0 / ps ** 2 if prior else np.zeros(nt)
        h_post = 1.0 / (1.0 / h_like + h_prior)
        hist_t.append(theta.copy()); hist_h.append(h_post)
        np.testing.assert_allclose(glike_hist[i - 1], g_like, rtol=1e-12, atol=1e-13)
        theta = theta - alpha * h_post * (g_like + g_prior)
        n = i
    assert n_iter.value == n and 2 <= n <= maxsteps
    np.testing.assert_allclose(th_final, theta, rtol=1e-12)
    np.testing.assert_allclose(th_hist[:n], np.array(hist_t), rtol=1e-12)
    np.testing.assert_allclose(hpost_hist[:n], np.array(hist_h), rtol=1e-12)
    gs = canned[n - 1, 1:]
    Jref = np.atleast_2d(np.cov(gs, rowvar=False, ddof=1))
    Href = Hs.mean(axis=0)
    Sinv_ref = Href.T @ np.linalg.inv(Jref) @ Href + (np.diag(1.0 / ps ** 2) if prior else 0.0)
    np.testing.assert_allclose(J, Jref, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(step, 0.1 / gs.std(axis=0, ddof=1), rtol=1e-12)
    np.testing.assert_allclose(H, Href, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(Sinv, Sinv_ref, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(S, np.linalg.inv(Sinv_ref), rtol=1e-8, atol=1e-10)
