"""Full MUSE solve on the CPU with the batch body in C (TEST INFRASTRUCTURE ONLY).

Same outer arithmetic as oracle/muse.py (restating /root/reference/src/muse.jl:112-250, 296-333,
407-450, 484-549), but the mapped per-simulation blocks run through oracle/csrc/muse_oracle.c on
all host threads.  This is what ``bench.py`` times as ``cpu_baseline`` and ``--impl reference``:
the reference itself (Julia) cannot run here, its default pool is a serial ``map`` and its
gradients come from AD, so this analytic-gradient, threaded port is strictly *faster* than the
real reference would be on the same cores.  PARITY UNPINNED (oracle/__init__.py).
"""
from __future__ import annotations

import math

import numpy as np

from . import cport
from .muse import MuseResult, finalize_result_bang


def muse_cpu(prob, theta0, *, nsims, gradz_logLike_atol=1e-2, maxsteps=50, theta_rtol=1e-1, alpha=0.7,
             get_covariance=True, nthreads=0):
    """Returns (MuseResult, units) where units = MAP+score bodies executed."""
    fam, dr = prob.family, prob.draws
    fid = fam.family_id
    kw = dict(P=fam.P, L=fam.L) if fid == 3 else {}
    res = MuseResult()
    theta = np.atleast_1d(np.asarray(theta0, dtype=np.float64)).copy()
    hist = res.history
    xi, nu = dr.xi[:nsims], dr.nu[:nsims]
    z = None
    units = 0
    for i in range(1, maxsteps + 1):
        if i > 2:
            dth = hist[-1]["theta"] - hist[-2]["theta"]
            if math.sqrt(-(dth @ hist[-1]["H_inv_post"] @ dth)) < theta_rtol:
                break
        out = cport.map_score(fid, xi, nu, prob.x, theta, theta, gradz_logLike_atol, True, 0 if z is None else 1,
                              z_start=z, want_z=True, nthreads=nthreads, **kw)
        z = out["z"]
        units += nsims + 1
        g_dat, g_sims = out["g"][0], out["g"][1:]
        g_like = g_dat - g_sims.mean(axis=0)
        g_post = g_like + prob.prior.grad(theta)
        H_inv_like = np.diag(-1.0 / np.var(g_sims, axis=0, ddof=1))
        H_inv_post = np.linalg.inv(np.linalg.inv(H_inv_like) + prob.prior.hess(theta))
        hist.append(dict(theta=theta.copy(), g_like=g_like, g_post=g_post, H_inv_post=H_inv_post))
        theta_unreg = theta - alpha * (H_inv_post @ g_post)
        theta = theta_unreg.copy()
        res.theta = theta_unreg.copy()
        res.gs = [g.copy() for g in g_sims]
    if get_covariance:
        gs = np.array(res.gs)
        res.J = np.array([[np.var(gs[:, 0], ddof=1)]]) if theta.size == 1 else np.cov(gs, rowvar=False, ddof=1)
        th0 = res.theta
        nH = max(1, nsims // 10)
        step = 0.1 / np.std(gs, axis=0, ddof=1)
        fo = cport.map_score(fid, dr.xi_master[None], dr.nu_master[None], None, th0, th0, gradz_logLike_atol, False, 0,
                             want_z=True, nthreads=nthreads, **kw)
        units += 1
        zfid = np.repeat(fo["z"], nH, axis=0)
        nt = th0.size
        Hs = np.zeros((nH, nt, nt))
        for n in range(nt):
            gpm = []
            for sgn in (-1.0, 1.0):
                th = th0.copy()
                th[n] = th0[n] + (0.0 + step[n] * sgn)
                o = cport.map_score(fid, dr.xi[:nH], dr.nu[:nH], None, th, th0, gradz_logLike_atol, False, 1,
                                    z_start=zfid, nthreads=nthreads, **kw)
                units += nH
                gpm.append(o["g"])
            Hs[:, :, n] = ((gpm[0] * -0.5 + 0.0) + gpm[1] * 0.5) / step[n]
        res.Hs = list(Hs)
        res.H = Hs.mean(axis=0)
        finalize_result_bang(res, prob)
    return res, units
