"""Model of the device-resident outer loop (TEST INFRASTRUCTURE ONLY).

csrc/muse_outer.cu runs the θ iteration of ``muse!`` (/root/reference/src/muse.jl:159-236) without a host round trip between
passes: the host enqueues *chunks* of passes (two in the first chunk, then three at a time); after each pass a single-CTA
kernel (``theta_step_kernel``) reduces the gathered scores with a fixed parallel tree, applies src/muse.jl:183-224, evaluates
the convergence test that the reference makes at the top of the NEXT iteration (:163-166) and raises ``done``; passes enqueued
after ``done`` see ``skip`` and return at once; the covariance stage rides the chunk in which ``done`` was raised.  This module
restates that control flow and the kernel's summation order in NumPy, so that CPU tests can hold it against the
line-by-line restatement (oracle/muse.py): same iteration count, same θ to round-off, and the number of host
synchronisations the scheme needs."""
from __future__ import annotations

import math

import numpy as np

THREADS = 1024        # kStepThreads
FIRST_CHUNK, CHUNK = 2, 3


def tree_sum(v: np.ndarray) -> float:
    """Σ v in the order of ``block_sums``: thread t adds elements t, t + T, t + 2T, … sequentially; a butterfly over the 32
    lanes of each warp; a butterfly over the 32 warp results."""
    n = v.size
    acc = np.zeros(THREADS)
    for start in range(0, n, THREADS):                    # every thread adds its next element (same per-thread order)
        blk = v[start:start + THREADS]
        acc[:blk.size] += blk
    w = acc.reshape(32, 32).copy()                        # [warp, lane]
    for o in (16, 8, 4, 2, 1):                            # __shfl_xor butterfly: lane l adds lane l ^ o
        w = w + w[:, np.arange(32) ^ o]
    s = w[:, 0].copy()                                    # lane 0 of every warp
    for o in (16, 8, 4, 2, 1):
        s = s + s[np.arange(32) ^ o]
    return float(s[0])


def theta_step(theta, g_dat, g_sims, alpha, prior_mean=None, prior_sigma=None):
    """One ``theta_step_kernel``: returns (θ_new, h_inv_post) — src/muse.jl:183-224 with the kernel's reduction order."""
    n, nt = g_sims.shape
    th_new, h_post = np.empty(nt), np.empty(nt)
    for c in range(nt):
        mean = tree_sum(g_sims[:, c]) / n
        var = tree_sum((g_sims[:, c] - mean) ** 2) / (n - 1)
        g_like = g_dat[c] - mean
        g_prior = -(theta[c] - prior_mean[c]) / prior_sigma[c] ** 2 if prior_sigma is not None else 0.0
        h_like = -1.0 / var
        h_prior = -1.0 / prior_sigma[c] ** 2 if prior_sigma is not None else 0.0
        h_post[c] = 1.0 / (1.0 / h_like + h_prior)
        th_new[c] = theta[c] - alpha * (h_post[c] * (g_like + g_prior))
    return th_new, h_post


def run(pass_fn, theta0, maxsteps=50, theta_rtol=1e-1, alpha=0.7, prior_mean=None, prior_sigma=None):
    """``pass_fn(i, θ) -> (g_dat, g_sims)`` is the solver pass of iteration i.  Returns dict(theta, n_iter, syncs, launched,
    skipped, theta_hist): ``launched`` passes were enqueued, ``skipped`` of them returned at once, ``syncs`` host
    synchronisations were needed (one per chunk)."""
    theta = np.array(theta0, dtype=np.float64)
    hist_theta, hist_hpost = [], []
    n_iter, done, syncs, launched, skipped = 0, False, 0, 0, 0
    while not done and n_iter < maxsteps:
        first = n_iter + 1
        last = min(maxsteps, n_iter + (FIRST_CHUNK if n_iter == 0 else CHUNK))
        for i in range(first, last + 1):                  # what the host enqueues; the device decides what runs
            launched += 1
            if done:
                skipped += 1
                continue
            g_dat, g_sims = pass_fn(i, theta.copy())
            hist_theta.append(theta.copy())
            theta, h_post = theta_step(theta, g_dat, g_sims, alpha, prior_mean, prior_sigma)
            hist_hpost.append(h_post)
            n_iter = i
            if i >= 2:                                    # the test at the top of iteration i + 1 > 2
                dth = hist_theta[-1] - hist_theta[-2]
                q = -float(np.sum(dth * hist_hpost[-1] * dth))
                if q < 0:
                    raise ValueError("DomainError in the θ convergence test")
                if math.sqrt(q) < theta_rtol:
                    done = True
            if i >= maxsteps:
                done = True
        syncs += 1
    return dict(theta=theta, n_iter=n_iter, syncs=syncs, launched=launched, skipped=skipped, theta_hist=np.array(hist_theta))
