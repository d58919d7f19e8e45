"""Test double for ``B200Backend`` built on the CPU oracle (tests only).

Lets the host-side driver (museinference.jl_b200/muse.py, parallel.py) run on a machine without a
GPU: same method signatures and semantics as the real backend, per-sim work done by oracle/*.py.
"""
import numpy as np

import oracle as O


class FakeBackend:
    def __init__(self, family, d, nsims, *, sim_offset=0, nsims_h=0, h_sim_offset=0, device=0, group=0, cluster=0,
                 kernel=0, stream=None, lbfgs_m=0, max_iters=0, P=None, L=None):
        self.family, self.d, self.nsims, self.nsims_h = family, d, nsims, nsims_h
        self.sim_offset, self.h_sim_offset = sim_offset, h_sim_offset
        self.fam = O.make_family(family, d)
        self.ntheta = self.fam.ntheta
        self.x = None
        self.draws = None
        self.draws_h = None
        self.z = [np.zeros(d) for _ in range(nsims + 1)]
        self.z0 = None
        self.calls = []

    def close(self):
        pass

    def set_data(self, x):
        self.x = np.array(x, dtype=np.float64)

    def set_draws(self, xi, nu, xm, nm):
        self.draws = O.Draws(np.array(xi), np.array(nu), np.array(xm), np.array(nm))

    def set_draws_h(self, xi, nu):
        self.draws_h = (np.array(xi), np.array(nu))

    def seed_draws(self, seed):
        self.draws = O.Draws.from_philox(seed, self.nsims, self.d, offset=self.sim_offset)
        if self.nsims_h:
            h = O.Draws.from_philox(seed, self.nsims_h, self.d, offset=self.h_sim_offset)
            self.draws_h = (h.xi, h.nu)

    def set_z0(self, z0):
        self.z0 = np.array(z0, dtype=np.float64)

    def fd_start(self, start):
        self.fd_user_start = int(start) == 3        # MUSE_START_USER

    def _prob(self):
        return O.OracleProblem(self.fam, self.x if self.x is not None else np.zeros(self.d), self.draws)

    def map_score(self, theta_sim, theta_eval, atol, *, include_data, warm_start, first_sim=0, count=None):
        count = self.nsims - first_sim if count is None else count
        # the real backend rejects ranges outside the handle's shard (muse_b200_map_score_async) and empty launches
        assert first_sim >= 0 and count >= 0 and first_sim + count <= self.nsims, "sim range outside the handle's shard"
        assert count + (1 if include_data else 0) > 0, "empty launch"
        prob = self._prob()
        units = ([0] if include_data else []) + [1 + first_sim + i for i in range(count)]
        self.calls.append(("map_score", tuple(np.atleast_1d(theta_eval)), include_data, warm_start, first_sim, count))
        g, it, fg, gn, st = [], [], [], [], []
        for u in units:
            if u == 0:
                x, ztrue = prob.x, None
            else:
                x, ztrue = prob.sample_x_z(u - 1, theta_sim)
            if warm_start == 0:
                z0 = np.zeros(self.d)
            elif warm_start == 1:
                z0 = self.z[u]
            elif warm_start == 2:
                z0 = ztrue if ztrue is not None else np.zeros(self.d)
            else:
                z0 = self.z0
            zh, gi, soln = O.map_score_unit(prob, x, z0, theta_eval, atol)
            self.z[u] = zh
            g.append(gi); it.append(soln.iterations); fg.append(soln.f_calls); gn.append(soln.g_residual)
            st.append(0 if soln.g_converged else (1 if soln.converged else 2))
        return dict(g=np.array(g).reshape(len(units), self.ntheta), iters=np.array(it, dtype=np.int32),
                    fg_evals=np.array(fg, dtype=np.int32), gnorm=np.array(gn), status=np.array(st, dtype=np.int32))

    def fd_jacobian(self, theta0, step, nsims_H, atol):
        if self.nsims_h:
            dr = O.Draws(self.draws_h[0], self.draws_h[1], self.draws.xi_master, self.draws.nu_master)
        else:
            dr = self.draws
        prob = O.OracleProblem(self.fam, self.x, dr)
        res = O.MuseResult(theta=np.array(theta0, dtype=np.float64))
        O.get_H_bang(res, prob, theta0, nsims=nsims_H, step=np.atleast_1d(step), gradz_logLike_atol=atol,
                     z0=self.z0 if getattr(self, "fd_user_start", False) else None)
        Hs = np.array(res.Hs).reshape(nsims_H, self.ntheta, self.ntheta)
        return Hs, np.zeros((nsims_H, self.ntheta, 2), dtype=np.int32)

    def implicit_h(self, theta0, nsims_H, start=0, cg_maxiter=100):
        prob = O.OracleProblem(self.fam, self.x, self.draws)
        Hs, its = [], []
        for k in range(nsims_H):
            H, it = O.implicit_diff_H(prob, k, theta0, self.z0 if start == 3 else None, cg_maxiter)
            Hs.append(H); its.append(it)
        return np.array(Hs).reshape(nsims_H, self.ntheta, self.ntheta), np.array(its, dtype=np.int32), np.zeros(nsims_H, dtype=np.int32)

    def fd_scores(self, theta_eval, theta_sims, nsims_H, atol):
        if self.nsims_h:
            dr = O.Draws(self.draws_h[0], self.draws_h[1], self.draws.xi_master, self.draws.nu_master)
        else:
            dr = self.draws
        prob = O.OracleProblem(self.fam, self.x, dr)
        theta_eval = np.atleast_1d(np.asarray(theta_eval, dtype=np.float64))
        theta_sims = np.asarray(theta_sims, dtype=np.float64).reshape(2 * self.ntheta, self.ntheta)
        self.calls.append(("fd_scores", tuple(theta_eval), nsims_H))
        xm, zm = prob.sample_x_z("master", theta_eval)
        zfid, _ = prob.z_at_theta(xm, self.z0 if getattr(self, "fd_user_start", False) else np.zeros(self.d), theta_eval, atol)
        g = np.empty((nsims_H, 2 * self.ntheta, self.ntheta))
        for k in range(nsims_H):
            for p in range(2 * self.ntheta):
                x, _ = prob.sample_x_z(k, theta_sims[p])
                zh, _ = prob.z_at_theta(x, zfid, theta_eval, atol)
                g[k, p] = prob.grad_theta(x, zh, theta_eval)
        return g, np.zeros((nsims_H, 2 * self.ntheta), dtype=np.int32)

    def get_maps(self, first_unit, count):
        return np.array(self.z[first_unit:first_unit + count])


class FusedFakeBackend(FakeBackend):
    """FakeBackend plus the in-library outer loops (``muse_iterate`` / ``muse_covariance`` / ``muse_solve`` of
    include/muse_b200.h) restated on the host from the same per-sim pieces, single rank: lets the glue that turns the
    library's history arrays back into ``MuseResult.history`` (museinference.jl_b200/muse.py, fused path) run on the CPU."""

    def muse_iterate(self, theta0, nsims_total, counts, maxsteps, theta_rtol, atol, alpha, first_start, prior_mean=None, prior_sigma=None):
        assert counts is None and nsims_total == self.nsims
        nt, N, K, units = self.ntheta, self.nsims, int(maxsteps), self.nsims + 1
        f = lambda *s: np.zeros(s)
        cache = self.__dict__.setdefault("_bufs", {})      # like the real wrapper: buffers are reused, the caller copies
        if (K, N) not in cache:
            cache[(K, N)] = dict(theta_final=f(nt), theta_hist=f(K, nt), g_dat_hist=f(K, nt), g_sims_hist=f(K, N, nt), g_like_hist=f(K, nt),
                                 g_prior_hist=f(K, nt), h_inv_like_hist=f(K, nt), h_prior_hist=f(K, nt), h_inv_post_hist=f(K, nt),
                                 seconds_hist=f(K), iters_hist=np.zeros((K, units), dtype=np.int32),
                                 fg_hist=np.zeros((K, units), dtype=np.int32), gnorm_hist=f(K, units),
                                 status_hist=np.zeros((K, units), dtype=np.int32))
        r = dict(cache[(K, N)], n_iter=0)
        theta = np.array(theta0, dtype=np.float64)
        pm = np.asarray(prior_mean, dtype=np.float64) if prior_mean is not None else None
        ps = np.asarray(prior_sigma, dtype=np.float64) if prior_sigma is not None else None
        for i in range(1, K + 1):
            if i > 2:
                dth = r["theta_hist"][i - 2] - r["theta_hist"][i - 3]
                if np.sqrt(-(dth * r["h_inv_post_hist"][i - 2] * dth).sum()) < theta_rtol:
                    break
            out = self.map_score(theta, theta, atol, include_data=True, warm_start=first_start if i == 1 else 1)
            k = i - 1
            gs = out["g"][1:]
            g_like = out["g"][0] - gs.mean(axis=0)
            g_prior = -(theta - pm) / ps ** 2 if ps is not None else np.zeros(nt)
            h_like = -1.0 / gs.var(axis=0, ddof=1)
            h_prior = -1.0 / ps ** 2 if ps is not None else np.zeros(nt)
            h_post = 1.0 / (1.0 / h_like + h_prior)
            r["theta_hist"][k], r["g_dat_hist"][k], r["g_sims_hist"][k], r["g_like_hist"][k] = theta, out["g"][0], gs, g_like
            r["g_prior_hist"][k], r["h_inv_like_hist"][k], r["h_prior_hist"][k], r["h_inv_post_hist"][k] = g_prior, h_like, h_prior, h_post
            r["iters_hist"][k], r["fg_hist"][k], r["gnorm_hist"][k], r["status_hist"][k] = out["iters"], out["fg_evals"], out["gnorm"], out["status"]
            r["seconds_hist"][k] = 1e-3
            theta = theta - alpha * (h_post * (g_like + g_prior))
            r["n_iter"] = i
            r["theta_final"][:] = theta
        return r

    def muse_covariance(self, theta, gs, nsims_h_total, counts_h, atol, prior_sigma=None):
        nt = self.ntheta
        gs = np.asarray(gs, dtype=np.float64).reshape(-1, nt)
        J = np.atleast_2d(np.cov(gs, rowvar=False, ddof=1)) if nt > 1 else np.array([[gs[:, 0].var(ddof=1)]])
        step = 0.1 / gs.std(axis=0, ddof=1)
        Hs, _ = self.fd_jacobian(theta, step, nsims_h_total, atol)
        H = Hs.mean(axis=0)
        Hp = np.diag(1.0 / np.asarray(prior_sigma, dtype=np.float64) ** 2) if prior_sigma is not None else np.zeros((nt, nt))
        Sinv = H.T @ np.linalg.inv(J) @ H + Hp
        return dict(J=J, step=step, Hs=Hs, H=H, Sigma_inv=Sinv, Sigma=np.linalg.inv(Sinv))

    def muse_solve(self, theta0, nsims_total, counts, maxsteps, theta_rtol, atol, alpha, first_start, prior_mean=None, prior_sigma=None,
                   get_covariance=False, nsims_h_total=0, counts_h=None):
        r = self.muse_iterate(theta0, nsims_total, counts, maxsteps, theta_rtol, atol, alpha, first_start, prior_mean, prior_sigma)
        c = None
        if get_covariance:
            c = self.muse_covariance(r["theta_final"], r["g_sims_hist"][r["n_iter"] - 1], nsims_h_total, counts_h, atol, prior_sigma)
        return r, c
