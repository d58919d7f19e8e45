"""The generic solver kernel's code — the INIT / TRIAL sweeps, the slow-path vector operations and the whole scalar
Controller (Hager–Zhang, two-loop recursion, L-BFGS loop, unit set-up) of museinference.jl_b200/csrc/muse_iso_ctl.cuh and
muse_iso_solver.cu — executed on the CPU: the blocks marked ``[host-test:…]`` are pasted verbatim between shims
(tests/csrc/generic_solver_host.cpp.in), compiled with g++ (and UBSan) and run as a one-thread "group" over host arrays.  A
complete launch is then held against the oracle, unit by unit: iterations, evaluations, status, ẑ and score."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ZERO, OWN, TRUTH, SHARED, KEEP = 0, 1, 2, 3, 4              # StartKind
FAMILY_ID = {"funnel": 1, "hiergauss": 2}


@pytest.fixture(scope="module")
def gs(tmp_path_factory):
    csrc = os.path.join(ROOT, "museinference.jl_b200", "csrc")
    ctl = open(os.path.join(csrc, "muse_iso_ctl.cuh")).read()
    sol = open(os.path.join(csrc, "muse_iso_solver.cu")).read()

    def block(src, tag):
        m = re.search(r"// \[host-test:begin %s\][^\n]*\n(.*?)// \[host-test:end %s\]" % (tag, tag), src, flags=re.S)
        assert m, tag
        return m.group(1)

    src = open(os.path.join(ROOT, "tests", "csrc", "generic_solver_host.cpp.in")).read()
    stream = open(os.path.join(csrc, "muse_iso_stream.cu")).read()
    src = src.replace("@CTL_A@", block(ctl, "ctl-a")).replace("@CTL_B@", block(ctl, "ctl-b")).replace("@SWEEPS@", block(sol, "sweeps"))
    for tag, key in (("red-slots", "@RED_SLOTS@"), ("item-desc", "@ITEM_DESC@"), ("elem3", "@ELEM3@"), ("fast-replay", "@FAST_REPLAY@"),
                     ("publish", "@PUBLISH@"), ("warp-item", "@WARP_ITEM@")):
        src = src.replace(key, block(stream, tag))
    d = tmp_path_factory.mktemp("gs")
    (d / "gs.cpp").write_text(src)
    out = str(d / "libgs.so")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fsanitize=undefined", "-fno-sanitize-recover=undefined",
                    "-I", csrc, "-I", "/usr/local/cuda/include", "-o", out, str(d / "gs.cpp")], check=True)
    return C.CDLL(out)


class HostSolver:
    """Host arrays of one handle and launches of the generic solver over them."""

    def __init__(self, lib, family, d, draws, xdat, lbfgs_m=10, max_iters=1000, single_pass=False):
        self.lib, self.family, self.d, self.n = lib, family, d, draws.xi.shape[0]
        self.single_pass = single_pass          # True: the streaming kernels' single pass first, generic solver on the hand-backs
        self.redo_total = np.zeros(1, dtype=np.uint64)
        self.last_redo = 0
        self.ld = d + (d & 1) + 2
        self.nt = 2 if family == "hiergauss" else 1
        rows = self.n + 1
        pad = lambda a: np.ascontiguousarray(np.pad(np.atleast_2d(a), ((0, 0), (0, self.ld - d))))
        self.xi = pad(np.vstack([draws.xi, draws.xi_master]))
        self.nu = pad(np.vstack([draws.nu, draws.nu_master]))
        self.xdat = pad(xdat)[0].copy()
        self.zA, self.zB = np.zeros((rows, self.ld)), np.zeros((rows, self.ld))
        self.zstate = np.zeros(rows, dtype=np.int32)
        self.m, self.max_iters = lbfgs_m, max_iters
        self.xslot, self.sbuf = np.zeros(self.ld), np.zeros(self.ld)
        self.dxh, self.dgh = np.zeros((lbfgs_m, self.ld)), np.zeros((lbfgs_m, self.ld))

    def consts(self, th_sim, th_eval):
        d = self.d
        if self.family == "funnel":
            a = np.exp(-th_eval[0]); ev = [a, 0.0, 0.5 * d * th_eval[0], 1.0 / (1.0 + a)]; smp = [np.exp(0.5 * th_sim[0]), 0.0]
        else:
            a = np.exp(-2.0 * th_eval[1]); ev = [a, th_eval[0], d * th_eval[1], 1.0 / (1.0 + a)]; smp = [np.exp(th_sim[1]), th_sim[0]]
        return np.array(ev), np.array(smp)

    def map_score(self, th_sim, th_eval, atol, include_data, start, first=0, count=None, zshared=None):
        count = self.n - first if count is None else count
        items = count + (1 if include_data else 0)
        ev, smp = self.consts(np.atleast_1d(th_sim), np.atleast_1d(th_eval))
        g = np.zeros((items, self.nt)); it = np.zeros(items, dtype=np.int32); fg = np.zeros(items, dtype=np.int32)
        gn = np.zeros(items); f = np.zeros(items); st = np.zeros(items, dtype=np.int32)
        zs = None if zshared is None else np.ascontiguousarray(np.pad(zshared, (0, self.ld - self.d)))
        redo_items = np.zeros(items, dtype=np.int32) if self.single_pass else None
        redo_count = np.zeros(2, dtype=np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        self.lib.muse_host_generic_run(FAMILY_ID[self.family], self.d, self.ld, items, 0, int(include_data), first, start, self.m, self.max_iters,
                                       C.c_double(atol), p(ev), p(smp), 1, p(self.xi), p(self.nu), p(self.xdat), p(zs), self.n,
                                       p(self.xslot), p(self.zA), p(self.zB), p(self.zstate), p(self.sbuf), p(self.dxh), p(self.dgh),
                                       p(g), p(it), p(fg), p(gn), p(f), p(st), p(redo_items), p(redo_count), p(self.redo_total) if self.single_pass else None, 0)
        self.last_redo = int(redo_count[0])
        return dict(g=g, iters=it, fg=fg, gnorm=gn, f=f, status=st)

    def z(self, unit):
        s = self.zstate[unit]
        return np.zeros(self.d) if s == 0 else (self.zA if s == 1 else self.zB)[unit, :self.d].copy()


def _status(soln):
    return 0 if soln.g_converged else (1 if soln.converged else 2)


@pytest.mark.parametrize("single_pass", [False, True])
@pytest.mark.parametrize("family,d,atol", [("funnel", 37, 1e-2), ("funnel", 64, 1e-8), ("hiergauss", 51, 1e-2), ("hiergauss", 20, 1e-6)])
def test_generic_solver_code_on_the_host_matches_the_oracle(gs, family, d, atol, single_pass):
    rng = np.random.default_rng(d)
    nsims = 14
    fam = O.make_family(family, d)
    draws = O.Draws.from_numpy(int(rng.integers(1 << 30)), nsims, d)
    th_true = np.zeros(fam.ntheta)
    xd, _ = fam.sample(th_true, rng.standard_normal(d), rng.standard_normal(d))
    prob = O.OracleProblem(fam, xd, draws)
    hs = HostSolver(gs, family, d, draws, xd, single_pass=single_pass)
    th0 = np.array([0.7]) if family == "funnel" else np.array([0.4, 0.25])

    def check(out, units, theta_sim, theta_eval, starts, z_prev):
        for i, u in enumerate(units):
            x = xd if u == 0 else prob.sample_x_z(u - 1, theta_sim)[0]
            _, g, soln = O.map_score_unit(prob, x, starts[i], theta_eval, atol)
            assert (out["iters"][i], out["fg"][i], out["status"][i]) == (soln.iterations, soln.f_calls, _status(soln)), (u, out["iters"][i], soln.iterations)
            np.testing.assert_allclose(hs.z(u), soln.minimizer, rtol=1e-11, atol=1e-12)
            np.testing.assert_allclose(out["g"][i], g, rtol=1e-10, atol=1e-9)
            np.testing.assert_allclose(out["f"][i], soln.minimum, rtol=1e-12)
            np.testing.assert_allclose(out["gnorm"][i], soln.g_residual, rtol=1e-6, atol=1e-12)

    units = list(range(nsims + 1))
    # cold pass: data + sims from zero(z)                                        (src/muse.jl:169-176, iteration 1)
    out = hs.map_score(th0, th0, atol, True, ZERO)
    check(out, units, th0, th0, [np.zeros(d)] * (nsims + 1), None)
    assert (out["iters"] == 1).all() and (out["fg"] == 3).all()
    assert hs.last_redo == 0                 # single pass: every unit finished on the fast path, nothing handed back
    zprev = [hs.z(u) for u in units]
    # warm pass at a moved θ from the previous ẑ
    th1 = th0 - 0.3
    out = hs.map_score(th1, th1, atol, True, OWN)
    check(out, units, th1, th1, zprev, None)
    # the same θ again: the start already satisfies the tolerance → 0 iterations, 1 evaluation, ẑ kept
    zprev = [hs.z(u) for u in units]
    out = hs.map_score(th1, th1, max(atol, 1e-7), True, OWN)
    assert (out["iters"] == 0).all() and (out["fg"] == 1).all()
    for u in units:
        np.testing.assert_array_equal(hs.z(u), zprev[u])
    # get_J!: a sub-range of sims from the simulated latent                     (src/muse.jl:508-514)
    out = hs.map_score(th0, th0, atol, False, TRUTH, first=3, count=6)
    check(out, list(range(4, 10)), th0, th0, [prob.sample_x_z(k, th0)[1] for k in range(3, 9)], None)
    # user z₀ kept as the start of every unit                                   (`z₀` keyword)
    z0 = rng.normal(0, 0.5, d)
    out = hs.map_score(th0, th0, atol, True, KEEP, zshared=z0)
    check(out, units, th0, th0, [z0] * (nsims + 1), None)
    # sims at one θ, MAP and score at another                                   (get_H!: src/muse.jl:430-432)
    out = hs.map_score(th0 + 0.05, th0, atol, False, ZERO)
    check(out, list(range(1, nsims + 1)), th0 + 0.05, th0, [np.zeros(d)] * nsims, None)


@pytest.mark.parametrize("family,d", [("funnel", 33), ("hiergauss", 48)])
def test_generic_solver_history_path_iteration_cap_and_non_finite_data_on_the_host(gs, family, d):
    """atol below round-off drives the history path (two-loop recursion over stored (dx, dg), direction resets, x/f
    stagnation exits) — here under UBSan; the iterate must stay at the closed-form MAP.  max_iters caps; NaN data reports."""
    rng = np.random.default_rng(7 + d)
    nsims = 6
    fam = O.make_family(family, d)
    draws = O.Draws.from_numpy(99, nsims, d)
    xd, _ = fam.sample(np.zeros(fam.ntheta), rng.standard_normal(d), rng.standard_normal(d))
    prob = O.OracleProblem(fam, xd, draws)
    th = np.array([0.8]) if family == "funnel" else np.array([0.3, -0.2])
    hs = HostSolver(gs, family, d, draws, xd, single_pass=True)
    out = hs.map_score(th, th, 1e-300, True, ZERO)
    assert hs.last_redo == nsims + 1         # the single pass cannot finish such a unit: all handed back, re-solved from the untouched start
    assert (out["iters"] >= 2).all() and np.isin(out["status"], [0, 1, 3]).all()
    for u in range(nsims + 1):
        x = xd if u == 0 else prob.sample_x_z(u - 1, th)[0]
        np.testing.assert_allclose(hs.z(u), fam.exact_map(x, th), rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(out["g"][u], fam.score(x, fam.exact_map(x, th), th), rtol=1e-9, atol=1e-9)
        _, _, soln = O.map_score_unit(prob, x, np.zeros(d), th, 1e-300)
        assert soln.iterations >= 2
    hs2 = HostSolver(gs, family, d, draws, xd, max_iters=1)
    out = hs2.map_score(th, th, 1e-300, True, ZERO)
    assert (out["iters"] == 1).all() and np.isin(out["status"], [1, 2]).all()
    xbad = xd.copy(); xbad[d // 2] = np.nan
    hs3 = HostSolver(gs, family, d, draws, xbad)
    out = hs3.map_score(th, th, 1e-2, True, ZERO)
    assert out["status"][0] == 4 and (out["status"][1:] == 0).all()


@pytest.mark.parametrize("family,d", [("funnel", 40), ("hiergauss", 27)])
def test_generic_solver_fd_launches_on_the_host_match_the_oracle(gs, family, d):
    """get_H!'s two launches (src/muse.jl:417-433): mode 2 — the fiducial MAP of the master stream's draw from zero(z) — then
    mode 1 — 2·nθ virtual sims per H sim at θ₀ ∓ h·eₙ, MAP and score at θ₀ from the shared fiducial start."""
    rng = np.random.default_rng(3 * d)
    nsims, nH, atol = 9, 4, 1e-2
    fam = O.make_family(family, d)
    nt = fam.ntheta
    draws = O.Draws.from_numpy(5, nsims, d)
    xd, _ = fam.sample(np.zeros(nt), rng.standard_normal(d), rng.standard_normal(d))
    prob = O.OracleProblem(fam, xd, draws)
    th0 = np.array([0.6]) if family == "funnel" else np.array([0.2, 0.15])
    step = 0.03 * (1 + np.arange(nt))
    hs = HostSolver(gs, family, d, draws, xd)
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    ev, smp0 = hs.consts(th0, th0)
    # fiducial: one item, mode 2, start zeros, own buffers (row 0 of a 1-row pair)
    zfA, zfB, zfs = np.zeros((1, hs.ld)), np.zeros((1, hs.ld)), np.zeros(1, dtype=np.int32)
    g1, it1, fg1, gn1, f1, st1 = np.zeros((1, nt)), np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1), np.zeros(1), np.zeros(1, np.int32)
    gs.muse_host_generic_run(FAMILY_ID[family], d, hs.ld, 1, 2, 0, 0, ZERO, 10, 1000, C.c_double(atol), p(ev), p(smp0), 1, p(hs.xi), p(hs.nu),
                             p(hs.xdat), None, nsims, p(hs.xslot), p(zfA), p(zfB), p(zfs), p(hs.sbuf), p(hs.dxh), p(hs.dgh),
                             p(g1), p(it1), p(fg1), p(gn1), p(f1), p(st1), None, None, None, 0)
    xm, _ = prob.sample_x_z("master", th0)
    zfid_ref, soln = prob.z_at_theta(xm, np.zeros(d), th0, atol)
    zfid = (zfA if zfs[0] == 1 else zfB)[0, :d].copy()
    np.testing.assert_allclose(zfid, zfid_ref, rtol=1e-11, atol=1e-12)
    assert (it1[0], fg1[0]) == (soln.iterations, soln.f_calls)
    # virtual sims: items (k·nθ + n)·2 + s, constants of the sample points in smp[2n + s], shared start = the fiducial ẑ
    items = nH * nt * 2
    smp = np.zeros((2 * nt, 2))
    pts = []
    for n in range(nt):
        for s in (0, 1):
            th = th0.copy(); th[n] += step[n] * (1.0 if s else -1.0)
            smp[2 * n + s] = hs.consts(th, th0)[1]
            pts.append(th)
    zHA, zHB = np.zeros((items, hs.ld)), np.zeros((items, hs.ld))
    g, it, fg, gn, f, st = np.zeros((items, nt)), np.zeros(items, np.int32), np.zeros(items, np.int32), np.zeros(items), np.zeros(items), np.zeros(items, np.int32)
    zsh = np.ascontiguousarray(np.pad(zfid, (0, hs.ld - d)))
    gs.muse_host_generic_run(FAMILY_ID[family], d, hs.ld, items, 1, 0, 0, SHARED, 10, 1000, C.c_double(atol), p(ev), p(np.ascontiguousarray(smp)), 2 * nt,
                             p(hs.xi), p(hs.nu), p(hs.xdat), p(zsh), nsims, p(hs.xslot), p(zHA), p(zHB), None, p(hs.sbuf), p(hs.dxh), p(hs.dgh),
                             p(g), p(it), p(fg), p(gn), p(f), p(st), None, None, None, 0)
    for k in range(nH):
        for n in range(nt):
            for s in (0, 1):
                item = (k * nt + n) * 2 + s
                x, _ = prob.sample_x_z(k, pts[2 * n + s])
                zh, sol = prob.z_at_theta(x, zfid_ref, th0, atol)
                assert (it[item], fg[item], st[item]) == (sol.iterations, sol.f_calls, _status(sol))
                np.testing.assert_allclose(g[item], prob.grad_theta(x, zh, th0), rtol=1e-10, atol=1e-9)
    # the Jacobian the host forms from them (central_fdm(3,1), src/util.jl:13-19) against the oracle's get_H!
    res = O.MuseResult(theta=th0.copy())
    O.get_H_bang(res, prob, th0, nsims=nH, step=step, gradz_logLike_atol=atol)
    for k in range(nH):
        Hk = np.stack([(g[(k * nt + n) * 2 + 1] * 0.5 + g[(k * nt + n) * 2] * -0.5) / step[n] for n in range(nt)], axis=1)
        np.testing.assert_allclose(Hk, res.Hs[k], rtol=1e-7, atol=1e-7 * np.abs(res.Hs[k]).max())


@pytest.mark.parametrize("single_pass", [False, True])
def test_zero_iteration_solve_from_zeros_leaves_a_zero_map_not_a_stale_one(gs, single_pass):
    """A unit whose zero start already satisfies the tolerance ends with ẑ = zero(z).  The single-pass kernels used to leave
    the unit's state cell untouched in that case, so the buffer of an EARLIER solve passed for its ẑ (found by the random
    sweep of this harness; publish_unit now records kZZero, as the generic kernel always did)."""
    d, nsims = 6, 5
    fam = O.make_family("funnel", d)
    draws = O.Draws.from_numpy(3, nsims, d)
    xd = np.full(d, 0.3)
    hs = HostSolver(gs, "funnel", d, draws, xd, single_pass=single_pass)
    th = np.array([0.5])
    hs.map_score(th, th, 1e-2, True, ZERO)                       # a real solve: every unit now holds a non-zero ẑ
    assert all(np.abs(hs.z(u)).max() > 0 for u in range(nsims + 1))
    out = hs.map_score(th, th, 1e3, True, ZERO)                  # tolerance so loose that zero(z) is already "converged"
    assert (out["iters"] == 0).all() and (out["fg"] == 1).all() and (out["status"] == 0).all()
    for u in range(nsims + 1):
        np.testing.assert_array_equal(hs.z(u), np.zeros(d))
