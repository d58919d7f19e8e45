"""The streaming kernels' scalar decision code — ``fast_replay`` of museinference.jl_b200/csrc/muse_iso_stream.cu, which
decides from a unit's 15 sums whether the single-pass speculative solve IS what the reference's L-BFGS + Hager–Zhang would have
done — fuzzed on the CPU.  The function's source text is taken verbatim from the .cu (blocks marked ``[host-test:…]``),
compiled for the host between shims (tests/csrc/fast_replay_host.cpp.in), fed sums computed in NumPy exactly as ``elem3``
defines them, and held against the oracle's full optimiser on the same unit: whenever the fast path accepts, the oracle must
have taken the same number of iterations and evaluations, ended with the same status, at the same ẑ and score."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SLOTS = ["rR0", "rS2_0", "rS1_0", "rGG0", "rR1", "rS2_1", "rDP1", "rRT", "rS2T", "rS1T", "rDPT", "rGGT", "rGM0", "rGMT", "rXC"]
START_ZERO, START_OWN, START_TRUTH, START_KEEP = 0, 1, 2, 4       # StartKind of muse_common.cuh


@pytest.fixture(scope="module")
def replay(tmp_path_factory):
    with open(os.path.join(ROOT, "museinference.jl_b200", "csrc", "muse_iso_stream.cu")) as fh:
        cu = fh.read()

    def block(tag):
        m = re.search(r"// \[host-test:begin %s\][^\n]*\n(.*?)// \[host-test:end %s\]" % (tag, tag), cu, flags=re.S)
        assert m, tag
        return m.group(1)

    with open(os.path.join(ROOT, "tests", "csrc", "fast_replay_host.cpp.in")) as fh:
        src = fh.read().replace("@RED_SLOTS@", block("red-slots")).replace("@FAST_REPLAY@", block("fast-replay"))
    d = tmp_path_factory.mktemp("fr")
    (d / "fast_replay_host.cpp").write_text(src)
    out = str(d / "libfr.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-I", os.path.join(ROOT, "museinference.jl_b200", "csrc"),
                    "-I", "/usr/local/cuda/include", "-o", out, str(d / "fast_replay_host.cpp")], check=True)
    lib = C.CDLL(out)
    assert lib.muse_host_nred() == 15 and [lib.muse_host_slot(i) for i in range(15)] == list(range(15))
    return lib


def _consts(family, theta, d):
    if family == "funnel":
        a, mu, half = np.exp(-theta[0]), 0.0, 0.5 * d * theta[0]
    else:
        a, mu, half = np.exp(-2.0 * theta[1]), theta[0], d * theta[1]
    return a, mu, half, 1.0 / (1.0 + a)


def _sums(x, z0, a, mu, cspec):
    """The 15 sums of one unit, as ``elem3`` accumulates them (up to summation order)."""
    r0, w0 = x - z0, z0 - mu
    g0 = a * w0 - r0
    z1 = z0 - g0
    r1, w1 = x - z1, z1 - mu
    g1 = a * w1 - r1
    zt = z0 - cspec * g0
    rt, wt = x - zt, zt - mu
    gt = a * wt - rt
    t = np.empty(15)
    t[0], t[1], t[2], t[3] = r0 @ r0, w0 @ w0, w0.sum(), g0 @ g0
    t[4], t[5], t[6] = r1 @ r1, w1 @ w1, -(g1 @ g0)
    t[7], t[8], t[9], t[10], t[11] = rt @ rt, wt @ wt, wt.sum(), -(gt @ g0), gt @ gt
    t[12], t[13], t[14] = np.abs(g0).max(), np.abs(gt).max(), np.abs(zt - z0).max()
    return t, zt


def _call(lib, T, A, MU, HALF, CSPEC, ATOL, MAXIT, SK, lean=False):
    n = len(T)
    f64 = lambda v: np.ascontiguousarray(v, dtype=np.float64)
    i32 = lambda v: np.ascontiguousarray(v, dtype=np.int32)
    out = np.zeros((n, 8))
    arrs = [f64(np.array(T)), f64(A), f64(MU), f64(HALF), f64(CSPEC), f64(ATOL), i32(MAXIT), i32(SK)]
    fn = lib.muse_host_fast_replay_lean if lean else lib.muse_host_fast_replay
    fn(C.c_int(n), *[a.ctypes.data_as(C.c_void_p) for a in arrs], out.ctypes.data_as(C.c_void_p))
    return out


def test_fast_replay_agrees_with_the_oracle_optimiser_on_random_units(replay):
    rng = np.random.default_rng(20261017)
    cases, T, A, MU, HALF, CSPEC, ATOL, MAXIT, SK = [], [], [], [], [], [], [], [], []
    for _ in range(3000):
        family = "funnel" if rng.random() < 0.5 else "hiergauss"
        d = int(rng.integers(2, 48))
        theta = np.array([rng.normal(0, 1.5)]) if family == "funnel" else np.array([rng.normal(0, 2), rng.normal(0, 0.8)])
        fam = O.make_family(family, d)
        xi, nu = rng.standard_normal(d), rng.standard_normal(d)
        x, ztrue = fam.sample(theta + (rng.normal(0, 0.3, theta.size) if rng.random() < 0.5 else 0.0), xi, nu)
        kind = rng.choice([START_ZERO, START_OWN, START_TRUTH, START_KEEP])
        if kind == START_ZERO:
            z0 = np.zeros(d)
        elif kind == START_TRUTH:
            z0 = ztrue
        elif rng.random() < 0.15:
            z0 = fam.exact_map(x, theta) + rng.normal(0, 1e-9, d)        # warm start at the MAP: 0 iterations
        else:
            z0 = rng.normal(0, 1.0, d) * rng.choice([1e-3, 1.0, 30.0])
        atol = float(rng.choice([1e-2, 1e-5, 1e-9, 1.0]))
        a, mu, half, cspec = _consts(family, theta, d)
        t, zt = _sums(x, z0, a, mu, cspec)
        cases.append((fam, x, z0, theta, atol, zt, kind))
        T.append(t); A.append(a); MU.append(mu); HALF.append(half); CSPEC.append(cspec); ATOL.append(atol); MAXIT.append(1000); SK.append(kind)
    out = _call(replay, T, A, MU, HALF, CSPEC, ATOL, MAXIT, SK)
    # The lean form (the one-launch solve: φ(1), φ′(1) from their closed forms a‖∇f‖², no max|Δz|) must take the decisions of the
    # form that evaluates the α = 1 trial element by element: whatever it accepts, the other accepts with the same outputs; what
    # the other accepts and it does not is a unit whose honest secant step sits at the edge of the speculation test, or one
    # accepted only because x did not change at all (status XF_CONVERGED), which the lean form hands back by design.
    lean = _call(replay, T, A, MU, HALF, CSPEC, ATOL, MAXIT, SK, lean=True)
    only_honest = only_lean = 0
    for (fam, x, z0, theta, atol, zt, kind), o, l, t, cs in zip(cases, out, lean, T, CSPEC):
        if l[0]:
            if not o[0]:
                # starts within round-off of the MAP: the element-by-element φ′(1) is noise, its secant step misses c_spec and the
                # unit is handed back; the closed form is exact — and the oracle's optimiser does finish these units the way the
                # lean form says (one iteration, three evaluations, at the same point)
                only_lean += 1
                dphi0, dphi1 = -t[3], t[6]
                c = -dphi0 / (dphi1 - dphi0)
                assert not (abs(c - cs) <= 1e-11 * cs and dphi1 >= 0), (c, cs)
                soln = O.lbfgs_minimize(lambda z: fam.neg_loglike_and_grad(x, z, theta), z0, g_tol=atol)
                if min(abs(t[12] - atol), abs(t[13] - atol)) >= 1e-9 * atol:
                    assert (int(l[5]), int(l[6])) == (soln.iterations, soln.f_calls), (l, soln.iterations, soln.f_calls)
                    np.testing.assert_allclose(zt if int(l[5]) == 1 else z0, soln.minimizer, rtol=1e-9, atol=1e-10)
            else:
                np.testing.assert_array_equal(o, l)
        elif o[0]:
            only_honest += 1
            assert int(o[7]) == 1, o                                                                  # accepted through x_conv only
    assert only_lean < 150 and only_honest < 30, (only_lean, only_honest)
    accepted = one_iter = zero_iter = near = 0
    for (fam, x, z0, theta, atol, zt, kind), o, t in zip(cases, out, T):
        soln = O.lbfgs_minimize(lambda z: fam.neg_loglike_and_grad(x, z, theta), z0, g_tol=atol)
        if not o[0]:
            # Handed back — always safe (the generic kernel re-solves the unit).  A 0-iteration solve from a start that has to
            # be kept (truth / user z₀) is handed back by design.  Otherwise, if the oracle DID finish in the one speculated
            # iteration, the hand-back must be explained by the speculation test itself: the secant step computed from the
            # sums differs from c_spec by more than kSpecTol (starts within round-off of the MAP: tiny, noisy gradients).
            if soln.iterations == 1 and soln.f_calls == 3 and kind not in (START_TRUTH, START_KEEP):
                dphi0, dphi1 = -t[3], t[6]
                c = (0.0 * dphi1 - 1.0 * dphi0) / (dphi1 - dphi0)
                cspec = 1.0 / (1.0 + (np.exp(-theta[0]) if fam.name == "funnel" else np.exp(-2.0 * theta[1])))
                assert abs(c - cspec) > 1e-11 * cspec, (c, cspec, atol, t[12])
                near += 1
            continue
        accepted += 1
        iters, fg, status = int(o[5]), int(o[6]), int(o[7])
        if min(abs(t[12] - atol), abs(t[13] - atol)) < 1e-9 * atol:
            near += 1
            continue
        assert (iters, fg) == (soln.iterations, soln.f_calls), (iters, fg, soln.iterations, soln.f_calls, atol, t[12], t[13])
        assert status == (0 if soln.g_converged else 1)
        zhat = zt if iters == 1 else z0
        np.testing.assert_allclose(zhat, soln.minimizer, rtol=1e-9, atol=1e-10)
        scale = max(1.0, np.abs(x).max(), np.abs(z0).max())
        np.testing.assert_allclose(o[2], soln.g_residual, rtol=1e-6, atol=1e-11 * scale)      # at the MAP the residual is round-off
        np.testing.assert_allclose(o[1], soln.minimum, rtol=1e-10, atol=1e-10)
        s1, s2 = o[3], o[4]
        score = fam.score(x, zhat, theta)
        a_ = np.exp(-theta[0]) if fam.name == "funnel" else np.exp(-2.0 * theta[1])
        mine = np.array([0.5 * a_ * s2 - 0.5 * fam.d]) if fam.name == "funnel" else np.array([a_ * s1, a_ * s2 - fam.d])
        np.testing.assert_allclose(mine, score, rtol=1e-9, atol=1e-9 * scale ** 2)
        one_iter += iters == 1
        zero_iter += iters == 0
    assert accepted > 2000 and one_iter > 1500 and zero_iter > 50 and near < 100, (accepted, one_iter, zero_iter, near)


def test_fast_replay_hands_back_what_it_cannot_finish(replay):
    """atol below round-off (a second iteration would follow), an iteration cap of 0, non-finite sums, a 0-iteration solve
    whose start must be kept: the unit goes back to the generic kernel (returns false), except that a non-finite objective
    at a start that need not be kept is reported at once (status NONFINITE, 0 iterations)."""
    rng = np.random.default_rng(5)
    d, theta = 32, np.array([0.4])
    fam = O.make_family("funnel", d)
    x, _ = fam.sample(theta, rng.standard_normal(d), rng.standard_normal(d))
    a, mu, half, cspec = _consts("funnel", theta, d)
    t, _ = _sums(x, np.zeros(d), a, mu, cspec)
    out = _call(replay, [t, t, t], [a] * 3, [mu] * 3, [half] * 3, [cspec] * 3, [1e-300, 1e-2, 1e-2], [1000, 0, 1000], [START_ZERO] * 3)
    assert out[0, 0] == 0 and out[1, 0] == 0 and out[2, 0] == 1 and out[2, 5] == 1 and out[2, 6] == 3
    tn = t.copy(); tn[0] = np.nan
    out = _call(replay, [tn, tn], [a] * 2, [mu] * 2, [half] * 2, [cspec] * 2, [1e-2] * 2, [1000] * 2, [START_ZERO, START_TRUTH])
    assert out[0, 0] == 1 and out[0, 7] == 4 and out[0, 5] == 0          # reported: NONFINITE, nothing to keep
    assert out[1, 0] == 0                                                 # the truth start has to be materialised: generic kernel
    zmap = fam.exact_map(x, theta)
    t0, _ = _sums(x, zmap, a, mu, cspec)
    out = _call(replay, [t0, t0], [a] * 2, [mu] * 2, [half] * 2, [cspec] * 2, [1e-2] * 2, [1000] * 2, [START_OWN, START_KEEP])
    assert out[0, 0] == 1 and out[0, 5] == 0 and out[0, 6] == 1 and out[0, 7] == 0
    assert out[1, 0] == 0
    # a speculated step that is not the secant step (here: c_spec off by 1e-6) is never accepted
    tb, _ = _sums(x, np.zeros(d), a, mu, cspec * (1 + 1e-6))
    out = _call(replay, [tb], [a], [mu], [half], [cspec * (1 + 1e-6)], [1e-2], [1000], [START_ZERO])
    assert out[0, 0] == 0
