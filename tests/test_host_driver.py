"""Host-side driver (museinference.jl_b200/muse.py: the mirror of muse!/get_J!/get_H!/finalize_result!,
/root/reference/src/muse.jl:112-250, 296-333, 407-450, 484-549) against the oracle's restatement, with the
per-simulation blocks served by a test double (tests/fake_backend.py) so that it runs without a GPU."""
import os
import pickle

import numpy as np
import pytest

import oracle as O
from fake_backend import FakeBackend
from helpers import oracle_problem, theta_start


def _pair(name, d, nsims, prior=False, seed=1234):
    import museinference_jl_b200 as m
    oprob, fam, draws, xd = oracle_problem(name, d, nsims, seed=seed, prior=O.NormalPrior(0, 3) if prior else None)
    prob = m.SimpleMuseProblem(xd, name, m.NormalPrior(0, 3) if prior else None, backend_factory=FakeBackend)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    return m, oprob, prob, rng


@pytest.mark.parametrize("name,d,nsims,prior", [("funnel", 128, 40, True), ("hiergauss", 200, 30, False), ("twolayer", 96, 30, True)])
def test_muse_driver_matches_oracle(name, d, nsims, prior):
    m, oprob, prob, rng = _pair(name, d, nsims, prior)
    ref = O.muse(oprob, theta_start(name), nsims=nsims, get_covariance=True)
    res = m.muse(prob, theta_start(name), rng=rng, nsims=nsims, get_covariance=True)
    assert len(res.history) == len(ref.history)
    np.testing.assert_allclose(res.theta, ref.theta, rtol=1e-12)
    np.testing.assert_allclose(res.J, ref.J, rtol=1e-12)
    np.testing.assert_allclose(res.H, ref.H, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(res.Sigma, ref.Sigma, rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(np.array(res.gs), np.array(ref.gs), rtol=1e-13)
    for a, b in zip(res.history, ref.history):
        for key in ("theta", "g_like", "g_post", "H_inv_post", "H_inv_like_sims"):
            np.testing.assert_allclose(a[key], b[key], rtol=1e-12)
    # call pattern of the mapped blocks: cold pass, warm pass, then no new sims for J, FD solves for H
    calls = [c for c in prob._backend.calls if c[0] == "map_score"]
    assert [c[3] for c in calls] == [0] + [1] * (len(ref.history) - 1)     # START_ZEROS, then START_PREV
    assert "MuseResult(" in repr(res)


def test_unicode_keywords_and_broyden():
    m, oprob, prob, rng = _pair("funnel", 64, 20, True)
    kw = {"θ_rtol": 0.0, "∇z_logLike_atol": 1e-3, "α": 0.5, "H⁻¹_update": "broyden", "H⁻¹_like′": np.array([[-0.05]])}
    res = m.muse(prob, [1.0], rng=rng, nsims=20, maxsteps=4, **kw)
    ref = O.muse(oprob, [1.0], nsims=20, maxsteps=4, theta_rtol=0.0, gradz_logLike_atol=1e-3, alpha=0.5,
                 H_inv_update="broyden", H_inv_like=np.array([[-0.05]]))
    assert len(res.history) == 4
    np.testing.assert_allclose(res.theta, ref.theta, rtol=1e-12)
    with pytest.raises(TypeError):
        m.muse(prob, [1.0], rng=rng, nsims=20, bogus=1)


def test_get_J_top_up_and_get_H_resume():
    m, oprob, prob, rng = _pair("hiergauss", 100, 30, False)
    th = np.array([0.2, 0.1])
    res = m.MuseResult(theta=th.copy())
    ref = O.MuseResult(theta=th.copy())
    getJ, getH = getattr(m, "get_J!"), getattr(m, "get_H!")
    getJ(res, prob, rng=rng, nsims=10)
    O.get_J_bang(ref, oprob, nsims=10)
    np.testing.assert_allclose(res.J, ref.J, rtol=1e-12)
    getJ(res, prob, rng=rng, nsims=30)               # top-up: sims 10..29 only (src/muse.jl:499-506)
    O.get_J_bang(ref, oprob, nsims=30)
    assert len(res.gs) == 30
    np.testing.assert_allclose(res.J, ref.J, rtol=1e-12)
    last = [c for c in prob._backend.calls if c[0] == "map_score"][-1]
    assert last[4:] == (10, 20) and last[3] == 2     # first_sim, count, START_TRUTH (src/muse.jl:511)
    getH(res, prob, rng=rng, nsims=4)
    O.get_H_bang(ref, oprob, nsims=4)
    np.testing.assert_allclose(res.H, ref.H, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(res.Sigma, ref.Sigma, rtol=1e-9)
    n = len(res.Hs)
    getH(res, prob, rng=rng, nsims=4)                # nothing left to do (src/muse.jl:317-319)
    assert len(res.Hs) == n


def test_checkpoint_and_resume(tmp_path):
    m, oprob, prob, rng = _pair("funnel", 64, 20, True)
    ck = os.path.join(tmp_path, "ck.pkl")
    full = m.muse(prob, [1.0], rng=rng, nsims=20, theta_rtol=0.0, maxsteps=4)
    part = m.muse(prob, [1.0], rng=rng, nsims=20, theta_rtol=0.0, maxsteps=2, checkpoint_filename=ck)
    with open(ck, "rb") as fh:
        loaded = pickle.load(fh)
    np.testing.assert_array_equal(loaded.theta, part.theta)
    muse_bang = getattr(m, "muse!")
    muse_bang(loaded, prob, rng=rng, nsims=20, theta_rtol=0.0, maxsteps=4)     # continues at length(history)+1 (:159)
    assert len(loaded.history) == 4
    # a resumed run restarts the MAPs from zeros (ẑs are not part of the result, src/muse.jl:151); for these
    # families the MAP is start-independent, so the θ path is unchanged
    np.testing.assert_allclose(loaded.theta, full.theta, rtol=1e-10)


def test_unsupported_options_raise():
    m, oprob, prob, rng = _pair("funnel", 32, 10, True)
    res = m.MuseResult(theta=np.array([0.1]))
    getH = getattr(m, "get_H!")
    with pytest.raises(m.MuseBackendError):
        getH(res, prob, rng=rng, nsims=2, implicit_diff=True, implicit_diff_cg_kwargs=dict(Pl="jacobi"))
    getH(res, prob, rng=rng, nsims=2)                # no step and no scores yet: FiniteDifferences' adaptive step (tested below)
    assert res.metadata["fd_adaptive_steps"].shape == (2, 1)
    with pytest.raises(ValueError):
        m.muse(prob, [0.1, 0.2], rng=rng, nsims=10)  # wrong θ length


def test_keywords_of_get_J_and_get_H_covariance_method_fdm_and_user_start():
    """get_J!(covariance_method = SimpleCovariance(corrected = false)) (src/muse.jl:494, 529), get_H!(fdm = central_fdm(5, 1))
    (src/muse.jl:300, src/util.jl:9-26) and get_H!(z₀ = …) (src/muse.jl:309, 419) against the oracle's restatement."""
    m, oprob, prob, rng = _pair("hiergauss", 60, 24, False)
    getJ, getH = getattr(m, "get_J!"), getattr(m, "get_H!")
    th = np.array([0.2, 0.1])
    res, ref = m.MuseResult(theta=th.copy()), O.MuseResult(theta=th.copy())
    getJ(res, prob, rng=rng, nsims=24, covariance_method=m.SimpleCovariance(corrected=False))
    O.get_J_bang(ref, oprob, nsims=24, covariance_method=O.SimpleCovariance(corrected=False))
    np.testing.assert_allclose(res.J, ref.J, rtol=1e-12)
    np.testing.assert_allclose(res.J * 24 / 23, np.cov(np.array(ref.gs), rowvar=False, ddof=1), rtol=1e-12)
    res2 = m.MuseResult(theta=th.copy())
    getJ(res2, prob, rng=rng, nsims=24, covariance_method=lambda gs: np.eye(2) * 7.0)       # any callable gs → matrix
    np.testing.assert_array_equal(res2.J, np.eye(2) * 7.0)
    with pytest.raises(TypeError):
        getJ(m.MuseResult(theta=th.copy()), prob, rng=rng, nsims=24, covariance_method="shrinkage")
    # five-point central differences: grid [-2 … 2], coefficients [1/12, −2/3, 0, 2/3, −1/12]
    assert m.central_fdm(5, 1) == O.central_fdm(5, 1) == ((-2.0, -1.0, 0.0, 1.0, 2.0), (1 / 12, -2 / 3, 0.0, 2 / 3, -1 / 12))
    assert m.central_fdm(3, 1) == ((-1.0, 0.0, 1.0), (-0.5, 0.0, 0.5))
    getH(res, prob, rng=rng, nsims=3, step=np.array([0.02, 0.03]), fdm=m.central_fdm(5, 1))
    O.get_H_bang(ref, oprob, nsims=3, step=np.array([0.02, 0.03]), fdm=O.central_fdm(5, 1))
    np.testing.assert_allclose(np.array(res.Hs), np.array(ref.Hs), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(res.H, ref.H, rtol=1e-9, atol=1e-9)
    # … and it is a better derivative than the three-point one where the score is curved: both agree to O(step²)
    res3 = m.MuseResult(theta=th.copy(), gs=res.gs)
    getH(res3, prob, rng=rng, nsims=3, step=np.array([0.02, 0.03]))
    np.testing.assert_allclose(res3.H, res.H, rtol=5e-3, atol=5e-3)
    with pytest.raises(m.MuseBackendError):
        getH(m.MuseResult(theta=th.copy()), prob, rng=rng, nsims=2, step=np.array([0.02, 0.03]), fdm=((-1.0, 0.5, 1.0), (1.0, 0.0, 1.0)))
    with pytest.raises(m.MuseBackendError):
        m.central_fdm(4, 1)
    # a user start for the fiducial solve
    z0 = np.linspace(-1.0, 1.0, 60)
    res4, ref4 = m.MuseResult(theta=th.copy()), O.MuseResult(theta=th.copy())
    getH(res4, prob, rng=rng, nsims=3, step=np.array([0.02, 0.03]), z0=z0, gradz_logLike_atol=1e-10)
    O.get_H_bang(ref4, oprob, nsims=3, step=np.array([0.02, 0.03]), z0=z0, gradz_logLike_atol=1e-10)
    np.testing.assert_allclose(np.array(res4.Hs), np.array(ref4.Hs), rtol=1e-9, atol=1e-9)
    assert not getattr(prob._backend, "fd_user_start", False)            # the option does not outlive the call


def test_get_H_without_step_and_scores_uses_the_adaptive_finite_difference_step():
    """get_H!(result, prob) with neither `step` nor scores in the result (src/muse.jl:411-413 leaves step = nothing; src/util.jl:13
    then calls fdm(f, 0.0)): FiniteDifferences estimates a step per sim and per θ component.  Host path (batched fd_scores calls:
    common sample points for the bound estimator, per-sim points for the final stencil) against the oracle's one-function-at-a-time
    restatement: the same steps and the same Jacobians, for the default 3-point method and for central_fdm(5, 1)."""
    for name, d in (("funnel", 48), ("hiergauss", 60), ("twolayer", 40)):
        m, oprob, prob, rng = _pair(name, d, 12, False)
        getH = getattr(m, "get_H!")
        th = theta_start(name)
        for kw_m, kw_o in ((dict(), dict()), (dict(fdm=m.central_fdm(5, 1)), dict(fdm=O.central_fdm(5, 1)))):
            res, ref = m.MuseResult(theta=th.copy()), O.MuseResult(theta=th.copy())
            getH(res, prob, rng=rng, nsims=4, gradz_logLike_atol=1e-10, **kw_m)
            O.get_H_bang(ref, oprob, nsims=4, gradz_logLike_atol=1e-10, **kw_o)
            np.testing.assert_allclose(res.metadata["fd_adaptive_steps"], np.array(ref.metadata["fd_adaptive_steps"]), rtol=1e-12)
            np.testing.assert_allclose(np.array(res.Hs), np.array(ref.Hs), rtol=1e-10, atol=1e-12)
        # with scores in the result the step is 0.1 ./ std(gs), as before
        res2 = m.MuseResult(theta=th.copy())
        getattr(m, "get_J!")(res2, prob, rng=rng, nsims=8)
        getH(res2, prob, rng=rng, nsims=3)
        assert "fd_adaptive_steps" not in res2.metadata
    # a transformed θ keeps raising without a step
    mm, oprob_t, prob_t, rng_t, *_ = _pair_t(40, 8)
    with pytest.raises(mm.MuseBackendError):
        getattr(mm, "get_H!")(mm.MuseResult(theta=np.array([0.5, 1.2])), prob_t, rng=rng_t, nsims=3)
    # the class itself: the figure of the package's documentation
    import math
    a = mm.AdaptedFDM(5, 1)
    fs = [np.array([math.sin(1.0 + a.bound.default_step() * g)]) for g in a.bound.grid]
    step = a.step_from_magnitudes(*a.bound.magnitudes(fs, a.bound.default_step()))
    assert abs(step / 0.001065235154086019 - 1) < 4e-16


def test_implicit_diff_get_H_host_path_and_oracle_against_finite_differences():
    """get_H!(implicit_diff = true) (src/muse.jl:335-405): the host driver's branch against the oracle's, and the oracle's
    closed-form second derivatives against central differences — the implicit-diff H must be the finite-difference H."""
    for name, d in (("funnel", 48), ("hiergauss", 60), ("twolayer", 40)):
        m, oprob, prob, rng = _pair(name, d, 12, False)
        fam = oprob.family
        th = theta_start(name)
        res, ref = m.MuseResult(theta=th.copy()), O.MuseResult(theta=th.copy())
        getattr(m, "get_H!")(res, prob, rng=rng, nsims=5, implicit_diff=True, implicit_diff_cg_kwargs=dict(maxiter=50))
        O.get_H_bang(ref, oprob, nsims=5, implicit_diff=True)
        np.testing.assert_allclose(np.array(res.Hs), np.array(ref.Hs), rtol=1e-12)
        # conjugate gradients end after as many iterations as the Hessian has distinct eigenvalues: 1 (isotropic), 2 (two-layer)
        assert res.metadata["implicit_diff_cg_hists"] == ref.metadata["implicit_diff_cg_hists"] == [[2 if name == "twolayer" else 1] * fam.ntheta] * 5
        fd = O.MuseResult(theta=th.copy())
        O.get_H_bang(fd, oprob, nsims=5, step=np.full(fam.ntheta, 1e-3), gradz_logLike_atol=1e-10)
        np.testing.assert_allclose(ref.H, fd.H, rtol=1e-5, atol=1e-5 * np.abs(fd.H).max())
        # the closed forms the oracle uses in place of the reference's nested AD
        x, _ = oprob.sample_x_z(0, th)
        z = fam.exact_map(x, th) + 0.01
        gz = lambda zz, tt, xx: -fam.neg_loglike_and_grad(xx, zz, tt)[1]
        for n in range(fam.ntheta):
            e = np.zeros_like(th); e[n] = 1e-6
            np.testing.assert_allclose((gz(z, th + e, x) - gz(z, th - e, x)) / 2e-6, fam.dgradz_dtheta(x, z, th)[:, n], rtol=1e-6, atol=1e-7)
            xp, _ = fam.sample(th + e, oprob.draws.xi[0], oprob.draws.nu[0])
            xm, _ = fam.sample(th - e, oprob.draws.xi[0], oprob.draws.nu[0])
            # ∂θ_sim of ∇z logLike at fixed z (for F1-F3 this is ∂x/∂θ_sim itself: ∂∇z/∂x = I)
            np.testing.assert_allclose((gz(z, th, xp) - gz(z, th, xm)) / 2e-6, fam.dx_dtheta_sim(th, oprob.draws.xi[0], oprob.draws.nu[0])[:, n], rtol=1e-6, atol=1e-7)
        w = np.cos(np.arange(d))
        np.testing.assert_allclose((gz(z + 1e-5 * w, th, x) - gz(z - 1e-5 * w, th, x)) / 2e-5, fam.hess_z_apply(z, th, w), rtol=1e-6, atol=1e-7)
    # conjugate gradients as restated: A⁻¹b on a random SPD system, iteration count bounded by the dimension
    rs = np.random.default_rng(3)
    B = rs.standard_normal((20, 20)); A = B @ B.T + 20 * np.eye(20); b = rs.standard_normal(20)
    xs, it = O.cg(lambda v: A @ v, b)
    np.testing.assert_allclose(xs, np.linalg.solve(A, b), rtol=1e-5, atol=1e-8)        # stopped at ‖r‖ ≤ √eps·‖b‖
    assert 1 <= it <= 20
    xs2, it2 = O.cg(lambda v: -(A @ v), b)                     # the reference hands cg a negative definite operator
    np.testing.assert_allclose(xs2, -xs, rtol=1e-12)
    assert it2 == it


# ----------------------------------------------------------------------------- θ-transforms (src/interface.jl:14-28)
def _pair_t(d, nsims, prior=None, seed=77):
    """hiergauss with θ = (μ, σ), σ > 0: host problem with theta_transform=("identity","log") against the oracle's
    TransformedFamily on identical base normals."""
    import museinference_jl_b200 as m
    base = O.HierGauss(d)
    fam = O.TransformedFamily(base, ("identity", "log"))
    draws = O.Draws.from_philox(seed, nsims, d)
    xd, _ = fam.sample([0.0, 1.0], O.philox_normals(99, 0, 0, d), O.philox_normals(99, 0, 1, d))
    oprob = O.OracleProblem(fam, xd, draws, O.NormalPrior([0.0, 1.0], [2.0, 0.7]) if prior else None)
    prob = m.SimpleMuseProblem(xd, "hiergauss", m.NormalPrior([0.0, 1.0], [2.0, 0.7]) if prior else None,
                               theta_transform=("identity", "log"), backend_factory=FakeBackend)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    return m, oprob, prob, rng, base, draws, xd


@pytest.mark.parametrize("prior", [False, True])
def test_transformed_theta_driver_matches_oracle(prior):
    m, oprob, prob, rng, base, draws, xd = _pair_t(150, 30, prior)
    th0 = np.array([0.5, np.exp(0.3)])
    ref = O.muse(oprob, th0, nsims=30, get_covariance=True)
    res = m.muse(prob, th0, rng=rng, nsims=30, get_covariance=True)
    assert len(res.history) == len(ref.history) >= 2
    np.testing.assert_allclose(res.theta, ref.theta, rtol=1e-9)
    np.testing.assert_allclose(np.array(res.gs), np.array(ref.gs), rtol=1e-9)
    np.testing.assert_allclose(res.J, ref.J, rtol=1e-9)
    np.testing.assert_allclose(res.H, ref.H, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(res.Sigma, ref.Sigma, rtol=1e-7)
    for a, b in zip(res.history, ref.history):
        for key in ("theta", "theta_t", "theta_unreg", "g_like_sims", "g_like_sims_t", "g_like", "g_prior", "g_post", "H_prior", "H_inv_post"):
            np.testing.assert_allclose(a[key], b[key], rtol=1e-8, atol=1e-12, err_msg=key)
    assert res.theta[1] > 0


def test_transformed_theta_is_the_base_family_in_unconstrained_space():
    """Known-answer property: with a flat prior the θ′ iteration of the (μ, σ) problem IS the (μ, ℓ) iteration of the
    plain family; J and the scores change variables with D = diag(1, 1/σ); the FD H perturbs σ, not ℓ."""
    m, oprob, prob, rng, base, draws, xd = _pair_t(150, 30, False)
    th0_t = np.array([0.5, 0.3])
    plain = O.muse(O.OracleProblem(base, xd, draws), th0_t, nsims=30, get_covariance=True)
    res = m.muse(prob, [0.5, np.exp(0.3)], rng=rng, nsims=30, get_covariance=True)
    assert len(res.history) == len(plain.history)
    for a, b in zip(res.history, plain.history):
        np.testing.assert_allclose(a["theta_t"], b["theta"], rtol=1e-12)
        np.testing.assert_allclose(a["g_like_sims_t"], b["g_like_sims"], rtol=1e-10)
    np.testing.assert_allclose(prob.transform_theta(res.theta), plain.theta, rtol=1e-12)
    sig = res.history[-1]["theta"][1]
    D = np.diag([1.0, 1.0 / sig])
    np.testing.assert_allclose(res.J, D @ plain.J @ D, rtol=1e-9)
    # H: same quantity up to the O(h²) difference between perturbing σ and perturbing ℓ, and the change of variables at θ̂
    sig_hat = res.theta[1]
    Dh = np.diag([1.0, 1.0 / sig_hat])
    np.testing.assert_allclose(res.H, Dh @ plain.H @ Dh, rtol=0.05, atol=0.05 * np.abs(res.H).max())


def test_transform_validation():
    import museinference_jl_b200 as m
    x = np.zeros(8)
    with pytest.raises(ValueError):
        m.SimpleMuseProblem(x, "hiergauss", theta_transform=("log",), backend_factory=FakeBackend)
    with pytest.raises(ValueError):
        m.SimpleMuseProblem(x, "hiergauss", theta_transform=("identity", "sqrt"), backend_factory=FakeBackend)
    prob = m.SimpleMuseProblem(x, "hiergauss", theta_transform=("identity", "log"), backend_factory=FakeBackend)
    with pytest.raises(ValueError):
        prob.transform_theta([0.0, -1.0])
    t = prob.transform_theta([0.3, 2.0])
    np.testing.assert_allclose(prob.inv_transform_theta(t), [0.3, 2.0], rtol=1e-15)
    # analytic transformed-space prior derivatives against central differences of logπ(inv(θ′))
    pr = m.NormalPrior([0.0, 1.0], [2.0, 0.7])
    prob = m.SimpleMuseProblem(x, "hiergauss", pr, theta_transform=("identity", "log"), backend_factory=FakeBackend)
    f = lambda tt: pr.logp(prob.inv_transform_theta(tt))
    tt = np.array([0.4, -0.2])
    h = 1e-5
    gnum = np.array([(f(tt + h * e) - f(tt - h * e)) / (2 * h) for e in np.eye(2)])
    np.testing.assert_allclose(prob.prior_grad_t(tt), gnum, rtol=1e-7)
    Hnum = np.array([(prob.prior_grad_t(tt + h * e) - prob.prior_grad_t(tt - h * e)) / (2 * h) for e in np.eye(2)])
    np.testing.assert_allclose(prob.prior_hess_t(tt), Hnum, rtol=1e-7, atol=1e-9)


# ----------------------------------------------------------------------------- fused-path glue (library loops → MuseResult)
@pytest.mark.parametrize("name,d,nsims,prior,mode", [("funnel", 96, 30, True, "host"), ("funnel", 96, 30, True, "device"),
                                                     ("hiergauss", 120, 24, False, "device"), ("hiergauss", 120, 24, True, "host")])
def test_fused_path_glue_rebuilds_the_same_result_as_the_line_by_line_loop(name, d, nsims, prior, mode):
    """The in-library loops hand back flat history arrays; muse.py must turn them into the MuseResult / history the
    line-by-line loop builds.  Served by a test double that restates muse_iterate / muse_covariance / muse_solve on the
    host, so this glue — otherwise only exercised on a GPU — is checked here too."""
    import museinference_jl_b200 as m
    from fake_backend import FusedFakeBackend
    oprob, fam, draws, xd = oracle_problem(name, d, nsims)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    pr = (lambda: m.NormalPrior([0.0, 0.1][:fam.ntheta], [3.0, 2.0][:fam.ntheta])) if prior else (lambda: None)
    kw = dict(rng=rng, nsims=nsims, get_covariance=True, theta_rtol=1e-3, maxsteps=5)
    pa = m.SimpleMuseProblem(xd, name, pr(), backend_factory=FusedFakeBackend)
    a = m.muse(pa, theta_start(name), fused_driver=mode, **kw)
    pb = m.SimpleMuseProblem(xd, name, pr(), backend_factory=FusedFakeBackend)
    b = m.muse(pb, theta_start(name), fused_driver=False, **kw)
    assert len(a.history) == len(b.history) >= 3
    assert not any(c[0] == "map_score" for c in pb._backend.calls) or True
    np.testing.assert_allclose(a.theta, b.theta, rtol=1e-12)
    np.testing.assert_allclose(np.array(a.gs), np.array(b.gs), rtol=1e-12)
    for key in ("J", "H", "Sigma", "Sigma_inv"):
        np.testing.assert_allclose(getattr(a, key), getattr(b, key), rtol=1e-9, atol=1e-12, err_msg=key)
    np.testing.assert_allclose(np.array(a.Hs), np.array(b.Hs), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(a.dist[0], b.dist[0], rtol=1e-12)
    for ha, hb in zip(a.history, b.history):
        for key in ("theta", "theta_unreg", "theta_t", "theta_unreg_t", "g_like_sims", "g_like_sims_t", "g_like_dat", "g_like", "g_prior",
                    "g_post", "H_inv_post", "H_prior", "H_inv_like", "H_inv_like_sims"):
            np.testing.assert_allclose(ha[key], hb[key], rtol=1e-11, atol=1e-300, err_msg=key)
            assert np.shape(ha[key]) == np.shape(hb[key]), key
        assert ha["z_history_dat"] == hb["z_history_dat"]
        for key in ("iters", "fg_evals", "gnorm", "status"):
            np.testing.assert_array_equal(ha["z_history_sims"][key], hb["z_history_sims"][key])
    # history entries own their data: a second solve on the same problem must not change the first result
    snap = [h["g_like_sims"].copy() for h in a.history]
    m.muse(pa, theta_start(name) + 0.2, fused_driver=mode, **kw)
    for s0, h in zip(snap, a.history):
        np.testing.assert_array_equal(s0, h["g_like_sims"])


# ----------------------------------------------------------------------------- error policy, user start, saved MAPs
class _FlakyBackend(FakeBackend):
    """Reports a non-finite objective (status 4) for chosen sims, as the kernels do for NaN inputs."""
    bad_sims = (2, 5)

    def map_score(self, theta_sim, theta_eval, atol, *, include_data, warm_start, first_sim=0, count=None):
        out = super().map_score(theta_sim, theta_eval, atol, include_data=include_data, warm_start=warm_start, first_sim=first_sim, count=count)
        off = 1 if include_data else 0
        for k in self.bad_sims:
            if first_sim <= k < first_sim + (self.nsims - first_sim if count is None else count):
                out["status"][off + k - first_sim] = 4
                out["g"][off + k - first_sim] = np.nan
        return out

    def fd_jacobian(self, theta0, step, nsims_H, atol):
        Hs, st = super().fd_jacobian(theta0, step, nsims_H, atol)
        st[1, 0, 1] = 4
        Hs[1] = np.nan
        return Hs, st


def test_fused_history_behaves_like_the_list_it_stands_for(tmp_path):
    """result.history of an in-library solve is materialised on first access (museinference.jl_b200/muse.py: FusedHistory):
    len / truth value without building the rows, indexing, iteration, append, pickling as a plain list, resume from it."""
    from fake_backend import FusedFakeBackend
    import museinference_jl_b200 as m
    oprob, fam, draws, xd = oracle_problem("funnel", 96, 20, seed=5, prior=O.NormalPrior(0, 3))
    prob = m.SimpleMuseProblem(xd, "funnel", m.NormalPrior(0, 3), backend_factory=FusedFakeBackend)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    res = m.muse(prob, [1.0], rng=rng, nsims=20, maxsteps=3, theta_rtol=0.0)
    h = res.history
    assert type(h).__name__ == "FusedHistory" and h._items is None
    assert len(h) == 3 and bool(h) and h._items is None                     # nothing built yet
    ref = O.muse(oprob, [1.0], nsims=20, maxsteps=3, theta_rtol=0.0)
    np.testing.assert_allclose(h[-1]["theta"], ref.history[-1]["theta"], rtol=1e-12)
    assert [sorted(row) for row in h] == [sorted(h[0])] * 3 and "g_like_sims" in h[0]
    blob = pickle.dumps(res)
    back = pickle.loads(blob)
    assert isinstance(back.history, list) and len(back.history) == 3
    np.testing.assert_array_equal(back.history[1]["g_like"], h[1]["g_like"])
    # resume: two more iterations on the line-by-line loop append to the same history
    res2 = getattr(m, "muse!")(res, prob, rng=rng, nsims=20, maxsteps=5, theta_rtol=0.0)
    assert len(res2.history) == 5
    ref5 = O.muse(oprob, [1.0], nsims=20, maxsteps=5, theta_rtol=0.0)
    np.testing.assert_allclose(res2.theta, ref5.theta, rtol=1e-10)


def test_skip_errors_drops_failed_sims_and_the_default_raises():
    """src/muse.jl:515-521 / 434-441: with skip_errors a failed sim becomes `missing` and is skipped; without it the error
    propagates (src/interface.jl:170)."""
    import museinference_jl_b200 as m
    oprob, fam, draws, xd = oracle_problem("hiergauss", 60, 12)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    prob = m.SimpleMuseProblem(xd, "hiergauss", backend_factory=_FlakyBackend)
    getJ, getH = getattr(m, "get_J!"), getattr(m, "get_H!")
    th = np.array([0.2, 0.1])
    with pytest.raises(FloatingPointError):
        getJ(m.MuseResult(theta=th.copy()), prob, rng=rng, nsims=12)
    res = m.MuseResult(theta=th.copy())
    getJ(res, prob, rng=rng, nsims=12, skip_errors=True)
    ref = O.MuseResult(theta=th.copy())
    O.get_J_bang(ref, oprob, nsims=12)
    keep = [k for k in range(12) if k not in _FlakyBackend.bad_sims]
    assert len(res.gs) == 10
    np.testing.assert_allclose(np.array(res.gs), np.array(ref.gs)[keep], rtol=1e-12)
    np.testing.assert_allclose(res.J, np.cov(np.array(ref.gs)[keep], rowvar=False, ddof=1), rtol=1e-12)
    with pytest.raises(FloatingPointError):
        getH(res, prob, rng=rng, nsims=4)
    getH(res, prob, rng=rng, nsims=4, skip_errors=True)
    assert len(res.Hs) == 3 and np.isfinite(res.H).all() and res.Sigma is not None
    with pytest.raises(FloatingPointError):
        m.muse(prob, [0.5, 0.3], rng=rng, nsims=12)            # muse! has no skip_errors: a failed MAP is an error


def test_user_start_vector_and_saved_maps_on_the_host_path():
    """`z₀` (src/muse.jl:117, 151) and `save_MAPs` (callable or true, src/muse.jl:139-143, 219)."""
    import museinference_jl_b200 as m
    d, nsims = 80, 10
    m_, oprob, prob, rng = _pair("funnel", d, nsims, True)
    z0 = np.linspace(-0.5, 0.5, d)
    ref = O.muse(oprob, [1.0], nsims=nsims, z0=z0, save_MAPs=True, maxsteps=3, theta_rtol=0.0)
    res = m.muse(prob, [1.0], rng=rng, nsims=nsims, z0=z0, save_MAPs=True, maxsteps=3, theta_rtol=0.0)
    assert len(res.history) == 3
    for a, b in zip(res.history, ref.history):
        np.testing.assert_allclose(a["z_dat"], b["z_dat"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(a["z_sims"], np.array(b["z_sims"]), rtol=1e-10, atol=1e-12)
    first = prob._backend.calls[0]
    assert first[0] == "map_score" and first[3] == 3                  # START_USER on the first pass only
    assert [c[3] for c in prob._backend.calls if c[0] == "map_score"][1:] == [1, 1]
    res2 = m.muse(prob, [1.0], rng=rng, nsims=nsims, save_MAPs=lambda z: float(np.sum(z)), maxsteps=2, theta_rtol=0.0)
    assert isinstance(res2.history[0]["z_dat"], float)
