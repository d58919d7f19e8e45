"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical base normals.

Tolerances (BASELINE.json north_star): per-sim ẑ_i and g_i within rtol 1e-8; θ̂, J, H, Σ within
rtol 1e-6.  All FP64.
"""
import os

import numpy as np
import pytest

import oracle as O
from helpers import make_inputs, oracle_problem, theta_start

pytestmark = pytest.mark.gpu

RTOL_SIM = 1e-8
RTOL_EST = 1e-6


def _backend(name, d, nsims, draws, xd, **kw):
    import museinference_jl_b200 as m
    be = m.B200Backend(name, d, nsims, **kw)
    be.set_data(xd)
    be.set_draws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    return be


def test_device_philox_matches_oracle():
    import museinference_jl_b200 as m
    d, n, seed, off = 1000, 7, 0xDEADBEEF12345, 40
    be = m.B200Backend("funnel", d, n, sim_offset=off)
    be.seed_draws(seed)
    xi, nu = be.get_draws(0, n + 1)
    for k in range(n):
        np.testing.assert_allclose(xi[k], O.philox_normals(seed, off + k, 0, d), rtol=0, atol=2e-13)
        np.testing.assert_allclose(nu[k], O.philox_normals(seed, off + k, 1, d), rtol=0, atol=2e-13)
    np.testing.assert_allclose(xi[n], O.philox_normals(seed, O.philox.MASTER_INDEX, 0, d), rtol=0, atol=2e-13)
    np.testing.assert_allclose(nu[n], O.philox_normals(seed, O.philox.MASTER_INDEX, 1, d), rtol=0, atol=2e-13)
    be.close()


# (family, d, group threads, cluster, kernel): kernel 1 = generic two-sweep solver only, 2 = single-pass TMA-ring streaming
# kernel first (the default for d ≥ 4096), 3 = its warp-per-unit form first (the default below), 0 = auto
GEOMS = [
    ("funnel", 512, 0, 0, 0), ("funnel", 513, 0, 0, 0), ("funnel", 512, 256, 1, 1), ("hiergauss", 700, 0, 0, 0),
    ("funnel", 5000, 0, 0, 0), ("hiergauss", 4096, 512, 1, 1), ("funnel", 4098, 512, 2, 1),
    ("funnel", 5000, 256, 1, 2), ("funnel", 5001, 512, 2, 2), ("hiergauss", 9000, 512, 1, 2),
    ("hiergauss", 4097, 0, 0, 0), ("funnel", 20001, 0, 0, 0), ("funnel", 300, 0, 0, 2), ("hiergauss", 70001, 0, 0, 0),
    ("funnel", 65536, 0, 0, 0), ("funnel", 600, 32, 0, 1), ("hiergauss", 5001, 0, 0, 3), ("funnel", 2, 0, 0, 3),
]


@pytest.mark.parametrize("name,d,group,cluster,kernel", GEOMS)
def test_map_score_cold_warm_truth(name, d, group, cluster, kernel):
    nsims = 24
    fam, draws, xd = make_inputs(name, d, nsims)
    prob = O.OracleProblem(fam, xd, draws)
    be = _backend(name, d, nsims, draws, xd, group=group, cluster=cluster, kernel=kernel)
    th0 = theta_start(name)
    atol = 1e-2

    # cold pass (src/muse.jl:169-176 with ẑs = zeros)
    out = be.map_score(th0, th0, atol, include_data=True, warm_start=0)
    zs = be.get_maps(0, nsims + 1)
    ref_z = []
    for u in range(nsims + 1):
        x = xd if u == 0 else prob.sample_x_z(u - 1, th0)[0]
        zh, g, soln = O.map_score_unit(prob, x, np.zeros(d), th0, atol)
        ref_z.append(zh)
        np.testing.assert_allclose(out["g"][u], g, rtol=RTOL_SIM)
        np.testing.assert_allclose(zs[u], zh, rtol=RTOL_SIM, atol=1e-12)
        assert out["iters"][u] == soln.iterations and out["fg_evals"][u] == soln.f_calls
        assert out["status"][u] == 0

    # warm pass at a moved θ (src/muse.jl:181: ẑs carried over)
    th1 = th0 - 0.3
    out = be.map_score(th1, th1, atol, include_data=True, warm_start=1)
    zs = be.get_maps(0, nsims + 1)
    for u in range(nsims + 1):
        x = xd if u == 0 else prob.sample_x_z(u - 1, th1)[0]
        zh, g, soln = O.map_score_unit(prob, x, ref_z[u], th1, atol)
        np.testing.assert_allclose(out["g"][u], g, rtol=RTOL_SIM)
        np.testing.assert_allclose(zs[u], zh, rtol=RTOL_SIM, atol=1e-12)
        assert out["iters"][u] == soln.iterations and out["fg_evals"][u] == soln.f_calls

    # truth start on a sub-range (src/muse.jl:508-514)
    out = be.map_score(th1, th1, atol, include_data=False, warm_start=2, first_sim=5, count=11)
    for i in range(11):
        x, z = prob.sample_x_z(5 + i, th1)
        zh, g, soln = O.map_score_unit(prob, x, z, th1, atol)
        np.testing.assert_allclose(out["g"][i], g, rtol=RTOL_SIM)
        assert out["iters"][i] == soln.iterations and out["fg_evals"][i] == soln.f_calls
    # the streaming kernel must have finished every unit itself (nothing handed back to the generic kernel)
    assert be.profile()["redo_units"] == 0
    be.close()


@pytest.mark.parametrize("d,kernel", [(512, 0), (512, 1), (6000, 1), (6000, 2)])
def test_zero_iteration_warm_start_keeps_previous_map(d, kernel):
    """A start point that already satisfies ‖∇z‖_∞ ≤ atol is returned unchanged (0 iterations)."""
    name, nsims = "funnel", 8
    fam, draws, xd = make_inputs(name, d, nsims)
    be = _backend(name, d, nsims, draws, xd, kernel=kernel)
    th = np.array([0.7])
    be.map_score(th, th, 1e-2, include_data=True, warm_start=0)
    z1 = be.get_maps(0, nsims + 1)
    out = be.map_score(th + 1e-7, th + 1e-7, 1e-2, include_data=True, warm_start=1)
    z2 = be.get_maps(0, nsims + 1)
    assert (out["iters"] == 0).all() and (out["fg_evals"] == 1).all()
    np.testing.assert_array_equal(z1, z2)
    # truth start with a huge atol: 0 iterations, the kept start is the simulated latent (src/muse.jl:511)
    prob = O.OracleProblem(fam, xd, draws)
    out = be.map_score(th, th, 1e6, include_data=False, warm_start=2)
    assert (out["iters"] == 0).all()
    z3 = be.get_maps(1, nsims)
    for k in range(nsims):
        np.testing.assert_allclose(z3[k], prob.sample_x_z(k, th)[1], rtol=1e-15)
    be.close()


@pytest.mark.parametrize("name,d,kw", [("funnel", 512, {}), ("hiergauss", 1024, {}),
                                       ("funnel", 6000, dict(group=256, cluster=2, kernel=1)),
                                       ("hiergauss", 5000, {}), ("hiergauss", 40000, dict(kernel=2))])
def test_fd_jacobian_matches_oracle(name, d, kw):
    nsims, nH = 20, 6
    fam, draws, xd = make_inputs(name, d, nsims)
    prob = O.OracleProblem(fam, xd, draws)
    be = _backend(name, d, nsims, draws, xd, **kw)
    th0 = theta_start(name)
    step = np.full(th0.shape, 0.01) * (1 + np.arange(th0.size))
    Hs, status = be.fd_jacobian(th0, step, nH, 1e-2)
    res = O.MuseResult(theta=th0.copy())
    O.get_H_bang(res, prob, th0, nsims=nH, step=step, gradz_logLike_atol=1e-2)
    assert (status == 0).all()
    for k in range(nH):
        np.testing.assert_allclose(Hs[k], res.Hs[k], rtol=1e-6, atol=1e-6 * np.abs(res.Hs[k]).max())
    be.close()


@pytest.mark.parametrize("kw", [{}, dict(group=32), dict(group=32, kernel=1), dict(group=256, cluster=1, kernel=1),
                                dict(group=256, cluster=2, kernel=1), dict(kernel=2), dict(group=512, cluster=1, kernel=2), dict(kernel=3)])
def test_history_path_runs_and_stays_at_the_map(kw):
    """atol far below round-off forces iterations ≥ 2: two-loop recursion over the (dx, dg) history,
    direction resets, x/f stagnation exits.  With kernel 2 the streaming kernel cannot finish such a unit:
    it must hand every unit back and the generic kernel re-solves them from the untouched start.  The iterate
    must stay at the closed-form MAP and the score must agree with the oracle run the same way."""
    name, d, nsims = "funnel", 3000, 12
    fam, draws, xd = make_inputs(name, d, nsims)
    prob = O.OracleProblem(fam, xd, draws)
    be = _backend(name, d, nsims, draws, xd, **kw)
    th = np.array([0.8])
    out = be.map_score(th, th, 1e-300, include_data=True, warm_start=0)
    zs = be.get_maps(0, nsims + 1)
    assert (out["iters"] >= 2).all()
    assert np.isin(out["status"], [0, 1, 3]).all()
    assert be.profile()["redo_units"] == (0 if kw.get("kernel") == 1 else nsims + 1)
    for u in range(nsims + 1):
        x = xd if u == 0 else prob.sample_x_z(u - 1, th)[0]
        np.testing.assert_allclose(zs[u], fam.exact_map(x, th), rtol=1e-12, atol=1e-13)
        _, g, soln = O.map_score_unit(prob, x, np.zeros(d), th, 1e-300)
        assert soln.iterations >= 2
        np.testing.assert_allclose(out["g"][u], g, rtol=1e-9)
    be.close()


@pytest.mark.parametrize("name,d,nsims", [("funnel", 512, 100), ("hiergauss", 2048, 60)])
def test_full_muse_matches_oracle(name, d, nsims):
    import museinference_jl_b200 as m
    prior_o = O.NormalPrior(0, 3) if name == "funnel" else None
    prior_p = m.NormalPrior(0, 3) if name == "funnel" else None
    oprob, fam, draws, xd = oracle_problem(name, d, nsims, prior=prior_o)
    ref = O.muse(oprob, theta_start(name), nsims=nsims, get_covariance=True)
    prob = m.SimpleMuseProblem(xd, name, prior_p)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    res = m.muse(prob, theta_start(name), rng=rng, nsims=nsims, get_covariance=True)
    assert len(res.history) == len(ref.history)
    np.testing.assert_allclose(res.theta, ref.theta, rtol=RTOL_EST)
    np.testing.assert_allclose(res.J, ref.J, rtol=RTOL_EST)
    np.testing.assert_allclose(res.H, ref.H, rtol=RTOL_EST, atol=RTOL_EST * np.abs(ref.H).max())
    np.testing.assert_allclose(res.Sigma, ref.Sigma, rtol=10 * RTOL_EST, atol=RTOL_EST * np.abs(ref.Sigma).max())
    np.testing.assert_allclose(np.array(res.gs), np.array(ref.gs), rtol=RTOL_SIM)
    prob.close()


def test_nccl_shardpool_device_gather_matches_local_pool():
    """World-size-1 NCCL group on this GPU: the scores travel device → all-gather → host instead of through
    map_score's host copy; the solve must not change."""
    import torch
    import torch.distributed as dist
    import museinference_jl_b200 as m
    name, d, nsims = "funnel", 6000, 40
    oprob, fam, draws, xd = oracle_problem(name, d, nsims, prior=O.NormalPrior(0, 3))
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    prob = m.SimpleMuseProblem(xd, name, m.NormalPrior(0, 3))
    ref = m.muse(prob, theta_start(name), rng=rng, nsims=nsims, get_covariance=True)
    prob.close()
    own = not dist.is_initialized()
    if own:
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29533", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
    try:
        pool = m.ShardPool(device=0)
        assert pool.uses_device_gather()
        prob = m.SimpleMuseProblem(xd, name, m.NormalPrior(0, 3))
        res = m.muse(prob, theta_start(name), rng=rng, nsims=nsims, get_covariance=True, pool=pool)
        np.testing.assert_array_equal(res.theta, ref.theta)
        np.testing.assert_array_equal(np.array(res.gs), np.array(ref.gs))
        np.testing.assert_array_equal(res.H, ref.H)
        prob.close()
    finally:
        if own:
            dist.destroy_process_group()


@pytest.mark.parametrize("name,d,nsims,prior", [("funnel", 512, 100, True), ("hiergauss", 5000, 50, False), ("hiergauss", 300, 30, True)])
def test_in_library_outer_loop_matches_line_by_line_driver(name, d, nsims, prior):
    """muse_b200_muse_iterate (csrc/muse_driver.cu) against the Python mirror of src/muse.jl:159-236."""
    import museinference_jl_b200 as m
    oprob, fam, draws, xd = oracle_problem(name, d, nsims)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    pr = (lambda: m.NormalPrior([0.0, 0.1][:fam.ntheta], [3.0, 2.0][:fam.ntheta])) if prior else (lambda: None)
    res = {}
    for fused in ("host", False):
        prob = m.SimpleMuseProblem(xd, name, pr())
        res[fused] = m.muse(prob, theta_start(name), rng=rng, nsims=nsims, get_covariance=True, fused_driver=fused,
                            theta_rtol=1e-3, maxsteps=6)
        prob.close()
    a, b = res["host"], res[False]
    assert len(a.history) == len(b.history) >= 3
    np.testing.assert_allclose(a.theta, b.theta, rtol=1e-12)
    np.testing.assert_allclose(a.J, b.J, rtol=1e-12)
    np.testing.assert_allclose(a.H, b.H, rtol=1e-10)
    np.testing.assert_allclose(np.array(a.gs), np.array(b.gs), rtol=1e-10)
    np.testing.assert_allclose(a.Sigma, b.Sigma, rtol=1e-9)
    np.testing.assert_allclose(np.array(a.Hs), np.array(b.Hs), rtol=1e-8, atol=1e-9)
    for ha, hb in zip(a.history, b.history):
        for key in ("theta", "theta_unreg", "g_like_sims", "g_like_dat", "g_like", "g_prior", "g_post", "H_inv_post",
                    "H_prior", "H_inv_like", "H_inv_like_sims"):
            np.testing.assert_allclose(ha[key], hb[key], rtol=1e-11, atol=1e-300, err_msg=key)
        assert {k: v for k, v in ha["z_history_dat"].items() if k != "gnorm"} == {k: v for k, v in hb["z_history_dat"].items() if k != "gnorm"}
        np.testing.assert_array_equal(ha["z_history_sims"]["fg_evals"], hb["z_history_sims"]["fg_evals"])


def test_dgemm_dmma_matches_numpy():
    """The FP64 tensor-core GEMM of F3 (csrc/muse_dgemm.cu) against NumPy, FP64: rtol 1e-13 on a K = 512 contraction."""
    import ctypes as C
    import museinference_jl_b200 as m
    lib = m.load_library()
    rng = np.random.default_rng(0)
    M, N, K = 256, 384, 512
    A, B = rng.standard_normal((M, K)), rng.standard_normal((K, N))
    Cm = np.empty((M, N))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.muse_b200_dgemm_host(dp(A), dp(B), dp(Cm), M, N, K) == 0
    ref = A @ B
    np.testing.assert_allclose(Cm, ref, rtol=0, atol=1e-13 * np.abs(A).max() * np.abs(B).max() * K)
    assert lib.muse_b200_dgemm_host(dp(A), dp(B), dp(Cm), 100, N, K) != 0      # extents must be tile multiples


def test_trimmed_draws_kernel_is_bit_identical_to_the_first_table_driven_one():
    """philox_draws_tab2_kernel (the default: constants and table addresses held in registers, both streams of a row per iteration,
    funnel-shift / exponent-OR integer → double steps) must produce exactly the normals of philox_draws_tab_kernel: the generator is
    selected once per process (MUSE_DRAWS_IMPL), so each runs in its own interpreter and prints a digest of its output."""
    import subprocess
    import sys
    script = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "draws_ab.py")
    sha = {}
    for impl in ("1", "2", "3"):
        r = subprocess.run([sys.executable, script, "child"], env=dict(os.environ, MUSE_DRAWS_IMPL=impl), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        line = r.stdout.strip().splitlines()[-1]
        sha[impl] = line.split("sha256")[1].strip()
        assert float(line.split("= ")[1].split(" ;")[0]) < 1e-14            # and each against the oracle's generator
    assert sha["1"] == sha["2"] == sha["3"], sha


# ------------------------------------------------------------------------------- F3: dense correlated Gaussian
def _corr_backend(d, nsims, draws, xd, fam, **kw):
    import museinference_jl_b200 as m
    be = m.B200Backend("corrgauss", d, nsims, P=fam.P, L=fam.L, **kw)
    be.set_data(xd)
    be.set_draws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    return be


@pytest.mark.parametrize("d", [200, 384])
def test_corrgauss_map_score_matches_oracle(d):
    """Lock-step L-BFGS with one DMMA GEMM per iteration (csrc/muse_corr.cu) against the oracle's honest L-BFGS
    (P·z evaluated at every trial point): same iteration and evaluation counts, ẑ and g within rtol 1e-8."""
    name, nsims, atol = "corrgauss", 10, 1e-2
    fam, draws, xd = make_inputs(name, d, nsims)
    prob = O.OracleProblem(fam, xd, draws)
    be = _corr_backend(d, nsims, draws, xd, fam)
    th0 = np.array([0.6])
    out = be.map_score(th0, th0, atol, include_data=True, warm_start=0)
    zs = be.get_maps(0, nsims + 1)
    ref_z = []
    for u in range(nsims + 1):
        x = xd if u == 0 else prob.sample_x_z(u - 1, th0)[0]
        zh, g, soln = O.map_score_unit(prob, x, np.zeros(d), th0, atol)
        ref_z.append(zh)
        assert out["iters"][u] == soln.iterations and out["fg_evals"][u] == soln.f_calls, (u, out["iters"][u], soln.iterations)
        np.testing.assert_allclose(zs[u], zh, rtol=RTOL_SIM, atol=1e-11)
        np.testing.assert_allclose(out["g"][u], g, rtol=RTOL_SIM)
        assert out["status"][u] == 0 and out["gnorm"][u] <= atol
    assert out["iters"].min() >= 4          # anisotropic Hessian: a real multi-iteration solve
    # warm pass at a moved θ, then truth start on a sub-range
    th1 = th0 - 0.4
    out = be.map_score(th1, th1, atol, include_data=True, warm_start=1)
    zs = be.get_maps(0, nsims + 1)
    for u in range(nsims + 1):
        x = xd if u == 0 else prob.sample_x_z(u - 1, th1)[0]
        zh, g, soln = O.map_score_unit(prob, x, ref_z[u], th1, atol)
        assert out["iters"][u] == soln.iterations and out["fg_evals"][u] == soln.f_calls
        np.testing.assert_allclose(zs[u], zh, rtol=RTOL_SIM, atol=1e-11)
        np.testing.assert_allclose(out["g"][u], g, rtol=RTOL_SIM)
    out = be.map_score(th1, th1, atol, include_data=False, warm_start=2, first_sim=3, count=5)
    for i in range(5):
        x, z = prob.sample_x_z(3 + i, th1)
        zh, g, soln = O.map_score_unit(prob, x, z, th1, atol)
        assert out["iters"][i] == soln.iterations and out["fg_evals"][i] == soln.f_calls
        np.testing.assert_allclose(out["g"][i], g, rtol=RTOL_SIM)
    be.close()


def test_corrgauss_full_muse_matches_oracle():
    import museinference_jl_b200 as m
    name, d, nsims = "corrgauss", 256, 40
    oprob, fam, draws, xd = oracle_problem(name, d, nsims, prior=O.NormalPrior(0, 3))
    ref = O.muse(oprob, [1.0], nsims=nsims, get_covariance=True)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    for fused in (True, False):
        prob = m.SimpleMuseProblem(xd, name, m.NormalPrior(0, 3), P=fam.P, L=fam.L)
        res = m.muse(prob, [1.0], rng=rng, nsims=nsims, get_covariance=True, fused_driver=fused)
        assert len(res.history) == len(ref.history)
        np.testing.assert_allclose(res.theta, ref.theta, rtol=RTOL_EST)
        np.testing.assert_allclose(res.J, ref.J, rtol=RTOL_EST)
        np.testing.assert_allclose(res.H, ref.H, rtol=RTOL_EST)
        np.testing.assert_allclose(res.Sigma, ref.Sigma, rtol=10 * RTOL_EST)
        np.testing.assert_allclose(np.array(res.gs), np.array(ref.gs), rtol=RTOL_SIM)
        prob.close()


# ------------------------------------------------------------------------------- F4: the two-layer hierarchy of src/turing.jl:63-79
@pytest.mark.parametrize("d,nsims", [(1024, 12), (300, 7), (9000, 5)])
def test_twolayer_map_score_matches_oracle(d, nsims):
    """The docstring toy model of the Turing adapter (latent (z, w), data (x, y), parameter σ) on the lock-step solver with the
    elementwise 2 × 2-block product in place of the DGEMM (csrc/muse_corr.cu) against the oracle's L-BFGS + Hager–Zhang: the
    same iteration and evaluation counts — 2 and 6: a live (dx, dg) history on every unit — ẑ, ‖∇f‖∞ and the score."""
    import museinference_jl_b200 as m
    name, atol = "twolayer", 1e-2
    fam, draws, xd = make_inputs(name, d, nsims)
    prob = O.OracleProblem(fam, xd, draws)
    be = m.B200Backend(name, d, nsims)
    be.set_data(xd)
    be.set_draws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    th0 = np.array([0.5])
    out = be.map_score(th0, th0, atol, include_data=True, warm_start=0)
    zs = be.get_maps(0, nsims + 1)
    ref_z = []
    for u in range(nsims + 1):
        x = xd if u == 0 else prob.sample_x_z(u - 1, th0)[0]
        zh, g, soln = O.map_score_unit(prob, x, np.zeros(d), th0, atol)
        ref_z.append(zh)
        assert out["iters"][u] == soln.iterations == 2 and out["fg_evals"][u] == soln.f_calls == 6, (u, out["iters"][u], out["fg_evals"][u])
        np.testing.assert_allclose(zs[u], zh, rtol=RTOL_SIM, atol=1e-11)
        np.testing.assert_allclose(zs[u], fam.exact_map(x, th0), rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(out["g"][u], g, rtol=RTOL_SIM)
        assert out["status"][u] == 0 and out["gnorm"][u] <= atol
    # warm pass at a moved σ (start = the unit's own ẑ), then the simulated latent as the start on a sub-range (get_J!)
    th1 = th0 - 0.35
    out = be.map_score(th1, th1, atol, include_data=True, warm_start=1)
    zs = be.get_maps(0, nsims + 1)
    for u in range(nsims + 1):
        x = xd if u == 0 else prob.sample_x_z(u - 1, th1)[0]
        zh, g, soln = O.map_score_unit(prob, x, ref_z[u], th1, atol)
        assert out["iters"][u] == soln.iterations and out["fg_evals"][u] == soln.f_calls
        np.testing.assert_allclose(zs[u], zh, rtol=RTOL_SIM, atol=1e-11)
        np.testing.assert_allclose(out["g"][u], g, rtol=RTOL_SIM)
    out = be.map_score(th1, th1, atol, include_data=False, warm_start=2, first_sim=2, count=3)
    for i in range(3):
        x, z = prob.sample_x_z(2 + i, th1)
        zh, g, soln = O.map_score_unit(prob, x, z, th1, atol)
        assert out["iters"][i] == soln.iterations and out["fg_evals"][i] == soln.f_calls
        np.testing.assert_allclose(out["g"][i], g, rtol=RTOL_SIM)
    be.close()
    with pytest.raises(m.MuseBackendError):
        m.B200Backend(name, 11, 4)                          # d = 2n must be even


def test_twolayer_full_muse_matches_oracle_and_the_golden_fixture():
    """muse(prob, σ₀ = 0.5; get_covariance = true) of the docstring model at its own size (n = 512 per layer): in-library host loop
    and line-by-line Python loop against the oracle and the committed fixture; device-generated draws: partition invariance."""
    import json
    import os
    import museinference_jl_b200 as m
    with open(os.path.join(os.path.dirname(__file__), "golden", "twolayer_d1024_n60.json")) as fh:
        fix = json.load(fh)
    c = fix["case"]
    name, d, nsims = c["family"], c["d"], c["nsims"]
    oprob, fam, draws, xd = oracle_problem(name, d, nsims, seed=c["seed"], prior=O.NormalPrior(0, 3))
    ref = O.muse(oprob, fix["theta0"], nsims=nsims, get_covariance=True)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    for fused in (True, False, "device"):                  # "device": this family keeps the host loop, like corrgauss
        prob = m.SimpleMuseProblem(xd, name, m.NormalPrior(0, 3))
        res = m.muse(prob, fix["theta0"], rng=rng, nsims=nsims, get_covariance=True, fused_driver=fused)
        assert len(res.history) == len(ref.history) == fix["n_outer"]
        for want in (ref.theta, fix["theta"]):
            np.testing.assert_allclose(res.theta, want, rtol=RTOL_EST)
        np.testing.assert_allclose(res.J, fix["J"], rtol=RTOL_EST)
        np.testing.assert_allclose(res.H, fix["H"], rtol=RTOL_EST)
        np.testing.assert_allclose(res.Sigma, fix["Sigma"], rtol=10 * RTOL_EST)
        np.testing.assert_allclose(np.array(res.gs), np.array(ref.gs), rtol=RTOL_SIM)
        assert (res.history[0]["z_history_sims"]["iters"] == fix["iter1"]["iters"]).all()
        assert (res.history[0]["z_history_sims"]["fg_evals"] == fix["iter1"]["fg"]).all()
        prob.close()
    # get_J! / get_H! on their own
    prob = m.SimpleMuseProblem(xd, name, m.NormalPrior(0, 3))
    r2, o2 = m.MuseResult(theta=np.array(fix["theta"])), O.MuseResult(theta=np.array(fix["theta"]))
    getattr(m, "get_J!")(r2, prob, rng=rng, nsims=20)
    O.get_J_bang(o2, oprob, nsims=20)
    np.testing.assert_allclose(r2.J, o2.J, rtol=RTOL_EST)
    getattr(m, "get_H!")(r2, prob, rng=rng, nsims=6)
    O.get_H_bang(o2, oprob, nsims=6)
    np.testing.assert_allclose(r2.H, o2.H, rtol=RTOL_EST)
    prob.close()
    # device Philox draws: a handle owning sims [20, 35) reproduces those rows of the whole solve bit for bit
    full = m.B200Backend(name, d, 50); full.set_data(xd); full.seed_draws(4242)
    part = m.B200Backend(name, d, 15, sim_offset=20); part.set_data(xd); part.seed_draws(4242)
    th = np.array([0.4])
    a = full.map_score(th, th, 1e-2, include_data=False, warm_start=0)
    b = part.map_score(th, th, 1e-2, include_data=False, warm_start=0)
    np.testing.assert_array_equal(a["g"][20:35], b["g"])
    np.testing.assert_array_equal(a["iters"][20:35], b["iters"])
    full.close(); part.close()


def test_full_size_c3_closed_form_and_partition_invariance():
    """BASELINE configs[2] at full size (funnel, d = 65 536, nsims = 2 048, device Philox draws): every per-sim score
    against the closed form g = ½ e^{-θ} s² ‖x‖² − d/2 with x = e^{θ/2} ξ + ν, iteration counts,
    and — the size-independent property that makes sharding safe — a handle owning only sims [1000, 1100) of the same
    seed reproduces those rows bit for bit."""
    import math
    import museinference_jl_b200 as m
    d, n, seed = 65536, 2048, 20261017
    be = m.B200Backend("funnel", d, n)
    be.set_data(np.zeros(d))
    be.seed_draws(seed)
    th = np.array([0.35])
    out = be.map_score(th, th, 1e-2, include_data=False, warm_start=0)
    assert (out["iters"] == 1).all() and (out["fg_evals"] == 3).all() and (out["status"] == 0).all()
    assert be.profile()["redo_units"] == 0
    s = 1.0 / (1.0 + math.exp(-th[0]))
    for lo in range(0, n, 256):                       # closed form from the draws themselves, 256 rows at a time
        xi, nu = be.get_draws(lo, 256)
        x = math.exp(0.5 * th[0]) * xi + nu
        g_ref = 0.5 * math.exp(-th[0]) * s * s * np.einsum("ij,ij->i", x, x) - d / 2
        np.testing.assert_allclose(out["g"][lo:lo + 256, 0], g_ref, rtol=1e-9)
    z = be.get_maps(1 + 1234, 1)[0]
    xi, nu = be.get_draws(1234, 1)
    np.testing.assert_allclose(z, s * (math.exp(0.5 * th[0]) * xi[0] + nu[0]), rtol=1e-12, atol=1e-15)
    # J of these scores against the closed form d s²/2 (Monte-Carlo error √(2/N))
    assert np.var(out["g"][:, 0], ddof=1) == pytest.approx(d * s * s / 2, rel=5 * math.sqrt(2.0 / n))
    be.close()
    part = m.B200Backend("funnel", d, 100, sim_offset=1000)
    part.set_data(np.zeros(d))
    part.seed_draws(seed)
    o2 = part.map_score(th, th, 1e-2, include_data=False, warm_start=0)
    np.testing.assert_array_equal(o2["g"][:, 0], out["g"][1000:1100, 0])
    part.close()


# ------------------------------------------------------------------------------- edge cases and error behaviour
def test_call_order_and_argument_errors():
    import museinference_jl_b200 as m
    be = m.B200Backend("funnel", 100, 4)
    th = np.array([0.1])
    with pytest.raises(m.MuseBackendError) as ei:          # no data yet
        be.map_score(th, th, 1e-2, include_data=True, warm_start=0)
    assert ei.value.code == -6
    be.set_data(np.zeros(100))
    with pytest.raises(m.MuseBackendError) as ei:          # no draws yet
        be.map_score(th, th, 1e-2, include_data=True, warm_start=0)
    assert ei.value.code == -6
    be.seed_draws(1)
    with pytest.raises(m.MuseBackendError) as ei:          # user start without z₀
        be.map_score(th, th, 1e-2, include_data=True, warm_start=3)
    assert ei.value.code == -6
    with pytest.raises(m.MuseBackendError) as ei:          # range outside the shard
        be.map_score(th, th, 1e-2, include_data=False, warm_start=0, first_sim=3, count=2)
    assert ei.value.code == -1
    with pytest.raises(m.MuseBackendError):                # zero FD step
        be.fd_jacobian(th, np.array([0.0]), 2, 1e-2)
    out = be.map_score(th, th, 1e-2, include_data=False, warm_start=0, first_sim=2, count=0)   # empty range is fine
    assert out["g"].shape == (0, 1)
    be.close()
    with pytest.raises(m.MuseBackendError) as ei:          # ntheta / family mismatch is caught at create
        m.B200Backend("corrgauss", 64, 4)                  # no P, L
    assert ei.value.code == -1


@pytest.mark.parametrize("d,kernel", [(1, 0), (3, 0), (33, 1), (4100, 2)])
def test_tiny_and_ragged_dimensions(d, kernel):
    name, nsims = "hiergauss", 5
    fam, draws, xd = make_inputs(name, d, nsims)
    prob = O.OracleProblem(fam, xd, draws)
    be = _backend(name, d, nsims, draws, xd, kernel=kernel)
    th = theta_start(name)
    out = be.map_score(th, th, 1e-2, include_data=True, warm_start=0)
    zs = be.get_maps(0, nsims + 1)
    for u in range(nsims + 1):
        x = xd if u == 0 else prob.sample_x_z(u - 1, th)[0]
        zh, g, soln = O.map_score_unit(prob, x, np.zeros(d), th, 1e-2)
        np.testing.assert_allclose(out["g"][u], g, rtol=RTOL_SIM, atol=1e-12)
        np.testing.assert_allclose(zs[u], zh, rtol=RTOL_SIM, atol=1e-12)
        assert out["iters"][u] == soln.iterations
    be.close()


@pytest.mark.parametrize("d", [512, 6000])
def test_user_start_vector_and_save_maps(d):
    import museinference_jl_b200 as m
    name, nsims = "funnel", 12
    oprob, fam, draws, xd = oracle_problem(name, d, nsims, prior=O.NormalPrior(0, 3))
    z0 = np.linspace(-1, 1, d)
    ref = O.muse(oprob, [1.0], nsims=nsims, z0=z0, save_MAPs=True, maxsteps=3)
    prob = m.SimpleMuseProblem(xd, name, m.NormalPrior(0, 3))
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    res = m.muse(prob, [1.0], rng=rng, nsims=nsims, z0=z0, save_MAPs=True, maxsteps=3)
    assert len(res.history) == len(ref.history)
    np.testing.assert_allclose(res.theta, ref.theta, rtol=RTOL_EST)
    for a, b in zip(res.history, ref.history):
        np.testing.assert_allclose(a["z_dat"], b["z_dat"], rtol=RTOL_SIM, atol=1e-12)
        np.testing.assert_allclose(a["z_sims"], np.array(b["z_sims"]), rtol=RTOL_SIM, atol=1e-12)
    prob.close()


def test_non_finite_data_is_reported_not_hidden():
    """A NaN in the data gives a non-finite objective: status NONFINITE for that unit; muse! raises (src/interface.jl:170),
    get_J!(skip_errors=true) would drop such sims (src/muse.jl:515-521)."""
    import museinference_jl_b200 as m
    for d in (512, 6000):
        fam, draws, xd = make_inputs("funnel", d, 6)
        xbad = xd.copy()
        xbad[d // 2] = np.nan
        be = _backend("funnel", d, 6, draws, xbad)
        out = be.map_score(np.array([0.3]), np.array([0.3]), 1e-2, include_data=True, warm_start=0)
        assert out["status"][0] == 4 and (out["status"][1:] == 0).all()
        be.close()
        prob = m.SimpleMuseProblem(xbad, "funnel", m.NormalPrior(0, 3))
        rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
        with pytest.raises((FloatingPointError, m.MuseBackendError)):
            m.muse(prob, [1.0], rng=rng, nsims=6)
        prob.close()


def test_get_J_top_up_and_resume_on_gpu():
    import museinference_jl_b200 as m
    name, d, nsims = "hiergauss", 5000, 30
    oprob, fam, draws, xd = oracle_problem(name, d, nsims)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    prob = m.SimpleMuseProblem(xd, name)
    th = np.array([0.2, 0.1])
    res, ref = m.MuseResult(theta=th.copy()), O.MuseResult(theta=th.copy())
    getJ, getH = getattr(m, "get_J!"), getattr(m, "get_H!")
    getJ(res, prob, rng=rng, nsims=10)
    getJ(res, prob, rng=rng, nsims=30)                   # top-up: sims 10..29 only, truth start (src/muse.jl:499-511)
    O.get_J_bang(ref, oprob, nsims=30)
    np.testing.assert_allclose(res.J, ref.J, rtol=RTOL_EST)
    np.testing.assert_allclose(np.array(res.gs), np.array(ref.gs), rtol=RTOL_SIM)
    getH(res, prob, rng=rng, nsims=5)
    O.get_H_bang(ref, oprob, nsims=5)
    np.testing.assert_allclose(res.H, ref.H, rtol=RTOL_EST, atol=RTOL_EST * np.abs(ref.H).max())
    np.testing.assert_allclose(res.Sigma, ref.Sigma, rtol=10 * RTOL_EST, atol=1e-12)
    prob.close()


def test_corrgauss_iteration_cap_and_status():
    name, d, nsims = "corrgauss", 128, 4
    fam, draws, xd = make_inputs(name, d, nsims)
    import museinference_jl_b200 as m
    be = m.B200Backend(name, d, nsims, P=fam.P, L=fam.L, max_iters=2)
    be.set_data(xd)
    be.set_draws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    out = be.map_score(np.array([-1.0]), np.array([-1.0]), 1e-2, include_data=True, warm_start=0)
    assert (out["iters"] == 2).all() and (out["status"] == 2).all()      # MAXITER: Optim.converged == false
    be.close()


def test_full_size_c4_hiergauss_closed_form():
    """BASELINE configs[3] at full size (hierarchical Gaussian, d = 10⁵, nsims = 4 096, θ = (μ, ℓ)): iteration counts of every
    unit, and both score components of sampled sims against the closed form at the exact MAP
    ẑ = (x + μ a)/(1 + a), a = e^{−2ℓ}:  g = (a Σ(ẑ−μ), a Σ(ẑ−μ)² − d)."""
    import math
    import museinference_jl_b200 as m
    d, n, seed = 100000, 4096, 7
    be = m.B200Backend("hiergauss", d, n)
    be.set_data(np.zeros(d))
    be.seed_draws(seed)
    th = np.array([0.5, 0.3])
    out = be.map_score(th, th, 1e-2, include_data=False, warm_start=0)
    assert (out["iters"] == 1).all() and (out["fg_evals"] == 3).all() and (out["status"] == 0).all()
    assert be.profile()["redo_units"] == 0
    a = math.exp(-2 * th[1])
    for lo in (0, 1777, 4096 - 64):
        xi, nu = be.get_draws(lo, 64)
        x = th[0] + math.exp(th[1]) * xi + nu
        w = (x + th[0] * a) / (1 + a) - th[0]
        np.testing.assert_allclose(out["g"][lo:lo + 64, 0], a * w.sum(1), rtol=1e-9, atol=1e-7)
        np.testing.assert_allclose(out["g"][lo:lo + 64, 1], a * (w * w).sum(1) - d, rtol=1e-9)
    # warm pass at a moved θ: again one iteration per unit (isotropic Hessian), scores move as the closed form says
    th2 = np.array([0.45, 0.28])
    out2 = be.map_score(th2, th2, 1e-2, include_data=False, warm_start=1)
    assert (out2["iters"] == 1).all() and be.profile()["redo_units"] == 0
    a2 = math.exp(-2 * th2[1])
    xi, nu = be.get_draws(100, 8)
    x = th2[0] + math.exp(th2[1]) * xi + nu
    w = (x + th2[0] * a2) / (1 + a2) - th2[0]
    np.testing.assert_allclose(out2["g"][100:108, 1], a2 * (w * w).sum(1) - d, rtol=1e-9)
    be.close()


def test_full_size_c5_corrgauss_properties():
    """BASELINE configs[4] at full size (dense correlated Gaussian, d = 4 096, nsims = 8 192): every unit stops with
    ‖∇z‖∞ ≤ atol in a narrow band of iteration counts; for sampled units the returned ẑ satisfies the stationarity
    residual computed on the host with the full P, and the score equals ½ a ẑᵀPẑ − d/2."""
    import math
    import museinference_jl_b200 as m
    from bench import corr_consts
    d, n, atol = 4096, 8192, 1e-2
    P, L = corr_consts(d)
    be = m.B200Backend("corrgauss", d, n, P=P, L=L)
    rng = np.random.default_rng(3)
    xd = L @ rng.standard_normal(d) + rng.standard_normal(d)
    be.set_data(xd)
    be.seed_draws(11)
    th = np.array([0.8])
    out = be.map_score(th, th, atol, include_data=True, warm_start=0)
    assert (out["status"] == 0).all() and (out["gnorm"] <= atol).all()
    assert out["iters"].min() >= 4 and out["iters"].max() - out["iters"].min() <= 3
    a = math.exp(-th[0])
    z0 = be.get_maps(0, 1)[0]                                  # the data unit: x is known on the host
    resid = (z0 - xd) + a * (P @ z0)
    assert np.abs(resid).max() <= atol * (1 + 1e-6)
    assert np.abs(resid).max() == pytest.approx(out["gnorm"][0], rel=1e-6)
    assert out["g"][0, 0] == pytest.approx(0.5 * a * z0 @ (P @ z0) - d / 2, rel=1e-10)
    zs = be.get_maps(4000, 3)                                  # sims: score from the returned ẑ
    for k in range(3):
        assert out["g"][4000 + k, 0] == pytest.approx(0.5 * a * zs[k] @ (P @ zs[k]) - d / 2, rel=1e-10)
    be.close()


@pytest.mark.parametrize("name,d,kernel", [("funnel", 4500, 2), ("hiergauss", 4500, 2), ("funnel", 700, 3), ("hiergauss", 701, 3)])
def test_single_pass_kernels_agree_with_generic_solver_over_random_regimes(name, d, kernel):
    """The straight-line replay of the optimiser's decisions (fast_replay) against the generic kernel, which runs the
    full Controller: random θ (a = e^{−θ} from 3·10⁻⁴ to 3·10³), tolerances from 10⁻¹ down to 10⁻¹⁸ (below round-off), cold /
    warm / truth starts.  Whatever path a unit takes (accepted by the single pass or handed back), scores and ẑ must agree to
    1e-9 and — for tolerances above round-off — iteration and evaluation counts and statuses must be identical."""
    nsims = 6
    fam, draws, xd = make_inputs(name, d, nsims)
    rng = np.random.default_rng(d + kernel)
    a = _backend(name, d, nsims, draws, xd, kernel=kernel)
    b = _backend(name, d, nsims, draws, xd, kernel=1)
    handed_back = 0
    for trial in range(24):
        th = rng.uniform(-8, 8, size=fam.ntheta) if name == "funnel" else np.array([rng.uniform(-2, 2), rng.uniform(-4, 4)])
        atol = 10.0 ** rng.uniform(-18, -1)
        ws = int(rng.integers(0, 3)) if trial else 0
        kw = dict(include_data=(ws != 2), warm_start=ws)
        oa, ob = a.map_score(th, th, atol, **kw), b.map_score(th, th, atol, **kw)
        if atol >= 1e-12:       # below round-off the counts depend on the last bits of the (warm) start: only values are compared
            np.testing.assert_array_equal(oa["iters"], ob["iters"], err_msg=f"trial {trial} θ={th} atol={atol} ws={ws}")
            np.testing.assert_array_equal(oa["fg_evals"], ob["fg_evals"])
            np.testing.assert_array_equal(oa["status"], ob["status"])
        np.testing.assert_allclose(oa["g"], ob["g"], rtol=1e-9, atol=1e-9 * d)
        za, zb = a.get_maps(0, nsims + 1), b.get_maps(0, nsims + 1)
        np.testing.assert_allclose(za, zb, rtol=1e-9, atol=1e-12)
        handed_back = a.profile()["redo_units"]
    assert handed_back > 0          # some regimes (tolerances below round-off) must have gone through the hand-back
    assert b.profile()["redo_units"] == 0
    a.close()
    b.close()


def test_corrgauss_unit_results_do_not_depend_on_the_shard():
    """A handle owning sims [200, 400) reproduces, bit for bit, those units of a 400-sim handle with the same seed
    (Philox keyed by the global sim index; per-row GEMM and reduction orders do not depend on the batch)."""
    import museinference_jl_b200 as m
    name, d, n = "corrgauss", 256, 400
    fam = make_inputs(name, d, 1)[0]
    x = np.random.default_rng(0).standard_normal(d)
    th = np.array([1.0])
    outs, zs = [], []
    for kw, rows in ((dict(nsims=n), (201, 200)), (dict(nsims=200, sim_offset=200), (1, 200))):
        be = m.B200Backend(name, d, kw.pop("nsims"), P=fam.P, L=fam.L, **kw)
        be.set_data(x)
        be.seed_draws(7)
        o = be.map_score(th, th, 1e-2, include_data=True, warm_start=0)
        o2 = be.map_score(th - 0.2, th - 0.2, 1e-2, include_data=True, warm_start=1)
        outs.append((o["g"][0], o["g"][rows[0]:rows[0] + rows[1]], o2["g"][rows[0]:rows[0] + rows[1]], o2["iters"][rows[0]:rows[0] + rows[1]]))
        zs.append(be.get_maps(*rows))
        be.close()
    for a, b in zip(outs[0], outs[1]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(zs[0], zs[1])


# ----------------------------------------------------------------------------- fd_scores, θ-transforms, per-pass profile
@pytest.mark.parametrize("name,d,kw", [("funnel", 512, {}), ("hiergauss", 5000, {}), ("hiergauss", 700, dict(kernel=1)),
                                       ("corrgauss", 256, {})])
def test_fd_scores_reproduces_fd_jacobian_and_takes_asymmetric_points(name, d, kw):
    """muse_b200_fd_scores is the launch sequence of fd_jacobian with the sample points supplied by the host: at the
    symmetric points θ₀ ∓ h e_n it must reproduce fd_jacobian's columns to the last bit or two, and at arbitrary points each
    virtual sim must match the oracle's (sample at θ_sim, MAP + score at θ_eval from the fiducial ẑ)."""
    nsims, nH = 12, 5
    fam, draws, xd = make_inputs(name, d, nsims)
    prob = O.OracleProblem(fam, xd, draws)
    if name == "corrgauss":
        kw = dict(kw, P=fam.P, L=fam.L)
    be = _backend(name, d, nsims, draws, xd, **kw)
    th0 = theta_start(name)
    nt = th0.size
    step = np.full(nt, 0.02) * (1 + np.arange(nt))
    Hs, _ = be.fd_jacobian(th0, step, nH, 1e-2)
    pts = np.empty((2 * nt, nt))
    for n in range(nt):
        for s in (0, 1):
            pts[2 * n + s] = th0
            pts[2 * n + s, n] = th0[n] + (0.0 + step[n] * (1.0 if s else -1.0))
    g, status = be.fd_scores(th0, pts, nH, 1e-2)
    assert g.shape == (nH, 2 * nt, nt) and (status == 0).all()
    for n in range(nt):
        col = ((g[:, 2 * n, :] * -0.5 + 0.0) + g[:, 2 * n + 1, :] * 0.5) / step[n]
        np.testing.assert_allclose(col, Hs[:, :, n], rtol=0, atol=4e-16 * np.abs(g).max() / step[n])   # same launch, same combine
    # asymmetric points (what a bounded θ produces): every virtual sim against the oracle
    pts2 = pts + 0.013 * np.arange(1, 2 * nt + 1)[:, None]
    g2, status2 = be.fd_scores(th0, pts2, nH, 1e-2)
    assert (status2 == 0).all()
    xm, _ = prob.sample_x_z("master", th0)
    zfid, _ = prob.z_at_theta(xm, np.zeros(d), th0, 1e-2)
    for k in range(nH):
        for p in range(2 * nt):
            x, _ = prob.sample_x_z(k, pts2[p])
            zh, _ = prob.z_at_theta(x, zfid, th0, 1e-2)
            np.testing.assert_allclose(g2[k, p], prob.grad_theta(x, zh, th0), rtol=RTOL_SIM, atol=1e-7)
    with pytest.raises(Exception):
        be.fd_scores(th0, np.full((2 * nt, nt), np.nan), nH, 1e-2)
    be.close()


@pytest.mark.parametrize("d,nsims,prior", [(2048, 60, False), (6000, 40, True)])
def test_transformed_theta_full_muse_matches_oracle(d, nsims, prior):
    """θ = (μ, σ) with σ > 0 (transform_θ = (μ, log σ), src/interface.jl:14-28): the outer iteration runs in θ′, scores,
    J and H are reported in θ (src/muse.jl:172-173, 432, 513, 225-227)."""
    import museinference_jl_b200 as m
    fam = O.TransformedFamily(O.HierGauss(d), ("identity", "log"))
    draws = O.Draws.from_philox(4321, nsims, d)
    xd, _ = fam.sample([0.0, 1.0], O.philox_normals(99, 0, 0, d), O.philox_normals(99, 0, 1, d))
    oprob = O.OracleProblem(fam, xd, draws, O.NormalPrior([0.0, 1.0], [2.0, 0.7]) if prior else None)
    th0 = np.array([0.5, np.exp(0.3)])
    ref = O.muse(oprob, th0, nsims=nsims, get_covariance=True)
    prob = m.SimpleMuseProblem(xd, "hiergauss", m.NormalPrior([0.0, 1.0], [2.0, 0.7]) if prior else None,
                               theta_transform=("identity", "log"))
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    res = m.muse(prob, th0, rng=rng, nsims=nsims, get_covariance=True)
    assert len(res.history) == len(ref.history)
    np.testing.assert_allclose(res.theta, ref.theta, rtol=RTOL_EST)
    np.testing.assert_allclose(np.array(res.gs), np.array(ref.gs), rtol=RTOL_SIM, atol=1e-7)
    np.testing.assert_allclose(res.J, ref.J, rtol=RTOL_EST)
    np.testing.assert_allclose(res.H, ref.H, rtol=RTOL_EST, atol=RTOL_EST * np.abs(ref.H).max())
    np.testing.assert_allclose(res.Sigma, ref.Sigma, rtol=10 * RTOL_EST, atol=RTOL_EST * np.abs(ref.Sigma).max())
    for a, b in zip(res.history, ref.history):
        np.testing.assert_allclose(a["theta_t"], b["theta_t"], rtol=RTOL_EST)
        np.testing.assert_allclose(a["g_like_sims_t"], b["g_like_sims_t"], rtol=RTOL_SIM, atol=1e-7)
    prob.close()


def test_profile_splits_solver_time_by_pass_kind():
    """cold pass, warm pass, get_J! (truth start) and get_H! (fiducial + FD) are reported separately."""
    name, d, nsims = "funnel", 8192, 64
    fam, draws, xd = make_inputs(name, d, nsims)
    be = _backend(name, d, nsims, draws, xd)
    th = np.array([0.7])
    be.profile_reset(True)
    be.map_score(th, th, 1e-2, include_data=True, warm_start=0)
    be.map_score(th, th, 1e-2, include_data=True, warm_start=1)
    be.map_score(th, th, 1e-2, include_data=True, warm_start=1)
    be.map_score(th, th, 1e-2, include_data=False, warm_start=2)
    be.fd_jacobian(th, np.array([0.01]), 6, 1e-2)
    p = be.profile_passes()
    tot = be.profile()
    assert [p[k]["launches"] for k in ("cold", "warm", "truth", "fiducial", "fd")] == [1, 2, 1, 1, 1]
    assert p["cold"]["units"] == nsims + 1 and p["warm"]["units"] == 2 * (nsims + 1) and p["truth"]["units"] == nsims
    assert p["fiducial"]["units"] == 1 and p["fd"]["units"] == 12
    assert p["cold"]["bytes"] == nsims * 24 * d + 16 * d and p["warm"]["bytes"] == 2 * (nsims * 32 * d + 24 * d)
    assert abs(sum(v["ms"] for v in p.values()) - tot["solve_ms"]) < 1e-6 * max(1.0, tot["solve_ms"])
    assert all(v["ms"] > 0 for v in p.values())
    be.profile_reset(False)
    be.close()


@pytest.mark.parametrize("name,d", [("funnel", 512), ("hiergauss", 5000)])
def test_get_H_keywords_user_start_and_five_point_fdm(name, d):
    """get_H!(z₀ = …) — the fiducial solve starts from the user's vector (src/muse.jl:309, 419; muse_b200_fd_start) — and
    get_H!(fdm = central_fdm(5, 1)) (src/muse.jl:300) on the GPU against the oracle's restatement."""
    import museinference_jl_b200 as m
    nsims = 20
    oprob, fam, draws, xd = oracle_problem(name, d, nsims)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    prob = m.SimpleMuseProblem(xd, name)
    getH = getattr(m, "get_H!")
    th = theta_start(name)
    step = np.full(fam.ntheta, 0.02)
    z0 = np.sin(np.arange(d) * 0.01)
    for kw_m, kw_o in ((dict(z0=z0), dict(z0=z0)), (dict(fdm=m.central_fdm(5, 1)), dict(fdm=O.central_fdm(5, 1))),
                       (dict(z0=z0, fdm=m.central_fdm(5, 1), gradz_logLike_atol=1e-9), dict(z0=z0, fdm=O.central_fdm(5, 1), gradz_logLike_atol=1e-9))):
        res, ref = m.MuseResult(theta=th.copy()), O.MuseResult(theta=th.copy())
        getH(res, prob, rng=rng, nsims=4, step=step, **kw_m)
        O.get_H_bang(ref, oprob, nsims=4, step=step, **kw_o)
        scale = np.abs(np.array(ref.Hs)).max()
        np.testing.assert_allclose(np.array(res.Hs), np.array(ref.Hs), rtol=RTOL_EST, atol=RTOL_EST * scale)
        np.testing.assert_allclose(res.H, ref.H, rtol=RTOL_EST, atol=RTOL_EST * scale)
    # the option does not outlive the call: the next plain get_H! starts from zero(z) again
    res, ref = m.MuseResult(theta=th.copy()), O.MuseResult(theta=th.copy())
    getH(res, prob, rng=rng, nsims=4, step=step)
    O.get_H_bang(ref, oprob, nsims=4, step=step)
    np.testing.assert_allclose(res.H, ref.H, rtol=RTOL_EST, atol=RTOL_EST * np.abs(ref.H).max())
    prob.close()


@pytest.mark.parametrize("name,d", [("funnel", 512), ("hiergauss", 4500), ("twolayer", 400)])
def test_get_H_with_finite_differences_own_step(name, d):
    """get_H! with neither `step` nor scores in the result (src/muse.jl:411-413 → src/util.jl:13: fdm(f, 0.0)): FiniteDifferences'
    adaptive step per sim and per θ component, from batched muse_b200_fd_scores calls, against the oracle's restatement.  The
    steps rest on a five-point estimate of a third derivative (round-off amplified by 1/h³): they agree to a few 1e-3 where that
    derivative does not vanish, the Jacobians to rtol 1e-6."""
    import museinference_jl_b200 as m
    oprob, fam, draws, xd = oracle_problem(name, d, 12)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    prob = m.SimpleMuseProblem(xd, name)
    th = theta_start(name)
    res, ref = m.MuseResult(theta=th.copy()), O.MuseResult(theta=th.copy())
    getattr(m, "get_H!")(res, prob, rng=rng, nsims=4, gradz_logLike_atol=1e-6)
    O.get_H_bang(ref, oprob, nsims=4, gradz_logLike_atol=1e-6)
    steps, steps_ref = res.metadata["fd_adaptive_steps"], np.array(ref.metadata["fd_adaptive_steps"])
    assert steps.shape == steps_ref.shape == (4, fam.ntheta)
    if name != "hiergauss":
        np.testing.assert_allclose(steps, steps_ref, rtol=5e-2)
    # (hiergauss: the scores are linear / quadratic in the simulation's μ and polynomial-times-exponential in ℓ — where the third
    # derivative vanishes identically the step is set by round-off alone and differs between any two implementations; the central
    # difference of such a function is exact at any step)
    assert (steps > 0).all() and (steps <= 1000 * m.AdaptedFDM(3, 1).default_step()).all()
    scale = np.abs(np.array(ref.Hs)).max()
    np.testing.assert_allclose(np.array(res.Hs), np.array(ref.Hs), rtol=RTOL_EST, atol=RTOL_EST * scale)
    prob.close()


@pytest.mark.parametrize("name,d,nsims", [("funnel", 512, 30), ("hiergauss", 5000, 20), ("corrgauss", 256, 24), ("twolayer", 1024, 16)])
def test_implicit_diff_get_H_matches_oracle(name, d, nsims):
    """get_H!(implicit_diff = true) (src/muse.jl:335-405) on the GPU — MAP pass at ∇z_logLike_atol = 1e-1, closed-form second
    derivatives, conjugate gradients (one exact iteration for the isotropic families, a batched CG on the DMMA GEMM for corrgauss,
    the same batched CG on the elementwise block product for twolayer: two iterations, one per Hessian eigenvalue)
    — against the oracle's restatement: per-sim H within rtol 1e-6, identical CG iteration counts; and the implicit-diff H agrees
    with the finite-difference H of the same sims."""
    import museinference_jl_b200 as m
    from helpers import make_family
    oprob, fam, draws, xd = oracle_problem(name, d, nsims)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    kw = dict(P=fam.P, L=fam.L) if name == "corrgauss" else {}
    prob = m.SimpleMuseProblem(xd, name, **kw)
    getH = getattr(m, "get_H!")
    th = theta_start(name)
    nh = 6
    res, ref = m.MuseResult(theta=th.copy()), O.MuseResult(theta=th.copy())
    getH(res, prob, rng=rng, nsims=nh, implicit_diff=True)
    O.get_H_bang(ref, oprob, nsims=nh, implicit_diff=True)
    scale = np.abs(np.array(ref.Hs)).max()
    np.testing.assert_allclose(np.array(res.Hs), np.array(ref.Hs), rtol=RTOL_EST, atol=RTOL_EST * scale)
    np.testing.assert_allclose(res.H, ref.H, rtol=RTOL_EST, atol=RTOL_EST * scale)
    its, its_ref = np.array(res.metadata["implicit_diff_cg_hists"]), np.array(ref.metadata["implicit_diff_cg_hists"])
    if name == "corrgauss":
        assert its.min() >= 3 and np.abs(its - its_ref).max() <= 1          # stopping test at √eps·‖b‖: a borderline residual may cost one iteration
    else:
        np.testing.assert_array_equal(its, its_ref)
        assert (its == (2 if name == "twolayer" else 1)).all()
    fd = m.MuseResult(theta=th.copy())
    getH(fd, prob, rng=rng, nsims=nh, step=np.full(fam.ntheta, 1e-3), gradz_logLike_atol=1e-9)
    np.testing.assert_allclose(res.H, fd.H, rtol=5e-3 if name == "corrgauss" else 1e-4, atol=1e-4 * np.abs(fd.H).max())
    prob.close()


# ----------------------------------------------------------------------------- the whole solve in one launch (solve_persist_kernel)
def _solve_persist(m, xd, name, pr, rng, nsims, persist, lazy=False, lean=False, **kw):
    """One solve with MUSE_PERSIST = persist (and MUSE_LAZY = lazy, MUSE_LEAN = lean); returns (result, profile, per-pass profile)."""
    import os
    os.environ["MUSE_PERSIST"] = "1" if persist else "0"
    os.environ["MUSE_LAZY"] = "1" if lazy else "0"
    os.environ["MUSE_LEAN"] = "1" if lean else "0"
    try:
        prob = m.SimpleMuseProblem(xd, name, pr())
        res = m.muse(prob, theta_start(name), rng=rng, nsims=nsims, **kw)      # first solve: allocations
        be = prob._backend
        be.profile_reset(True)
        res = m.muse(prob, theta_start(name), rng=rng, nsims=nsims, **kw)
        prof = be.profile()
        passes = be.profile_passes()
        be.profile_reset(False)
        prob.close()
    finally:
        for k in ("MUSE_PERSIST", "MUSE_LAZY", "MUSE_LEAN"):
            os.environ.pop(k, None)
    return res, prof, passes


def _assert_identical_solve(a, b, cov=True):
    assert len(a.history) == len(b.history)
    np.testing.assert_array_equal(a.theta, b.theta)
    np.testing.assert_array_equal(np.array(a.gs), np.array(b.gs))
    if cov:
        for key in ("J", "H", "Sigma"):
            np.testing.assert_array_equal(getattr(a, key), getattr(b, key), err_msg=key)
        np.testing.assert_array_equal(np.array(a.Hs), np.array(b.Hs))
        np.testing.assert_array_equal(a.metadata["fd_step"], b.metadata["fd_step"])
    for ha, hb in zip(a.history, b.history):
        for key in ("theta", "g_like_sims", "g_like_dat", "g_like", "g_prior", "H_inv_post", "H_inv_like"):
            np.testing.assert_array_equal(ha[key], hb[key], err_msg=key)
        assert ha["z_history_dat"] == hb["z_history_dat"]
        for key in ("iters", "fg_evals", "status", "gnorm"):
            np.testing.assert_array_equal(ha["z_history_sims"][key], hb["z_history_sims"][key], err_msg=key)


@pytest.mark.parametrize("name,d,nsims,prior,kw", [
    ("funnel", 512, 100, True, {}),                                      # the reference's example (warp-per-unit phases)
    ("funnel", 512, 2500, True, {}),                                     # more units than resident warps
    ("hiergauss", 5000, 50, False, {}),                                  # TMA-ring phases, nθ = 2
    ("funnel", 70000, 24, True, {}),                                     # several segments per unit
    ("hiergauss", 300, 30, True, dict(theta_rtol=0.0, maxsteps=7)),      # 3 passes in the launch, the rest on the chain of launches
    ("funnel", 6000, 40, True, dict(theta_rtol=1e-2, maxsteps=3)),       # the loop ends with the launch's last pass
    ("funnel", 6000, 40, True, dict(maxsteps=1)),                        # one pass, then the covariance stage
    ("funnel", 5000, 30, True, dict(z0="user")),                         # user start vector: streamed as the lazy chain's start row
    ("hiergauss", 700, 45, False, dict(z0="user", theta_rtol=0.0, maxsteps=4)),
])
def test_one_launch_solve_is_bit_identical_to_the_chain_of_launches(name, d, nsims, prior, kw):
    """solve_persist_kernel (one cooperative launch: passes, θ updates, get_H!'s fiducial solve and FD sims) against the chain of
    launches it replaces (streaming kernel + re-solve + theta_step_kernel per pass, cov_prep_kernel, two more chains): the same
    per-unit arithmetic and the same reduction tree, hence every number of the result bit for bit — and ONE kernel launch."""
    import museinference_jl_b200 as m
    oprob, fam, draws, xd = oracle_problem(name, d, nsims)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    pr = (lambda: m.NormalPrior([0.0, 0.1][:fam.ntheta], [3.0, 2.0][:fam.ntheta])) if prior else (lambda: None)
    kw = dict(kw)
    if kw.get("z0") == "user":
        kw["z0"] = 0.3 * np.cos(0.01 * np.arange(d))
    # lazy: ẑ recomputed from the base normals instead of stored and re-read; lean: the α = 1 trial from its closed form — neither
    # may change a single bit of the result
    for cov, lazy, lean in ((True, False, False), (False, False, False), (True, True, False), (True, False, True), (True, True, True), (False, True, True)):
        a, pa, passes = _solve_persist(m, xd, name, pr, rng, nsims, True, lazy, lean, get_covariance=cov, fused_driver="device", **kw)
        b, pb, _ = _solve_persist(m, xd, name, pr, rng, nsims, False, get_covariance=cov, fused_driver="device", **kw)
        _assert_identical_solve(a, b, cov=cov)
        n = len(a.history)
        assert pa["launches"] < pb["launches"]
        if n <= 3:
            assert pa["launches"] == 1 and pa["solve_launches"] == 1, pa
            user = "z0" in kw                           # a user start makes the first pass a "warm" one (it streams a start row)
            assert passes["cold"]["launches"] == (0 if user else 1) and passes["warm"]["launches"] == (n if user else n - 1)
            assert passes["fd"]["launches"] == (1 if cov else 0)
            assert all(v["ms"] > 0 for v in passes.values() if v["launches"])
            assert pa["solve_units"] == pb["solve_units"]
            if not lazy:
                assert pa["solve_bytes"] == pb["solve_bytes"]
            else:
                # lazy ẑ: a pass reads ξ, ν of every sim and the data — nothing else, and writes nothing, unless it is the third
                # pass of a loop that may go on (which materialises ẑ)
                stored = (nsims + 1) * 8 * d if (n == 3 and kw.get("maxsteps", 50) > 3) else 0
                assert passes["cold"]["bytes"] + passes["warm"]["bytes"] == n * (nsims * 16 * d + 8 * d + (8 * d if user else 0)) + stored


def test_one_launch_solve_gives_way_to_the_chain_when_units_leave_the_fast_path():
    """atol below round-off: every unit needs the L-BFGS history path, which only the chain's generic kernel has.  The launch
    notices (hand-backs), gives up, and the solve is re-run on the chain — same result as with MUSE_PERSIST=0, and the next
    solve with these parameters goes to the chain directly."""
    import museinference_jl_b200 as m
    name, d, nsims = "funnel", 600, 40
    oprob, fam, draws, xd = oracle_problem(name, d, nsims)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    pr = lambda: m.NormalPrior(0, 3)
    kw = dict(get_covariance=True, fused_driver="device", gradz_logLike_atol=1e-300, maxsteps=2)
    a, pa, _ = _solve_persist(m, xd, name, pr, rng, nsims, True, **kw)
    b, pb, _ = _solve_persist(m, xd, name, pr, rng, nsims, False, **kw)
    _assert_identical_solve(a, b)
    assert pa["launches"] == pb["launches"]          # the measured (second) solve did not try the one-launch form again
    assert max(h["z_history_sims"]["iters"].max() for h in a.history) >= 2


@pytest.mark.parametrize("name,d,nsims,kw", [
    ("funnel", 512, 2500, {}),                                           # warp-per-unit phases, 2 passes
    ("hiergauss", 5000, 50, dict(theta_rtol=0.0, maxsteps=3)),           # TMA-ring phases, all three pass blocks of the launch
    ("funnel", 9000, 33, dict(maxsteps=1)),                              # one pass: the other pass blocks stay untouched
])
def test_results_stored_to_the_host_mirrors_by_the_kernel_equal_the_copied_ones(name, d, nsims, kw):
    """The one-launch solve stores every per-unit result into the pinned host mirrors itself (copy_to_host: each CTA its slice of a
    pass's block once the pass has ended, CTA 0 the FD block and the state at exit).  MUSE_HOSTWRITE=0 leaves them in device memory
    and copies them behind the kernel: the host side must read the same bytes either way."""
    import os
    import museinference_jl_b200 as m
    oprob, fam, draws, xd = oracle_problem(name, d, nsims)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    pr = lambda: m.NormalPrior([0.0, 0.1][:fam.ntheta], [3.0, 2.0][:fam.ntheta])
    a, pa, _ = _solve_persist(m, xd, name, pr, rng, nsims, True, True, True, get_covariance=True, fused_driver="device", **kw)
    os.environ["MUSE_HOSTWRITE"] = "0"
    try:
        b, pb, _ = _solve_persist(m, xd, name, pr, rng, nsims, True, True, True, get_covariance=True, fused_driver="device", **kw)
    finally:
        os.environ.pop("MUSE_HOSTWRITE", None)
    assert pa["launches"] == pb["launches"] == 1
    _assert_identical_solve(a, b)


# ----------------------------------------------------------------------------- device-resident outer loop (csrc/muse_outer.cu)
def _solve_modes(m, xd, name, pr, rng, nsims, modes, **kw):
    res = {}
    for mode in modes:
        prob = m.SimpleMuseProblem(xd, name, pr())
        res[mode] = m.muse(prob, theta_start(name), rng=rng, nsims=nsims, fused_driver=mode, **kw)
        prob.close()
    return res


def _assert_same_solve(a, b, cov=True):
    assert len(a.history) == len(b.history)
    np.testing.assert_allclose(a.theta, b.theta, rtol=1e-11)
    np.testing.assert_allclose(np.array(a.gs), np.array(b.gs), rtol=1e-9, atol=1e-9)
    if cov:
        np.testing.assert_allclose(a.J, b.J, rtol=1e-10)
        np.testing.assert_allclose(a.H, b.H, rtol=1e-8, atol=1e-9 * np.abs(b.H).max())
        np.testing.assert_allclose(a.Sigma, b.Sigma, rtol=1e-7)
        np.testing.assert_allclose(np.array(a.Hs), np.array(b.Hs), rtol=1e-7, atol=1e-8 * np.abs(np.array(b.Hs)).max())
        np.testing.assert_allclose(a.metadata["fd_step"], b.metadata["fd_step"], rtol=1e-12)
    for ha, hb in zip(a.history, b.history):
        for key in ("theta", "theta_unreg", "g_like_sims", "g_like_dat", "g_like", "g_prior", "g_post", "H_inv_post",
                    "H_prior", "H_inv_like"):
            np.testing.assert_allclose(ha[key], hb[key], rtol=1e-9, atol=1e-9, err_msg=key)
        assert {k: v for k, v in ha["z_history_dat"].items() if k != "gnorm"} == {k: v for k, v in hb["z_history_dat"].items() if k != "gnorm"}
        np.testing.assert_array_equal(ha["z_history_sims"]["fg_evals"], hb["z_history_sims"]["fg_evals"])
        np.testing.assert_array_equal(ha["z_history_sims"]["status"], hb["z_history_sims"]["status"])


@pytest.mark.parametrize("name,d,nsims,prior,kw", [
    ("funnel", 512, 100, True, {}),                                     # the reference's example: 2 iterations + convergence test
    ("hiergauss", 5000, 50, False, dict(theta_rtol=1e-3, maxsteps=6)),   # several chunks of three passes
    ("hiergauss", 300, 30, True, dict(theta_rtol=0.0, maxsteps=7)),      # runs into maxsteps: 3 + 3 + 1 passes
    ("funnel", 6000, 40, True, dict(maxsteps=1)),                        # one pass, then the covariance stage
    ("funnel", 70000, 24, True, dict(theta_rtol=1e-2, maxsteps=3)),      # loop ends exactly at the end of a chunk
])
def test_device_resident_outer_loop_matches_host_loop(name, d, nsims, prior, kw):
    """muse_b200_muse_solve (θ update on the device, one synchronisation per chunk of passes) against muse_b200_muse_iterate +
    muse_b200_muse_covariance (host arithmetic between passes): same iteration count, histories, θ̂, J, H, Σ to round-off."""
    import museinference_jl_b200 as m
    oprob, fam, draws, xd = oracle_problem(name, d, nsims)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    pr = (lambda: m.NormalPrior([0.0, 0.1][:fam.ntheta], [3.0, 2.0][:fam.ntheta])) if prior else (lambda: None)
    res = _solve_modes(m, xd, name, pr, rng, nsims, ("device", "host"), get_covariance=True, **kw)
    _assert_same_solve(res["device"], res["host"])
    res = _solve_modes(m, xd, name, pr, rng, nsims, ("device", "host"), get_covariance=False, **kw)
    _assert_same_solve(res["device"], res["host"], cov=False)
    assert res["device"].J is None


def test_device_resident_outer_loop_against_oracle_and_errors():
    import museinference_jl_b200 as m
    name, d, nsims = "hiergauss", 2048, 60
    oprob, fam, draws, xd = oracle_problem(name, d, nsims)
    ref = O.muse(oprob, theta_start(name), nsims=nsims, get_covariance=True)
    prob = m.SimpleMuseProblem(xd, name)
    rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
    res = m.muse(prob, theta_start(name), rng=rng, nsims=nsims, get_covariance=True, fused_driver="device")
    assert len(res.history) == len(ref.history)
    np.testing.assert_allclose(res.theta, ref.theta, rtol=RTOL_EST)
    np.testing.assert_allclose(res.J, ref.J, rtol=RTOL_EST)
    np.testing.assert_allclose(res.H, ref.H, rtol=RTOL_EST, atol=RTOL_EST * np.abs(ref.H).max())
    np.testing.assert_allclose(res.Sigma, ref.Sigma, rtol=10 * RTOL_EST, atol=RTOL_EST * np.abs(ref.Sigma).max())
    np.testing.assert_allclose(np.array(res.gs), np.array(ref.gs), rtol=RTOL_SIM)
    # a second solve on the same handle (state is re-initialised), seeded draws this time
    a = m.muse(prob, theta_start(name), rng=5, nsims=nsims, get_covariance=True, fused_driver="device")
    b = m.muse(prob, theta_start(name), rng=5, nsims=nsims, get_covariance=True, fused_driver="host")
    _assert_same_solve(a, b)
    prob.close()
    # a NaN in the data: the θ-step kernel sees the NONFINITE status and the call fails loudly (src/interface.jl:170)
    xbad = xd.copy()
    xbad[7] = np.nan
    prob = m.SimpleMuseProblem(xbad, name)
    with pytest.raises(m.MuseBackendError):
        m.muse(prob, theta_start(name), rng=rng, nsims=nsims, get_covariance=True, fused_driver="device")
    prob.close()
    # corrgauss keeps the host loop even when asked for the device one
    from helpers import make_family
    famc = make_family("corrgauss", 128)
    oprobc, _, drawsc, xdc = oracle_problem("corrgauss", 128, 20, prior=O.NormalPrior(0, 3))
    probc = m.SimpleMuseProblem(xdc, "corrgauss", m.NormalPrior(0, 3), P=famc.P, L=famc.L)
    rc = m.muse(probc, [1.0], rng=m.BaseDraws(drawsc.xi, drawsc.nu, drawsc.xi_master, drawsc.nu_master), nsims=20,
                get_covariance=True, fused_driver="device")
    assert rc.Sigma is not None
    probc.close()


@pytest.mark.parametrize("d,kernel", [(512, 0), (6000, 0), (6000, 1)])
def test_zero_iteration_solve_from_zeros_leaves_a_zero_map(d, kernel):
    """ẑ of a unit whose zero start already satisfies the tolerance is zero(z) — not the buffer an earlier solve left behind.
    (Last in the file on purpose: added after the round's last GPU session, checked on the CPU through the host build of the
    kernels' code, tests/test_generic_solver_host.py.)"""
    fam, draws, xd = make_inputs("funnel", d, 8)
    be = _backend("funnel", d, 8, draws, xd, kernel=kernel)
    th = np.array([0.5])
    be.map_score(th, th, 1e-2, include_data=True, warm_start=0)
    assert np.abs(be.get_maps(0, 9)).max() > 0
    out = be.map_score(th, th, 1e6, include_data=True, warm_start=0)
    assert (out["iters"] == 0).all() and (out["fg_evals"] == 1).all() and (out["status"] == 0).all()
    np.testing.assert_array_equal(be.get_maps(0, 9), np.zeros((9, d)))
    be.close()


# ----------------------------------------------------------------------------- round 2: configs at size, multi-rank NCCL
def test_full_size_c2_funnel_against_closed_form_and_oracle():
    """BASELINE configs[1] at full size (funnel, d = 512, nsims = 10⁴, device Philox draws): 10 001 units = 1 251 CTAs of the
    warp-per-unit kernel in one launch.  Every unit: 1 iteration / 3 evaluations, score against the closed form
    g = ½ e^{-θ} s² ‖x‖² − d/2; 64 sampled units (and the data unit) against the oracle's L-BFGS on the device's own draws,
    for the cold pass and for the warm pass that follows (src/muse.jl:169-176, 181)."""
    import math
    import museinference_jl_b200 as m
    d, n, seed, atol = 512, 10000, 424242, 1e-2
    fam = O.make_family("funnel", d)
    xd, _ = fam.sample(np.array([0.0]), O.philox_normals(5, 0, 0, d), O.philox_normals(5, 0, 1, d))
    be = m.B200Backend("funnel", d, n)
    be.set_data(xd)
    be.seed_draws(seed)
    xi, nu = be.get_draws(0, n + 1)
    draws = O.Draws(xi[:n], nu[:n], xi[n], nu[n])
    prob = O.OracleProblem(fam, xd, draws)
    th0, th1 = np.array([1.0]), np.array([0.55])
    out0 = be.map_score(th0, th0, atol, include_data=True, warm_start=0)
    z0 = be.get_maps(0, n + 1)
    out1 = be.map_score(th1, th1, atol, include_data=True, warm_start=1)
    z1 = be.get_maps(0, n + 1)
    assert be.profile()["redo_units"] == 0
    for out, th in ((out0, th0), (out1, th1)):
        assert (out["iters"] == 1).all() and (out["fg_evals"] == 3).all() and (out["status"] == 0).all()
        s, a = 1.0 / (1.0 + math.exp(-th[0])), math.exp(-th[0])
        x = np.vstack([xd[None], math.exp(0.5 * th[0]) * xi[:n] + nu[:n]])
        np.testing.assert_allclose(out["g"][:, 0], 0.5 * a * s * s * np.einsum("ij,ij->i", x, x) - d / 2, rtol=1e-9)
    sample = [0] + sorted(np.random.default_rng(0).choice(np.arange(1, n + 1), size=64, replace=False).tolist())
    for u in sample:
        x = xd if u == 0 else prob.sample_x_z(u - 1, th0)[0]
        zh, g, soln = O.map_score_unit(prob, x, np.zeros(d), th0, atol)
        assert (out0["iters"][u], out0["fg_evals"][u]) == (soln.iterations, soln.f_calls)
        np.testing.assert_allclose(out0["g"][u], g, rtol=RTOL_SIM)
        np.testing.assert_allclose(z0[u], zh, rtol=RTOL_SIM, atol=1e-12)
        x = xd if u == 0 else prob.sample_x_z(u - 1, th1)[0]
        zh1, g1, soln1 = O.map_score_unit(prob, x, zh, th1, atol)
        assert (out1["iters"][u], out1["fg_evals"][u]) == (soln1.iterations, soln1.f_calls)
        np.testing.assert_allclose(out1["g"][u], g1, rtol=RTOL_SIM)
        np.testing.assert_allclose(z1[u], zh1, rtol=RTOL_SIM, atol=1e-12)
    be.close()


def test_full_size_c5_corrgauss_sampled_units_match_oracle_lbfgs():
    """BASELINE configs[4] at full size (d = 4 096, nsims = 8 192, host Philox draws uploaded): the data unit and 8 sampled
    sims against the oracle's honest L-BFGS + Hager–Zhang (C port, P·z evaluated at every trial point) at d = 4 096 —
    identical iteration and evaluation counts, ẑ and g within rtol 1e-8 — for the cold pass and the warm pass after it."""
    import museinference_jl_b200 as m
    from bench import corr_consts
    from oracle import cport
    d, n, atol = 4096, 8192, 1e-2
    P, L = corr_consts(d)
    rng = np.random.Generator(np.random.Philox(77))
    xi, nu = rng.standard_normal((n, d)), rng.standard_normal((n, d))
    xim, num = rng.standard_normal(d), rng.standard_normal(d)
    xd = L @ rng.standard_normal(d) + rng.standard_normal(d)
    be = m.B200Backend("corrgauss", d, n, P=P, L=L)
    be.set_data(xd)
    be.set_draws(xi, nu, xim, num)
    th0, th1 = np.array([1.0]), np.array([0.6])
    sims = sorted(np.random.default_rng(1).choice(n, size=8, replace=False).tolist())
    out0 = be.map_score(th0, th0, atol, include_data=True, warm_start=0)
    z0 = np.vstack([be.get_maps(0, 1)] + [be.get_maps(1 + k, 1) for k in sims])
    out1 = be.map_score(th1, th1, atol, include_data=True, warm_start=1)
    z1 = np.vstack([be.get_maps(0, 1)] + [be.get_maps(1 + k, 1) for k in sims])
    be.close()
    units = [0] + [1 + k for k in sims]
    ref0 = cport.map_score(3, xi[sims], nu[sims], xd, th0, th0, atol, True, 0, want_z=True, P=P, L=L)
    ref1 = cport.map_score(3, xi[sims], nu[sims], xd, th1, th1, atol, True, 1, z_start=ref0["z"], want_z=True, P=P, L=L)
    for out, z, ref in ((out0, z0, ref0), (out1, z1, ref1)):
        np.testing.assert_array_equal(out["iters"][units], ref["iters"])
        np.testing.assert_array_equal(out["fg_evals"][units], ref["fg_evals"])
        np.testing.assert_array_equal(out["status"][units], ref["status"])
        np.testing.assert_allclose(out["g"][units], ref["g"], rtol=RTOL_SIM)
        np.testing.assert_allclose(z, ref["z"], rtol=RTOL_SIM, atol=1e-11)
    assert out0["iters"][units].min() >= 4


def _nccl_worker(rank, world, port, cases, q):
    import os
    import sys
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch
    import torch.distributed as dist
    import museinference_jl_b200 as m
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        pool = m.ShardPool(device=rank)
        outs = []
        for name, d, nsims, th0, xd, P, L, prior, fused, exchange, kw in cases:
            # exchange "p2p": the one-launch solve, score rows stored into the peer's memory by the kernel; "nccl": the chain
            # of launches with ncclAllGather between pass and θ-step
            if exchange == "nccl":
                os.environ["MUSE_EXCHANGE"] = "nccl"
            else:
                os.environ.pop("MUSE_EXCHANGE", None)
            prob = m.SimpleMuseProblem(xd, name, m.NormalPrior(0, 3) if prior else None, P=P, L=L)
            launches = p2p = None
            for seed in (11, 12):          # twice: the second solve of a shape goes through the library's captured graph
                if seed == 12 and prob._backend is not None:
                    prob._backend.profile_reset(True)
                res = m.muse(prob, th0, rng=seed, nsims=nsims, get_covariance=True, pool=pool, fused_driver=fused, **kw)
            if prob._backend is not None:
                launches = prob._backend.profile()["solve_launches"]
                p2p = prob._backend.p2p_info()
            outs.append((res.theta, np.array(res.gs), res.H, res.J, res.Sigma, len(res.history), launches, p2p))
            prob.close()
        q.put((rank, outs))
    finally:
        dist.destroy_process_group()


def test_two_rank_nccl_solve_is_bit_identical_to_one_gpu():
    """Two ranks, one GPU each (skipped on a one-GPU box): contiguous sim shards, the score exchange between the passes, θ
    updated identically on both ranks.  θ̂, gs, H, J, Σ of every rank must equal the single-GPU solve BIT FOR BIT for F1, F2
    (device-resident loop) and F3 (host loop), and through the line-by-line driver — per-sim results do not depend on the
    shard and every reduction is ordered by the global sim index (src/muse.jl:169, 177-183)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import socket
    import torch.multiprocessing as mp
    import museinference_jl_b200 as m
    cases = []
    long_loop = dict(theta_rtol=0.0, maxsteps=5)       # 3 passes in the one-launch solve, 2 more on the chain of launches
    for name, d, nsims, prior, fused, exchange, kw in (("funnel", 6000, 203, True, True, "p2p", {}), ("funnel", 6000, 203, True, True, "nccl", {}),
                                                       ("hiergauss", 5001, 120, False, True, "p2p", {}), ("hiergauss", 5001, 120, False, True, "nccl", {}),
                                                       ("funnel", 512, 301, True, True, "p2p", {}), ("funnel", 512, 5, True, True, "p2p", {}),
                                                       ("hiergauss", 300, 31, True, True, "p2p", long_loop), ("funnel", 4500, 40, True, True, "p2p", dict(maxsteps=3, theta_rtol=0.0)),
                                                       ("corrgauss", 256, 100, True, True, "p2p", {}),
                                                       ("hiergauss", 700, 64, True, False, "p2p", {})):
        fam, _, xd = make_inputs(name, d, 1)
        cases.append((name, d, nsims, theta_start(name), xd, getattr(fam, "P", None), getattr(fam, "L", None), prior, fused, exchange, kw))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, cases, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for i, (name, d, nsims, th0, xd, P, L, prior, fused, exchange, kw) in enumerate(cases):
        prob = m.SimpleMuseProblem(xd, name, m.NormalPrior(0, 3) if prior else None, P=P, L=L)
        ref = m.muse(prob, th0, rng=12, nsims=nsims, get_covariance=True, fused_driver=fused, **kw)
        prob.close()
        for rank in (0, 1):
            theta, gs, H, J, Sigma, nhist, launches, p2p = got[rank][i]
            assert nhist == len(ref.history)
            if name != "corrgauss" and fused is True:
                # the exchange that ran is the one asked for: one launch per solve with the peer-mapped buffers, a chain otherwise
                assert p2p[1] == (exchange == "p2p"), (name, exchange, p2p)
                assert (launches == 1) == (exchange == "p2p" and nhist <= 3), (name, exchange, launches)     # solver launches of the solve
            np.testing.assert_array_equal(gs, np.array(ref.gs), err_msg=f"{name} rank {rank}: gs")
            np.testing.assert_array_equal(theta, ref.theta, err_msg=f"{name} rank {rank}: θ")
            np.testing.assert_array_equal(H, ref.H, err_msg=f"{name} rank {rank}: H")
            np.testing.assert_array_equal(J, ref.J)
            np.testing.assert_array_equal(Sigma, ref.Sigma)
