"""Multi-rank path on CPU: world_size 2, gloo.  Sims are sharded in contiguous blocks, each rank solves its
block (test double for the GPU backend), the N×nθ score matrix is all-gathered — the one exchange step of
the path — and every rank must arrive at the single-process result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, d, nsims, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import museinference_jl_b200 as m
    import oracle as O
    from fake_backend import FakeBackend
    from helpers import oracle_problem, theta_start
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oprob, fam, draws, xd = oracle_problem(name, d, nsims, prior=None)
        prob = m.SimpleMuseProblem(xd, name, backend_factory=FakeBackend)
        pool = m.ShardPool()
        rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
        res = m.muse(prob, theta_start(name), rng=rng, nsims=nsims, get_covariance=True, pool=pool)
        off, cnt = pool.shard(nsims)
        be = prob._backend
        # seeded (throughput-mode) draws are keyed by the global sim index: shard-invariant
        res2 = m.muse(prob, theta_start(name), rng=77, nsims=nsims, get_covariance=True, pool=pool)
        q.put((rank, res.theta, res.J, res.H, np.array(res.gs), (off, cnt, be.nsims, be.nsims_h), res2.theta, res2.H))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,d,nsims", [("funnel", 96, 31), ("hiergauss", 64, 24)])
def test_two_rank_gloo_matches_single_process(name, d, nsims):
    import torch.multiprocessing as mp
    import museinference_jl_b200 as m
    import oracle as O
    from fake_backend import FakeBackend
    from helpers import oracle_problem, theta_start

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, d, nsims, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=240) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    oprob, fam, draws, xd = oracle_problem(name, d, nsims, prior=None)
    ref = O.muse(oprob, theta_start(name), nsims=nsims, get_covariance=True)
    prob = m.SimpleMuseProblem(xd, name, backend_factory=FakeBackend)
    single2 = m.muse(prob, theta_start(name), rng=77, nsims=nsims, get_covariance=True)
    offs, cnts = m.block_partition(nsims, 2)
    hoffs, hcnts = m.block_partition(max(1, nsims // 10), 2)
    for rank, theta, J, H, gs, geo, theta2, H2 in outs:
        assert geo == (offs[rank], cnts[rank], cnts[rank], hcnts[rank])
        np.testing.assert_allclose(theta, ref.theta, rtol=1e-12)
        np.testing.assert_allclose(J, ref.J, rtol=1e-12)
        np.testing.assert_allclose(H, ref.H, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(gs, np.array(ref.gs), rtol=1e-13)
        np.testing.assert_allclose(theta2, single2.theta, rtol=1e-12)
        np.testing.assert_allclose(H2, single2.H, rtol=1e-10, atol=1e-12)
    # ranks agree bit for bit (identical deterministic host arithmetic on the gathered scores)
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    np.testing.assert_array_equal(outs[0][3], outs[1][3])


def _worker_topup(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import museinference_jl_b200 as m
    from fake_backend import FakeBackend
    from helpers import oracle_problem
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oprob, fam, draws, xd = oracle_problem("hiergauss", 48, 21)
        prob = m.SimpleMuseProblem(xd, "hiergauss", backend_factory=FakeBackend)
        pool = m.ShardPool()
        rng = m.BaseDraws(draws.xi, draws.nu, draws.xi_master, draws.nu_master)
        res = m.MuseResult(theta=np.array([0.2, 0.1]))
        getJ, getH = getattr(m, "get_J!"), getattr(m, "get_H!")
        getJ(res, prob, rng=rng, nsims=8, pool=pool)
        getJ(res, prob, rng=rng, nsims=21, pool=pool)           # top-up across the shard boundary (src/muse.jl:499-506)
        getH(res, prob, rng=rng, nsims=5, pool=pool)
        # top-up whose existing rows already cover the whole of rank 0's shard (11 sims): rank 0 has nothing to simulate
        res3 = m.MuseResult(theta=np.array([0.2, 0.1]))
        getJ(res3, prob, rng=rng, nsims=14, pool=pool)
        getJ(res3, prob, rng=rng, nsims=21, pool=pool)
        assert np.array_equal(np.array(res3.gs), np.array(res.gs)) and np.array_equal(res3.J, res.J)
        # the same problem with θ = (μ, σ), σ > 0 (transform_θ = (μ, log σ)): full solve sharded over the two ranks
        import oracle as O
        tprob = m.SimpleMuseProblem(xd, "hiergauss", theta_transform=("identity", "log"), backend_factory=FakeBackend)
        tres = m.muse(tprob, [0.5, np.exp(0.3)], rng=rng, nsims=21, get_covariance=True, pool=pool)
        q.put((rank, np.array(res.gs), res.J, res.H, res.Sigma, tres.theta, tres.J, tres.H))
    finally:
        dist.destroy_process_group()


def test_two_rank_get_J_top_up_and_get_H():
    """get_J! called twice (8, then 21 sims) and get_H! on two gloo ranks against the single-process oracle: the top-up
    only simulates the missing sims, whichever rank owns them, and every rank ends with the same gs, J, H, Σ."""
    import torch.multiprocessing as mp
    import oracle as O
    from helpers import oracle_problem
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_topup, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=240) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    oprob, *_ = oracle_problem("hiergauss", 48, 21)
    ref = O.MuseResult(theta=np.array([0.2, 0.1]))
    O.get_J_bang(ref, oprob, nsims=21)
    O.get_H_bang(ref, oprob, nsims=5)
    tfam = O.TransformedFamily(oprob.family, ("identity", "log"))
    tref = O.muse(O.OracleProblem(tfam, oprob.x, oprob.draws), [0.5, np.exp(0.3)], nsims=21, get_covariance=True)
    for rank, gs, J, H, Sigma, t_theta, t_J, t_H in outs:
        np.testing.assert_allclose(t_theta, tref.theta, rtol=1e-9)
        np.testing.assert_allclose(t_J, tref.J, rtol=1e-9)
        np.testing.assert_allclose(t_H, tref.H, rtol=1e-7, atol=1e-9)
        assert gs.shape == (21, 2)
        np.testing.assert_allclose(gs, np.array(ref.gs), rtol=1e-12)
        np.testing.assert_allclose(J, ref.J, rtol=1e-12)
        np.testing.assert_allclose(H, ref.H, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(Sigma, ref.Sigma, rtol=1e-9)
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    np.testing.assert_array_equal(outs[0][5], outs[1][5])


def test_block_partition():
    import museinference_jl_b200 as m
    assert m.block_partition(10, 4) == ([0, 3, 6, 8], [3, 3, 2, 2])
    assert m.block_partition(2, 4) == ([0, 1, 2, 2], [1, 1, 0, 0])
    assert m.block_partition(0, 2) == ([0, 0], [0, 0])
