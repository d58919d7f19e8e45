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


def test_block_partition():
    import museinference_jl_b200 as m
    assert m.block_partition(10, 4) == ([0, 3, 6, 8], [3, 3, 2, 2])
    assert m.block_partition(2, 4) == ([0, 1, 2, 2], [1, 1, 0, 0])
    assert m.block_partition(0, 2) == ([0, 0], [0, 0])
