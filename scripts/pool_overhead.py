"""Software overhead of the exchange step: LocalPool vs a world-size-1 NCCL ShardPool on one GPU (diagnostics)."""
import os, sys, time
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import museinference_jl_b200 as m
d, n = 65536, 2048
x = np.random.default_rng(0).standard_normal(d) * 1.4
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29544", rank=0, world_size=1, device_id=torch.device("cuda", 0))
for name, pool in (("local", m.LocalPool()), ("nccl-1", m.ShardPool(device=0))):
    prob = m.SimpleMuseProblem(x, "funnel", m.NormalPrior(0, 3))
    for _ in range(3):
        m.muse(prob, [1.0], rng=5, nsims=n, get_covariance=True, pool=pool)
    acc = {}
    for nm in ("allgather_device_scores", "allgather_host_rows", "allgather_rows"):
        if hasattr(pool, nm):
            def timed(*a, _f=getattr(pool, nm), _n=nm, **k):
                t = time.perf_counter(); r = _f(*a, **k); acc[_n] = acc.get(_n, 0.0) + time.perf_counter() - t; return r
            setattr(pool, nm, timed)
    K = 50
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(K):
        m.muse(prob, [1.0], rng=5, nsims=n, get_covariance=True, pool=pool)
    torch.cuda.synchronize()
    print(name, "ms per solve %.3f" % ((time.perf_counter() - t0) / K * 1e3), {k: round(v / K * 1e3, 3) for k, v in acc.items()})
    prob.close()
dist.destroy_process_group()
