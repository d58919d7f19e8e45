"""Correlated Gaussian: per-iteration scores of a sharded (2-rank NCCL) solve vs the single-process solve (diagnostics).
Run rank-less for the single-process half, under torchrun for the sharded half; both write gpurun_out/f3_hist_*.npz."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import museinference_jl_b200 as m
from bench import corr_consts, observed_data
d, n = 512, 400
P, L = corr_consts(d)
x = observed_data("corrgauss", d, L)
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    import torch, torch.distributed as dist
    lr = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    pool = m.ShardPool(device=lr)
else:
    pool = m.LocalPool()
prob = m.SimpleMuseProblem(x, "corrgauss", m.NormalPrior(0, 3), P=P, L=L)
res = m.muse(prob, [1.0], rng=314159, nsims=n, get_covariance=True, pool=pool)
if pool.rank == 0:
    np.savez(f"gpurun_out/f3_hist_w{world}.npz", theta=np.array([h["theta"] for h in res.history]), gs=np.array([h["g_like_sims"] for h in res.history]),
             gdat=np.array([h["g_like_dat"] for h in res.history]), theta_final=res.theta, Hs=np.array(res.Hs), J=res.J, H=res.H)
prob.close()
if world > 1:
    dist.destroy_process_group()
