"""Small driver for ncu: one cold pass of the correlated-Gaussian family at the C5 shape (d=4096, nsims=8192 unless
MUSE_N is set) through the C ABI.  Launch order: philox, dgemm (W = ξ·Lᵀ), init, start, then per round dgemm + iter."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import museinference_jl_b200 as m
from bench import corr_consts
d = int(os.environ.get("MUSE_D", 4096)); n = int(os.environ.get("MUSE_N", 8192))
P, L = corr_consts(d)
be = m.B200Backend("corrgauss", d, n, P=P, L=L)
be.set_data(L @ np.random.default_rng(0).standard_normal(d)); be.seed_draws(42)
th = np.array([1.0])
be.profile_reset(True)
o = be.map_score(th, th, 1e-2, include_data=True, warm_start=0)
print("profile:", be.profile())
print("iters", np.bincount(o["iters"]), "fg", np.bincount(o["fg_evals"])[-6:], "status", np.bincount(o["status"]))
be.close()
