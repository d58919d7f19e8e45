"""Small driver for ncu: seeds draws, then runs cold / warm / truth passes and one FD Jacobian on the
C3 shape (funnel d=65536, nsims=2048) through the C ABI.  Kernel launches in order:
philox ×1, iso_solver (cold), iso_solver (warm), iso_solver (cold), iso_solver (warm), fid, FD."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import museinference_jl_b200 as m  # noqa: E402

d = int(os.environ.get("MUSE_D", 65536))
n = int(os.environ.get("MUSE_N", 2048))
group = int(os.environ.get("MUSE_GROUP", 0))
cluster = int(os.environ.get("MUSE_CLUSTER", 0))
kernel = int(os.environ.get("MUSE_KERNEL", 0))
be = m.B200Backend("funnel", d, n, group=group, cluster=cluster, kernel=kernel)
rng = np.random.Generator(np.random.Philox(1))
be.set_data(rng.standard_normal(d) * 1.4)
be.seed_draws(42)
th0, th1 = np.array([1.0]), np.array([0.45])
be.profile_reset(True)
for rep in range(2):
    o = be.map_score(th0, th0, 1e-2, include_data=True, warm_start=0)
    o = be.map_score(th1, th1, 1e-2, include_data=True, warm_start=1)
Hs, st = be.fd_jacobian(th1, np.array([1e-3]), n // 10, 1e-2)
p = be.profile()
print("profile:", p)
print("iters", np.bincount(o["iters"]), "fg", np.bincount(o["fg_evals"]), "H mean", Hs.mean())
be.close()
