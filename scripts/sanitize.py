"""Small end-to-end run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
streaming + replay + generic re-solve (history path), warp-per-unit kernel, FD pass, draws, F3 lock-step + DGEMM."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import museinference_jl_b200 as m
from bench import corr_consts

def run(family, d, n, **kw):
    P = L = None
    if family == "corrgauss":
        P, L = corr_consts(d)
    x = np.random.default_rng(0).standard_normal(d)
    prob = m.SimpleMuseProblem(x, family, None if family == "hiergauss" else m.NormalPrior(0, 3), P=P, L=L, **kw)
    th0 = [0.5, 0.3] if family == "hiergauss" else [1.0]
    r = m.muse(prob, th0, rng=3, nsims=n, get_covariance=True)
    be = prob._backend
    th = np.array(th0)
    o = be.map_score(th, th, 1e-300 if family != "corrgauss" else 1e-2, include_data=True, warm_start=0)   # history path / hand-back
    print(family, d, n, kw, "theta", r.theta, "iters", np.bincount(o["iters"])[:8], "redo", be.profile()["redo_units"])
    prob.close()

run("funnel", 4500, 40)                 # streaming kernel, 3 chunks, ragged tail
run("hiergauss", 20001, 24)             # streaming kernel, 2 segments, odd d
run("funnel", 300, 50)                  # warp-per-unit generic kernel
run("funnel", 3000, 16, kernel=1, group=256, cluster=2)   # cluster groups (DSMEM reductions)
run("corrgauss", 200, 20)               # F3: DGEMM + lock-step kernels
print("done")
