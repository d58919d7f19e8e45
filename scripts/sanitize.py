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
    if os.environ.get("MUSE_SANITIZE_NO_GENERIC"):      # synccheck stops at the generic kernel's named barriers (profiles/r01_sanitizer.md)
        print(family, d, n, kw, "theta", r.theta)
        prob.close()
        return
    o = be.map_score(th, th, 1e-300 if family != "corrgauss" else 1e-2, include_data=True, warm_start=0)   # history path / hand-back
    print(family, d, n, kw, "theta", r.theta, "iters", np.bincount(o["iters"])[:8], "redo", be.profile()["redo_units"])
    prob.close()

run("funnel", 4500, 40)                 # streaming kernel, 3 chunks, ragged tail
run("hiergauss", 20001, 24)             # streaming kernel, 2 segments, odd d
run("funnel", 300, 50)                  # warp-per-unit generic kernel
if not os.environ.get("MUSE_SANITIZE_NO_GENERIC"):
    run("funnel", 3000, 16, kernel=1, group=256, cluster=2)   # cluster groups (DSMEM reductions)
run("corrgauss", 256, 20)               # F3: TMA DGEMM + lock-step kernels
# round 2: the one-launch solve in its forms (the muse() calls above already ran it with the defaults), implicit diff, user start
for env in ({"MUSE_LAZY": "0", "MUSE_LEAN": "0"}, {"MUSE_FUNNEL_SPEC": "0"}, {"MUSE_PERSIST": "0"}):
    os.environ.update(env)
    run("funnel", 4500, 24)
    for k in env:
        os.environ.pop(k)
x = np.random.default_rng(1).standard_normal(4100)
prob = m.SimpleMuseProblem(x, "hiergauss")
r = m.muse(prob, [0.5, 0.3], rng=4, nsims=20, get_covariance=True, z0=0.1 * np.ones(4100), maxsteps=4, theta_rtol=0.0)
res = m.MuseResult(theta=r.theta.copy())
getattr(m, "get_H!")(res, prob, rng=4, nsims=5, implicit_diff=True)
prob.close()
P, L = corr_consts(256)
prob = m.SimpleMuseProblem(np.random.default_rng(2).standard_normal(256), "corrgauss", m.NormalPrior(0, 3), P=P, L=L)
res = m.MuseResult(theta=np.array([0.4]))
getattr(m, "get_H!")(res, prob, rng=4, nsims=6, implicit_diff=True)
prob.close()
print("done")
