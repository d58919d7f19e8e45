"""compute-sanitizer run over the code added late in round 2 only (the rest: scripts/sanitize.py): the one-launch solve with its
stores to the pinned host mirrors and the completion word, the trimmed draws kernel, the two-layer family on the lock-step solver
(sampling, elementwise block product, score, implicit-diff CG)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import museinference_jl_b200 as m

x = np.random.default_rng(0).standard_normal(4500)
prob = m.SimpleMuseProblem(x, "funnel", m.NormalPrior(0, 3))
r = m.muse(prob, [1.0], rng=3, nsims=40, get_covariance=True)          # seeded draws (philox_draws_tab2_kernel) + solve_persist_kernel
print("funnel", r.theta, len(r.history))
prob.close()
x = np.random.default_rng(1).standard_normal(301)
prob = m.SimpleMuseProblem(x, "funnel", m.NormalPrior(0, 3))           # warp-per-unit phases, odd d
r = m.muse(prob, [1.0], rng=5, nsims=33, get_covariance=True)
print("funnel small", r.theta)
prob.close()
x = np.random.default_rng(2).standard_normal(600)
prob = m.SimpleMuseProblem(x, "twolayer", m.NormalPrior(0, 3))
r = m.muse(prob, [0.5], rng=7, nsims=24, get_covariance=True)
res = m.MuseResult(theta=r.theta.copy())
getattr(m, "get_H!")(res, prob, rng=7, nsims=5, implicit_diff=True)
print("twolayer", r.theta, res.H)
prob.close()
print("done")
