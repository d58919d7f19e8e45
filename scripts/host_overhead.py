"""Where the host time of one full solve goes (diagnostics): wall time inside the C-ABI calls vs Python around them."""
import cProfile, pstats, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import museinference_jl_b200 as m
d = int(os.environ.get("MUSE_D", 65536)); n = int(os.environ.get("MUSE_N", 2048))
x = np.random.default_rng(0).standard_normal(d) * 1.4
prob = m.SimpleMuseProblem(x, "funnel", m.NormalPrior(0, 3))
for _ in range(3):
    m.muse(prob, [1.0], rng=5, nsims=n, get_covariance=True)
be = prob._backend
acc = {}
def wrap(name):
    f = getattr(be, name)
    def g(*a, **k):
        t = time.perf_counter(); r = f(*a, **k); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t; return r
    setattr(be, name, g)
for nm in ("map_score", "fd_jacobian", "muse_iterate", "muse_covariance", "muse_solve"):
    wrap(nm)
K = int(os.environ.get("MUSE_K", 50))
be.profile_reset(True)
t0 = time.perf_counter()
for _ in range(K):
    m.muse(prob, [1.0], rng=5, nsims=n, get_covariance=True)
wall = (time.perf_counter() - t0) / K
p = be.profile()
print("per solve: wall %.3f ms | in C calls %.3f ms %s | solver chains (events) %.3f ms | python outside C %.3f ms"
      % (wall * 1e3, sum(acc.values()) / K * 1e3, {k: round(v / K * 1e3, 3) for k, v in acc.items()}, p["solve_ms"] / K, (wall - sum(acc.values()) / K) * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    m.muse(prob, [1.0], rng=5, nsims=n, get_covariance=True)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
