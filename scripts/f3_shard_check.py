"""Is a correlated-Gaussian unit's result independent of its local row / shard?  (diagnostics)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import museinference_jl_b200 as m
from bench import corr_consts
d, n = 512, 400
P, L = corr_consts(d)
x = np.random.default_rng(0).standard_normal(d)
th = np.array([1.0])
full = m.B200Backend("corrgauss", d, n, P=P, L=L); full.set_data(x); full.seed_draws(7)
part = m.B200Backend("corrgauss", d, 200, sim_offset=200, P=P, L=L); part.set_data(x); part.seed_draws(7)
of = full.map_score(th, th, 1e-2, include_data=True, warm_start=0)
op = part.map_score(th, th, 1e-2, include_data=True, warm_start=0)
zf, zp = full.get_maps(201, 200), part.get_maps(1, 200)
print("data unit g equal:", of["g"][0, 0] == op["g"][0, 0], of["g"][0, 0] - op["g"][0, 0])
dg = of["g"][201:, 0] - op["g"][1:, 0]
print("sims g: max |diff| %.3e, #different %d of 200; iters equal %s" % (np.abs(dg).max(), (dg != 0).sum(), (of["iters"][201:] == op["iters"][1:]).all()))
print("sims z: max |diff| %.3e" % np.abs(zf - zp).max())
# the draws themselves (W = L ξ rows) through the first evaluation: cold start ⇒ x = σW + ν; compare ẑ after 0 rounds is not exposed, so compare W via a 1-iteration cap
a1 = m.B200Backend("corrgauss", d, n, P=P, L=L, max_iters=1); a1.set_data(x); a1.seed_draws(7)
b1 = m.B200Backend("corrgauss", d, 200, sim_offset=200, P=P, L=L, max_iters=1); b1.set_data(x); b1.seed_draws(7)
a1.map_score(th, th, 1e-2, include_data=True, warm_start=0); b1.map_score(th, th, 1e-2, include_data=True, warm_start=0)
print("after 1 iteration: max |z diff| %.3e" % np.abs(a1.get_maps(201, 200) - b1.get_maps(1, 200)).max())
a1.map_score(th, th, 1e-2, include_data=False, warm_start=2); b1.map_score(th, th, 1e-2, include_data=False, warm_start=2)
