"""Generates tests/golden/*.json from the NumPy oracle (run from the repo root:
``python tests/golden/make_golden.py``).  The reference (Julia) cannot run in this image and ships
no golden vectors; these fixtures freeze the *oracle's* outputs on seeded inputs so
that (a) oracle regressions are caught and (b) the GPU path can be checked on the GPU box, where
/root/reference does not exist, without re-deriving anything.  Inputs are regenerated from the
seeds with oracle/philox.py (counter-based, platform independent)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O  # noqa: E402
from helpers import oracle_problem, theta_start  # noqa: E402

CASES = [
    dict(name="funnel_d512_n100", family="funnel", d=512, nsims=100, seed=1234, prior=True, atol=1e-2),
    dict(name="funnel_d64_n16_tight", family="funnel", d=64, nsims=16, seed=77, prior=True, atol=1e-10),
    dict(name="hiergauss_d300_n40", family="hiergauss", d=300, nsims=40, seed=4321, prior=False, atol=1e-2),
    # θ = (μ, σ) with σ > 0: transform_θ = (μ, log σ) (oracle/families.py TransformedFamily), θ₀ = (0.5, e^0.3)
    dict(name="hiergauss_sigma_d200_n30", family="hiergauss", d=200, nsims=30, seed=2468, prior=False, atol=1e-2,
         transform=["identity", "log"]),
    # the toy hierarchy of the Turing adapter's docstring (src/turing.jl:63-79) at its own size: n = 512 components per layer
    dict(name="twolayer_d1024_n60", family="twolayer", d=1024, nsims=60, seed=1357, prior=True, atol=1e-2),
]


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for c in CASES:
        if len(sys.argv) > 1 and c["name"] not in sys.argv[1:]:
            continue
        prior = O.NormalPrior(0, 3) if c["prior"] else None
        prob, fam, draws, xd = oracle_problem(c["family"], c["d"], c["nsims"], seed=c["seed"], prior=prior)
        th0 = theta_start(c["family"])
        if c.get("transform"):
            prob = O.OracleProblem(O.TransformedFamily(fam, c["transform"]), xd, draws, prior)
            th0 = prob.inv_transform_theta(th0)
        res = O.muse(prob, th0, nsims=c["nsims"], gradz_logLike_atol=c["atol"], get_covariance=True, save_MAPs=True)
        h0 = res.history[0]
        fix = dict(
            case=c, theta0=th0.tolist(), theta=res.theta.tolist(), J=res.J.tolist(), H=res.H.tolist(),
            Sigma=res.Sigma.tolist(), n_outer=len(res.history),
            gs=np.array(res.gs).tolist(), Hs=np.array(res.Hs).tolist(),
            iter1=dict(g_dat=h0["g_like_dat"].tolist(), g_sims=h0["g_like_sims"].tolist(),
                       z_dat_sum=float(np.sum(h0["z_dat"])), z_dat_sumsq=float(np.dot(h0["z_dat"], h0["z_dat"])),
                       z_sims_sumsq=[float(np.dot(z, z)) for z in h0["z_sims"]],
                       iters=[s.iterations for s in h0["z_history_sims"]],
                       fg=[s.f_calls for s in h0["z_history_sims"]]),
            xdat_head=xd[:4].tolist(), xi0_head=draws.xi[0, :4].tolist(),
        )
        with open(os.path.join(out_dir, c["name"] + ".json"), "w") as fh:
            json.dump(fix, fh, indent=1)
        print("wrote", c["name"], "theta", res.theta, "n_outer", len(res.history))


if __name__ == "__main__":
    main()
