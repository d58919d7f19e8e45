"""CPU oracle for the MUSE per-simulation hot path  --  TEST INFRASTRUCTURE ONLY.

This package is a plain NumPy (and, under ``oracle/csrc``, plain C) restatement of the
algorithm the reference executes on the path named by ``BASELINE.json: north_star``:

    muse! / get_J! / get_H!           /root/reference/src/muse.jl:112-250, 296-333, 407-450, 484-549
    pjacobian, split_rng (semantics)  /root/reference/src/util.jl:9-26, 85-92
    default ẑ_at_θ (L-BFGS MAP)       /root/reference/src/interface.jl:162-171
    SimpleMuseProblem closures        /root/reference/src/simple.jl:58-95

It is the *checker*, never the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``museinference.jl_b200/`` imports it and the CUDA product has no CPU fallback.

PARITY UNPINNED.  The reference is pure Julia and Julia is not installed in this image (nor
on the GPU box); it ships no golden vectors, no known-answer tests and no fixtures for this
path (its only assertion is the statistical bound ``result.dist.μ / result.dist.σ < 2``,
/root/reference/test/runtests.jl:31,56,81).  The numerical core of the path lives in
third-party Julia packages that are not vendored under /root/reference and are pinned only
by compat ranges (Project.toml:34-50):

    Optim "1.5" (+ LineSearches, NLSolversBase)   L-BFGS(m=10) + InitialStatic + HagerZhang
    FiniteDifferences "0.12.20"                    central_fdm(p,1) with an explicit step, and its adaptive step estimate
    Statistics / CovarianceEstimation "0.2.7"      mean / var / std / cov (corrected)
    Distributions "0.25.36", Random                MvNormal draws, Xoshiro streams

Their published algorithms are restated here from their documentation/source as recalled
(module docstrings say which function follows which upstream routine).  What *does* pin this
oracle is optimiser-independent: closed-form MAPs / scores / J / H of the registered
families, central-difference checks of every analytic gradient, and the
reference's own statistical acceptance bound replayed over seeds (tests/test_oracle_*.py).
Two known answers of the UPSTREAM packages themselves are reproduced as well (quoted from memory of their documentation:
nothing is fetchable here): the L-BFGS run printed in Optim's manual — Rosenbrock from (0, 0), finite-difference gradient:
24 iterations, 67 f and ∇f calls, final objective 5.3784…e-17 — and FiniteDifferences'
``estimate_step(central_fdm(5, 1), sin, 1.0) = (0.001065235154086019, 1.9541865128909085e-13)``, to the last digit.

Julia's RNG bit streams cannot be reproduced; "identical draws" therefore means identical
base normals (ξ_k, ν_k) fed to both this oracle and the CUDA backend (common random numbers:
split_rng never advances the master, src/util.jl:87-92, so sim k sees the same base normals
at every θ).
"""

from .families import Funnel, HierGauss, CorrGauss, TwoLayer, TransformedFamily, make_family  # noqa: F401
from .hagerzhang import HagerZhang, LineSearchException  # noqa: F401
from .lbfgs import lbfgs_minimize, OptimResult  # noqa: F401
from .muse import (  # noqa: F401
    Draws, OracleProblem, MuseResult, muse, muse_bang, get_J_bang, get_H_bang, central_fdm, AdaptedFDM, pjacobian_adaptive, SimpleCovariance, cg, implicit_diff_H,
    finalize_result_bang, map_score_unit, NormalPrior, FlatPrior,
)
from .philox import philox_normals, philox4x32_10  # noqa: F401
