"""Shared test helpers: build matching (oracle, backend) problem pairs on identical base normals."""
import numpy as np

import oracle as O


def corr_consts(d, seed=5):
    """Σ₀ = A·Aᵀ/d + 0.1·I with A ~ N(0,1) from a stated seed; returns P = Σ₀⁻¹, L = chol Σ₀."""
    rng = np.random.Generator(np.random.Philox(seed))
    A = rng.standard_normal((d, d))
    S0 = A @ A.T / d + 0.1 * np.eye(d)
    return np.linalg.inv(S0), np.linalg.cholesky(S0)


def make_family(name, d):
    if name == "corrgauss":
        P, L = corr_consts(d)
        return O.make_family(name, d, P=P, L=L)
    return O.make_family(name, d)


def theta_true(name):
    if name == "twolayer":
        return np.array([0.3])                   # the docstring model draws σ ~ U(0, 1)
    return np.array([0.0, 0.0]) if name == "hiergauss" else np.array([0.0])


def theta_start(name):
    if name == "twolayer":
        return np.array([0.5])                   # muse(prob, (σ=0.5, θ=0)), src/turing.jl:78
    return np.array([0.5, 0.3]) if name == "hiergauss" else np.array([1.0])


def make_inputs(name, d, nsims, seed=1234, data_seed=99):
    """Philox base normals for nsims sims + master, and observed data drawn at θ_true."""
    fam = make_family(name, d)
    draws = O.Draws.from_philox(seed, nsims, d)
    xd, _ = fam.sample(theta_true(name), O.philox_normals(data_seed, 0, 0, d), O.philox_normals(data_seed, 0, 1, d))
    return fam, draws, xd


def oracle_problem(name, d, nsims, seed=1234, prior=None):
    fam, draws, xd = make_inputs(name, d, nsims, seed)
    return O.OracleProblem(fam, xd, draws, prior), fam, draws, xd
