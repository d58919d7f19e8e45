"""Registered model families (oracle side; TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Each family supplies what a ``SimpleMuseProblem`` holds as closures
(/root/reference/src/simple.jl:4-12, 79-95): ``sample_x_z``, ``logLike``,
``∇θ_logLike`` and ``logLike_and_∇z_logLike``.  The reference obtains the two
gradients by automatic differentiation of ``logLike`` (src/simple.jl:84-85); here they
are the analytic derivatives of the same expression, checked against central
differences in tests/test_oracle_families.py.

F1  Neal's funnel             /root/reference/src/simple.jl:58-76, docs/src/index.md:154-168
F2  hierarchical Gaussian     SURVEY.md §8(a) row F2 (BASELINE.json config 4; not in the reference)
F3  dense correlated Gaussian SURVEY.md §8(a) row F3 (BASELINE.json config 5; not in the reference)
F4  two-layer hierarchy       /root/reference/src/turing.jl:63-79 (the toy model of the Turing adapter's docstring)

Sampling is expressed on *base normals* (ξ, ν) so that common-random-number semantics
(src/util.jl:85-92) are explicit: the latent draws come first, then the noise draws
(src/simple.jl:62-63).

All functions minimise  f(z) = -logLike(x, z, θ)  exactly as the default ``ẑ_at_θ`` does
(src/interface.jl:163: ``z -> .-logLike_and_∇z_logLike(prob, x, z, θ)``).
"""
from __future__ import annotations

import numpy as np


class Funnel:
    """F1: z ~ N(0, e^θ I_d), x ~ N(z, I_d); scalar θ."""

    name = "funnel"
    family_id = 1
    ntheta = 1

    def __init__(self, d: int):
        self.d = int(d)

    # src/simple.jl:61-65
    def sample(self, theta, xi, nu):
        th = float(np.asarray(theta).reshape(-1)[0])
        z = np.exp(0.5 * th) * xi
        x = z + nu
        return x, z

    # src/simple.jl:66-68:  -(1//2) * (sum((x .- z).^2) + sum(z.^2) / exp(θ) + d*θ)
    def neg_loglike(self, x, z, theta):
        th = float(np.asarray(theta).reshape(-1)[0])
        r = x - z
        return 0.5 * (np.dot(r, r) + np.dot(z, z) / np.exp(th) + self.d * th)

    def neg_loglike_and_grad(self, x, z, theta):
        th = float(np.asarray(theta).reshape(-1)[0])
        a = np.exp(-th)
        r = x - z
        f = 0.5 * (np.dot(r, r) + a * np.dot(z, z) + self.d * th)
        g = -r + a * z
        return f, g

    # ∇θ logLike = ½ e^{-θ} Σ z² − d/2
    def score(self, x, z, theta):
        th = float(np.asarray(theta).reshape(-1)[0])
        return np.array([0.5 * np.exp(-th) * np.dot(z, z) - 0.5 * self.d])

    def exact_map(self, x, theta):
        th = float(np.asarray(theta).reshape(-1)[0])
        return x / (1.0 + np.exp(-th))

    # second derivatives of logLike the implicit-diff branch of get_H! obtains by nested AD (src/muse.jl:349-375), analytic here:
    # ∂θ ∇z logLike (d × nθ), ∂x/∂θ_sim of the sample (d × nθ; ∇z logLike depends on θ_sim only through x, with ∂∇z/∂x = I),
    # the Hessian ∇²z logLike applied to a vector, and ∂θ_sim of the score at fixed ẑ (these families: the score does not see x → 0)
    def dgradz_dtheta(self, x, z, theta):
        th = float(np.asarray(theta).reshape(-1)[0])
        return (np.exp(-th) * z)[:, None]

    def dx_dtheta_sim(self, theta, xi, nu):
        th = float(np.asarray(theta).reshape(-1)[0])
        return (0.5 * np.exp(0.5 * th) * xi)[:, None]

    def hess_z_apply(self, z, theta, w):
        th = float(np.asarray(theta).reshape(-1)[0])
        return -(1.0 + np.exp(-th)) * w


class HierGauss:
    """F2: z ~ N(μ, e^{2ℓ} I_d), x ~ N(z, I_d); θ = (μ, ℓ = log σ)."""

    name = "hiergauss"
    family_id = 2
    ntheta = 2

    def __init__(self, d: int):
        self.d = int(d)

    def sample(self, theta, xi, nu):
        mu, ell = (float(t) for t in np.asarray(theta).reshape(-1))
        z = mu + np.exp(ell) * xi
        x = z + nu
        return x, z

    def neg_loglike(self, x, z, theta):
        return self.neg_loglike_and_grad(x, z, theta)[0]

    def neg_loglike_and_grad(self, x, z, theta):
        mu, ell = (float(t) for t in np.asarray(theta).reshape(-1))
        a = np.exp(-2.0 * ell)
        r = x - z
        w = z - mu
        f = 0.5 * (np.dot(r, r) + a * np.dot(w, w) + 2.0 * self.d * ell)
        g = -r + a * w
        return f, g

    def score(self, x, z, theta):
        mu, ell = (float(t) for t in np.asarray(theta).reshape(-1))
        a = np.exp(-2.0 * ell)
        w = z - mu
        return np.array([a * np.sum(w), a * np.dot(w, w) - self.d])

    def exact_map(self, x, theta):
        mu, ell = (float(t) for t in np.asarray(theta).reshape(-1))
        a = np.exp(-2.0 * ell)
        return (x + mu * a) / (1.0 + a)

    def dgradz_dtheta(self, x, z, theta):          # ∇z logLike = (x − z) − a(z − μ):  ∂μ → a,  ∂ℓ → 2a(z − μ)
        mu, ell = (float(t) for t in np.asarray(theta).reshape(-1))
        a = np.exp(-2.0 * ell)
        return np.stack([np.full(self.d, a), 2.0 * a * (z - mu)], axis=1)

    def dx_dtheta_sim(self, theta, xi, nu):        # x = μ + e^ℓ ξ + ν
        mu, ell = (float(t) for t in np.asarray(theta).reshape(-1))
        return np.stack([np.ones(self.d), np.exp(ell) * xi], axis=1)

    def hess_z_apply(self, z, theta, w):
        mu, ell = (float(t) for t in np.asarray(theta).reshape(-1))
        return -(1.0 + np.exp(-2.0 * ell)) * w


class CorrGauss:
    """F3: z ~ N(0, e^θ Σ₀), x ~ N(z, I_d); scalar θ; P = Σ₀⁻¹ and L = chol(Σ₀) are inputs."""

    name = "corrgauss"
    family_id = 3
    ntheta = 1

    def __init__(self, d: int, P: np.ndarray, L: np.ndarray):
        self.d = int(d)
        self.P = np.ascontiguousarray(P, dtype=np.float64)
        self.L = np.ascontiguousarray(L, dtype=np.float64)
        assert self.P.shape == (d, d) and self.L.shape == (d, d)

    def sample(self, theta, xi, nu):
        th = float(np.asarray(theta).reshape(-1)[0])
        z = np.exp(0.5 * th) * (self.L @ xi)
        x = z + nu
        return x, z

    def neg_loglike(self, x, z, theta):
        return self.neg_loglike_and_grad(x, z, theta)[0]

    def neg_loglike_and_grad(self, x, z, theta):
        th = float(np.asarray(theta).reshape(-1)[0])
        a = np.exp(-th)
        r = x - z
        Pz = self.P @ z
        f = 0.5 * (np.dot(r, r) + a * np.dot(z, Pz) + self.d * th)
        g = -r + a * Pz
        return f, g

    def score(self, x, z, theta):
        th = float(np.asarray(theta).reshape(-1)[0])
        return np.array([0.5 * np.exp(-th) * np.dot(z, self.P @ z) - 0.5 * self.d])

    def exact_map(self, x, theta):
        th = float(np.asarray(theta).reshape(-1)[0])
        A = np.eye(self.d) + np.exp(-th) * self.P
        return np.linalg.solve(A, x)

    def dgradz_dtheta(self, x, z, theta):          # ∇z logLike = (x − z) − a P z:  ∂θ → a P z
        th = float(np.asarray(theta).reshape(-1)[0])
        return (np.exp(-th) * (self.P @ z))[:, None]

    def dx_dtheta_sim(self, theta, xi, nu):        # x = e^{θ/2} L ξ + ν
        th = float(np.asarray(theta).reshape(-1)[0])
        return (0.5 * np.exp(0.5 * th) * (self.L @ xi))[:, None]

    def hess_z_apply(self, z, theta, w):
        th = float(np.asarray(theta).reshape(-1)[0])
        return -(w + np.exp(-th) * (self.P @ w))


class TwoLayer:
    """F4: the toy hierarchy of the Turing adapter's docstring (/root/reference/src/turing.jl:63-79)

        z ~ MvNormal(zeros(n), exp(σ/2)*I)      (covariance e^{σ/2} I: z = e^{σ/4} ξ_z)
        w ~ MvNormal(z, I);  x ~ MvNormal(w, I);  y ~ MvNormal(x, I)

    with σ the parameter, (x, y) the data and (z, w) the latent space, as in the docstring (`model | (;sim.x, sim.y)`).  The
    docstring's second parameter θ ~ Normal(0, σ) has no descendants: it enters the joint density only through its own prior
    factor, identical for the data and every simulation, so it cancels out of every MUSE quantity and is not carried here.
    Everything is stacked to one length d = 2n: latent u = (z, w), data (x, y), latent normals ξ = (ξ_z, ξ_w), noise normals
    ν = (ν_x, ν_y).  logLike follows the funnel's convention of dropping the 2π constants (src/simple.jl:66-68):

        logLike = −½ [ e^{−σ/2} Σz² + Σ(w − z)² + Σ(x − w)² + Σ(y − x)² + n σ/2 ]

    The Hessian in u is [[e^{−σ/2} + 1, −1], [−1, 2]] ⊗ I_n: two distinct eigenvalues, not a multiple of the identity — the
    default ẑ_at_θ needs several L-BFGS iterations with a live (dx, dg) history."""

    name = "twolayer"
    family_id = 4
    ntheta = 1

    def __init__(self, d: int):
        self.d = int(d)
        if self.d % 2:
            raise ValueError("twolayer: d = 2n (latent (z, w) and data (x, y) stacked) must be even")
        self.n = self.d // 2

    def sample(self, theta, xi, nu):
        s = float(np.asarray(theta).reshape(-1)[0])
        n = self.n
        z = np.exp(0.25 * s) * xi[:n]
        w = z + xi[n:]
        x = w + nu[:n]
        y = x + nu[n:]
        return np.concatenate([x, y]), np.concatenate([z, w])

    def neg_loglike(self, x, z, theta):
        return self.neg_loglike_and_grad(x, z, theta)[0]

    def neg_loglike_and_grad(self, x, z, theta):
        s = float(np.asarray(theta).reshape(-1)[0])
        n = self.n
        b = np.exp(-0.5 * s)
        zz, ww, xx, yy = z[:n], z[n:], x[:n], x[n:]
        r1, r2, r3 = ww - zz, xx - ww, yy - xx
        f = 0.5 * (b * np.dot(zz, zz) + np.dot(r1, r1) + np.dot(r2, r2) + np.dot(r3, r3) + 0.5 * n * s)
        g = np.concatenate([b * zz - r1, r1 - r2])
        return f, g

    # ∇σ logLike = ¼ e^{−σ/2} Σ z² − n/4
    def score(self, x, z, theta):
        s = float(np.asarray(theta).reshape(-1)[0])
        zz = z[:self.n]
        return np.array([0.25 * np.exp(-0.5 * s) * np.dot(zz, zz) - 0.25 * self.n])

    def exact_map(self, x, theta):
        s = float(np.asarray(theta).reshape(-1)[0])
        b = np.exp(-0.5 * s)
        xx = x[:self.n]
        return np.concatenate([xx / (2.0 * b + 1.0), (b + 1.0) * xx / (2.0 * b + 1.0)])

    def dgradz_dtheta(self, x, z, theta):          # ∇u logLike = (−b z + (w − z), −(w − z) + (x − w)):  ∂σ → (½ b z, 0)
        s = float(np.asarray(theta).reshape(-1)[0])
        out = np.zeros((self.d, 1))
        out[:self.n, 0] = 0.5 * np.exp(-0.5 * s) * z[:self.n]
        return out

    def dx_dtheta_sim(self, theta, xi, nu):
        """∂θ_sim of ∇u logLike at fixed u: only the w-half sees the data (coefficient 1 on x), x = e^{σ/4} ξ_z + ξ_w + ν_x."""
        s = float(np.asarray(theta).reshape(-1)[0])
        out = np.zeros((self.d, 1))
        out[self.n:, 0] = 0.25 * np.exp(0.25 * s) * xi[:self.n]
        return out

    def hess_z_apply(self, z, theta, w):
        s = float(np.asarray(theta).reshape(-1)[0])
        n = self.n
        b = np.exp(-0.5 * s)
        return -np.concatenate([(b + 1.0) * w[:n] - w[n:], 2.0 * w[n:] - w[:n]])


class TransformedFamily:
    """A registered family whose user-facing θ has positive components: θ_user[i] = exp(θ_base[i]) where
    ``kinds[i] == "log"``.  Supplies the pair the reference's interface asks of a problem with a bounded θ
    (/root/reference/src/interface.jl:14-28, 36-58): ``transform_theta`` / ``inv_transform_theta`` and the score in both
    spaces, defined as the reference's Soss adapter defines them (src/soss.jl:96-117): the transformed-space gradient
    is the gradient of θ′ ↦ logLike(inv_transform_θ(θ′)), i.e. the base family's own score."""

    def __init__(self, base, kinds):
        self.base = base
        self.kinds = tuple(kinds)
        assert len(self.kinds) == base.ntheta and all(k in ("identity", "log") for k in self.kinds)
        self._log = np.array([k == "log" for k in self.kinds])
        self.name = base.name + "[" + ",".join(self.kinds) + "]"
        self.family_id = base.family_id
        self.ntheta = base.ntheta
        self.d = base.d

    def transform_theta(self, theta):
        t = np.array(theta, dtype=np.float64, copy=True).reshape(-1)
        if np.any(t[self._log] <= 0):
            raise ValueError("DomainError: a log-transformed θ component must be positive")
        t[self._log] = np.log(t[self._log])
        return t

    def inv_transform_theta(self, theta_t):
        t = np.array(theta_t, dtype=np.float64, copy=True).reshape(-1)
        t[self._log] = np.exp(t[self._log])
        return t

    def dinv_transform(self, theta_t):
        j = np.ones(self.ntheta)
        j[self._log] = np.exp(np.asarray(theta_t, dtype=np.float64).reshape(-1)[self._log])
        return j

    # everything below takes the UNtransformed θ, like sample_x_z / logLike of the reference (src/interface.jl:62-99)
    def sample(self, theta, xi, nu):
        return self.base.sample(self.transform_theta(theta), xi, nu)

    def neg_loglike(self, x, z, theta):
        return self.base.neg_loglike(x, z, self.transform_theta(theta))

    def neg_loglike_and_grad(self, x, z, theta):
        return self.base.neg_loglike_and_grad(x, z, self.transform_theta(theta))

    def score(self, x, z, theta):
        """∇θ logLike(x, z, θ, UnTransformedθ())  (src/muse.jl:172, 432, 513)."""
        tt = self.transform_theta(theta)
        return self.base.score(x, z, tt) / self.dinv_transform(tt)

    def score_t(self, x, z, theta_t):
        """∇θ′ logLike(x, z, θ′, Transformedθ())  (src/muse.jl:173)."""
        return self.base.score(x, z, np.asarray(theta_t, dtype=np.float64).reshape(-1))

    def exact_map(self, x, theta):
        return self.base.exact_map(x, self.transform_theta(theta))


def make_family(name: str, d: int, **consts):
    if name == "funnel":
        return Funnel(d)
    if name == "hiergauss":
        return HierGauss(d)
    if name == "corrgauss":
        return CorrGauss(d, consts["P"], consts["L"])
    if name == "twolayer":
        return TwoLayer(d)
    raise ValueError(f"unknown family {name!r}")
