"""Problem types of the B200 backend — the host-side mirror of the reference's problem interface.

``AbstractMuseProblem`` (/root/reference/src/interface.jl:4) is the reference's extension point;
``SimpleMuseProblem`` (src/simple.jl:4-12, 79-89) is the concrete type behind the registered
families.  The reference stores Julia closures (``sample_x_z``, ``logLike``, AD gradients); a
GPU backend cannot introspect closures, so the model is *named*: one of the registered families
whose sampling rule, log-density and analytic ∇z / ∇θ are compiled into the CUDA kernels
(csrc/muse_iso_solver.cu).  Anything else — in particular Turing/Soss-defined models
(src/turing.jl, src/soss.jl) — raises ``MuseBackendError`` (EUNSUPPORTED); there is no generic /
CPU path.

θ-transforms (``transform_θ`` / ``inv_transform_θ``, src/interface.jl:14-28) default to the identity, as for the
reference's ``SimpleMuseProblem``.  A problem whose θ has positive components (a standard deviation, a variance)
declares ``theta_transform=("identity", "log")``: the registered kernels are parameterised in the unconstrained
space, so the kernels' parameters ARE the transformed θ′ of src/muse.jl:136, their score is ∇θ′ logLike
(``Transformedθ()``, src/muse.jl:173) and the untransformed score (src/muse.jl:172, 432, 513) follows by the chain
rule, exactly as the reference's Soss adapter defines the pair (src/soss.jl:96-117).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from ._capi import MuseBackendError
from .backend import B200Backend, FAMILY_NTHETA


class AbstractMuseProblem:
    """src/interface.jl:4."""


# ----------------------------------------------------------------------------- priors (logPriorθ)
class FlatPrior:
    """``logPriorθ(θ) = 0`` — the default (src/interface.jl:121, src/simple.jl:79)."""

    def logp(self, theta):
        return 0.0

    def grad(self, theta):
        return np.zeros(np.size(theta))

    def hess(self, theta):
        return np.zeros((np.size(theta), np.size(theta)))


class NormalPrior:
    """Independent Normal(mean, sigma) per θ-component, e.g. the funnel's ``-θ^2/(2*3^2)``
    (src/simple.jl:69-71).  Gradient and Hessian are analytic (the reference uses ForwardDiff,
    src/muse.jl:184, 207, 539)."""

    def __init__(self, mean=0.0, sigma=3.0):
        self.mean, self.sigma = mean, sigma

    def _ms(self, theta):
        t = np.atleast_1d(np.asarray(theta, dtype=np.float64))
        return t, np.broadcast_to(np.asarray(self.mean, dtype=np.float64), t.shape), \
            np.broadcast_to(np.asarray(self.sigma, dtype=np.float64), t.shape)

    def logp(self, theta):
        t, m, s = self._ms(theta)
        return float(-np.sum((t - m) ** 2 / (2.0 * s ** 2)))

    def grad(self, theta):
        t, m, s = self._ms(theta)
        return -(t - m) / s ** 2

    def hess(self, theta):
        t, m, s = self._ms(theta)
        return np.diag(-1.0 / s ** 2)


# ----------------------------------------------------------------------------- base normals
@dataclass
class BaseDraws:
    """Explicit base normals (parity mode): row k of ``xi`` / ``nu`` are the latent / noise
    normals of child stream k; ``xi_master`` / ``nu_master`` the master stream's own draw.
    Stands in for the ``rng`` keyword (src/muse.jl:116, 134; split semantics src/util.jl:85-92)."""

    xi: np.ndarray
    nu: np.ndarray
    xi_master: np.ndarray
    nu_master: np.ndarray


# ----------------------------------------------------------------------------- problem
class SimpleMuseProblem(AbstractMuseProblem):
    """``SimpleMuseProblem(x, family; logPriorθ)`` for a registered family.

    Parameters
    ----------
    x        observed data, length d            (``prob.x``, src/simple.jl:5); "twolayer": the stacked (x, y), d = 2n
    family   "funnel" | "hiergauss" | "corrgauss" | "twolayer" (the toy hierarchy of src/turing.jl:63-79, parameter σ)
    prior    object with ``logp/grad/hess`` (``logPriorθ``); default flat
    group, cluster   solver geometry overrides (0 = auto), see DESIGN.md §3
    stream   raw cudaStream_t to launch on (e.g. ``torch.cuda.current_stream().cuda_stream``)
    """

    def __init__(self, x, family: str = "funnel", prior=None, *, theta_transform=None, P=None, L=None, group: int = 0,
                 cluster: int = 0, kernel: int = 0, stream=None, backend_factory=None):
        if family not in FAMILY_NTHETA:
            raise MuseBackendError(-5, f"model family {family!r} is not registered with the B200 backend; "
                                       "Turing/Soss-defined models are not supported and there is no CPU fallback")
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        if self.x.ndim != 1:
            raise ValueError("x must be a vector")
        self.family = family
        self.d = self.x.size
        self.ntheta = FAMILY_NTHETA[family]
        self.prior = prior or FlatPrior()
        kinds = tuple(theta_transform) if theta_transform is not None else ("identity",) * self.ntheta
        if len(kinds) != self.ntheta or any(k not in ("identity", "log") for k in kinds):
            raise ValueError(f"theta_transform must name {self.ntheta} component transform(s) out of 'identity' | 'log'")
        self.theta_transform = kinds
        self._log = np.array([k == "log" for k in kinds])
        self.has_transform = bool(self._log.any())
        self.P, self.L = P, L
        self.group, self.cluster, self.kernel = group, cluster, kernel
        self._backend_factory = backend_factory or B200Backend
        self._backend = None
        self._backend_key = None
        self._rng_key = None
        self._data_dirty = True
        self.stream = stream

    # src/interface.jl:134
    def standardize_theta(self, theta):
        t = np.atleast_1d(np.asarray(theta, dtype=np.float64)).copy()
        if t.shape != (self.ntheta,):
            raise ValueError(f"θ must have {self.ntheta} component(s) for family {self.family!r}")
        return t

    def logPrior(self, theta):
        return self.prior.logp(theta)

    # ------------------------------------------------------------------ θ-transforms (src/interface.jl:14-28)
    def transform_theta(self, theta):
        """θ → θ′ ∈ (−∞, ∞)ⁿ: ``log`` on the components declared positive.  θ′ is what the kernels take."""
        t = np.array(theta, dtype=np.float64, copy=True)
        if self.has_transform:
            if np.any(t[self._log] <= 0):
                raise ValueError("DomainError: a log-transformed θ component must be positive")
            t[self._log] = np.log(t[self._log])
        return t

    def inv_transform_theta(self, theta_t):
        t = np.array(theta_t, dtype=np.float64, copy=True)
        if self.has_transform:
            t[self._log] = np.exp(t[self._log])
        return t

    def dinv_transform(self, theta_t):
        """Diagonal of ∂θ/∂θ′ at θ′ (1, or θ for a log component): ∇θ′ = ∂θ/∂θ′ ⊙ ∇θ."""
        j = np.ones(self.ntheta)
        if self.has_transform:
            j[self._log] = np.exp(np.asarray(theta_t, dtype=np.float64)[self._log])
        return j

    def prior_grad_t(self, theta_t):
        """∇θ′ logPriorθ(θ′, Transformedθ()), with logPriorθ(θ′, Transformedθ()) = logPriorθ(inv_transform_θ(θ′))
        (src/soss.jl:107-108; used at src/muse.jl:184)."""
        th = self.inv_transform_theta(theta_t)
        return self.dinv_transform(theta_t) * np.asarray(self.prior.grad(th), dtype=np.float64)

    def prior_hess_t(self, theta_t):
        """∇²θ′ of the same function (src/muse.jl:207): D·∇²logπ·D + diag(∇logπ ⊙ ∂²θ/∂θ′²), D = diag(∂θ/∂θ′)."""
        th = self.inv_transform_theta(theta_t)
        D = self.dinv_transform(theta_t)
        H = np.asarray(self.prior.hess(th), dtype=np.float64) * np.outer(D, D)
        if self.has_transform:
            g = np.asarray(self.prior.grad(th), dtype=np.float64)
            H = H + np.diag(np.where(self._log, g * D, 0.0))      # ∂²e^{θ′}/∂θ′² = e^{θ′} = D
        return H

    # ------------------------------------------------------------------ backend management
    def set_data(self, x):
        """Replace the observed data (host buffer); uploaded on the next backend use."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.shape != (self.d,):
            raise ValueError("x has the wrong length")
        self.x = x
        self._data_dirty = True

    def backend_for(self, nsims_total: int, rng, pool, nsims_h_total: int = 0):
        """Handle holding this rank's shard of ``nsims_total`` sims (and of the first
        ``nsims_h_total`` sims for get_H!) with the draws of ``rng`` installed.  The handle (device
        memory) is reused while the shard geometry is unchanged; a new ``rng`` only re-installs draws."""
        off, cnt = pool.shard(nsims_total)
        hoff, hcnt = pool.shard(nsims_h_total) if (pool.world > 1 and nsims_h_total > 0) else (0, 0)
        gkey = (nsims_total, off, cnt, hoff, hcnt, pool.device)
        # installed draws are recognised by identity: the BaseDraws object itself is held (a bare id() could be reused by a
        # new object once the old one is collected, and stale draws would pass for the new ones).  Arrays mutated in place
        # are NOT detected — pass a new BaseDraws for new normals.
        rkey = ("seed", int(rng)) if isinstance(rng, (int, np.integer)) else ("draws", rng)
        if self._backend is None or self._backend_key != gkey:
            if self._backend is not None:
                self._backend.close()
            self._backend = self._backend_factory(
                self.family, self.d, cnt, sim_offset=off, nsims_h=hcnt, h_sim_offset=hoff, device=pool.device,
                group=self.group, cluster=self.cluster, kernel=self.kernel, stream=self.stream, P=self.P, L=self.L)
            self._backend_key, self._rng_key, self._data_dirty = gkey, None, True
        be = self._backend
        if self._data_dirty:
            be.set_data(self.x)
            self._data_dirty = False
        same = (self._rng_key is not None and self._rng_key[0] == rkey[0]
                and (self._rng_key[1] is rkey[1] if rkey[0] == "draws" else self._rng_key[1] == rkey[1]))
        if not same:
            if isinstance(rng, (int, np.integer)):
                be.seed_draws(int(rng))
            elif isinstance(rng, BaseDraws):
                if rng.xi.shape[0] < nsims_total:
                    raise ValueError("BaseDraws holds fewer simulations than requested")
                be.set_draws(rng.xi[off:off + cnt], rng.nu[off:off + cnt], rng.xi_master, rng.nu_master)
                if hcnt:
                    be.set_draws_h(rng.xi[hoff:hoff + hcnt], rng.nu[hoff:hoff + hcnt])
            else:
                raise TypeError("rng must be an integer seed or a BaseDraws")
            self._rng_key = rkey
        return be

    def close(self):
        if self._backend is not None:
            self._backend.close()
            self._backend = None
            self._backend_key = None
