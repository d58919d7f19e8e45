"""Counter-based base normals (oracle side; TEST INFRASTRUCTURE ONLY).

The reference draws its simulations from child RNGs produced by ``split_rng``
(/root/reference/src/util.jl:85-92), whose defining property is that sim ``k`` sees the *same*
stream at every call (the master is never advanced).  Julia's Xoshiro bit streams cannot be
reproduced here, so the backend's throughput mode uses Philox4x32-10 (Salmon et al., SC'11)
keyed by ``(seed, global sim index, stream, element pair)`` and a Box–Muller transform.  This
module restates that generator in NumPy so tests can check the device draws element by
element and so the oracle can be fed the identical base normals.

Layout of one draw (must match museinference.jl_b200/csrc/muse_rng.cuh):
    counter = (pair index p, global sim index G, stream t, 0)   t = 0 → ξ (latent), 1 → ν (noise)
    key     = (seed & 0xffffffff, seed >> 32)
    (r0, r1, r2, r3) = philox4x32_10(counter, key)
    u1 = ((r1 >> 5)·2^26 + (r0 >> 6) + 0.5)·2^-53,  u2 likewise from (r3, r2)
    element 2p   = sqrt(-2 ln u1)·cos(2π u2)
    element 2p+1 = sqrt(-2 ln u1)·sin(2π u2)
The master stream (reference: draws taken from ``copy(rng)`` itself, src/muse.jl:151, 418)
uses G = 0xFFFFFFFF.
"""
from __future__ import annotations

import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
MASTER_INDEX = 0xFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; all inputs broadcastable uint32 arrays / ints."""
    c0 = np.asarray(c0, dtype=np.uint64) & _MASK
    c1 = np.asarray(c1, dtype=np.uint64) & _MASK
    c2 = np.asarray(c2, dtype=np.uint64) & _MASK
    c3 = np.asarray(c3, dtype=np.uint64) & _MASK
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def _u53(lo, hi):
    lo = lo.astype(np.uint64)
    hi = hi.astype(np.uint64)
    v = ((hi >> np.uint64(5)) << np.uint64(26)) + (lo >> np.uint64(6))
    return (v.astype(np.float64) + 0.5) * (2.0 ** -53)


def philox_normals(seed: int, sim_index: int, stream: int, d: int) -> np.ndarray:
    """``d`` standard normals of sim ``sim_index`` (global), stream 0 (ξ) or 1 (ν)."""
    npairs = (d + 1) // 2
    p = np.arange(npairs, dtype=np.uint64)
    r0, r1, r2, r3 = philox4x32_10(p, sim_index, stream, 0, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u1 = _u53(r0, r1)
    u2 = _u53(r2, r3)
    rad = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * np.pi * u2
    out = np.empty(2 * npairs, dtype=np.float64)
    out[0::2] = rad * np.cos(ang)
    out[1::2] = rad * np.sin(ang)
    return out[:d]
