"""Sharding of simulations over ranks (one process per GPU) and the one exchange step.

The reference parallelises with ``pmap(pool, ...)`` over independent simulations and lets the
master reduce the gathered per-sim scores (/root/reference/src/muse.jl:169, 183, 188, 426, 446,
508, 529; "trivially parallelizable", docs/src/userapi.md:81-86).  Here the ``pool`` is a
``ShardPool``: rank r owns a contiguous block of the global simulation index space, solves it on
its GPU, and the ranks all-gather the N×nθ score matrix (≤ 64 KB; latency-bound, NCCL over
NVLink when the process group is NCCL, gloo in CPU tests).  Every rank then runs the identical,
deterministic O(nθ²) outer-solver arithmetic, so θ stays bit-identical across ranks without a
broadcast.  No other collective exists on this path.
"""
from __future__ import annotations

import numpy as np


def block_partition(n: int, world: int):
    """Contiguous blocks: returns (offsets, counts) with the remainder on the first ranks."""
    base, rem = divmod(int(n), int(world))
    counts = [base + (1 if r < rem else 0) for r in range(world)]
    offsets = [0]
    for c in counts[:-1]:
        offsets.append(offsets[-1] + c)
    return offsets, counts


class LocalPool:
    """Single process, single GPU (the analogue of the reference's ``LocalWorkerPool``,
    src/util.jl:73-76)."""

    rank = 0
    world = 1
    device = 0

    def shard(self, n):
        return 0, int(n)

    def allgather_rows(self, local: np.ndarray, n_total: int) -> np.ndarray:
        return np.asarray(local)

    def barrier(self):
        pass


class ShardPool:
    """``torch.distributed`` process group: one rank per GPU."""

    def __init__(self, group=None, device: int | None = None):
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("ShardPool needs an initialised torch.distributed process group")
        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self.device = device if device is not None else 0

    def shard(self, n):
        offs, cnts = block_partition(n, self.world)
        return offs[self.rank], cnts[self.rank]

    def allgather_rows(self, local: np.ndarray, n_total: int) -> np.ndarray:
        """Concatenate per-rank row blocks (block_partition order) into the global matrix."""
        import torch

        local = np.ascontiguousarray(local, dtype=np.float64)
        ncol = local.shape[1] if local.ndim == 2 else 1
        _, cnts = block_partition(n_total, self.world)
        if local.shape[0] != cnts[self.rank]:
            raise ValueError("local block does not match the partition")
        maxc = max(cnts) if cnts else 0
        dev = torch.device("cuda", self.device) if self.backend == "nccl" else torch.device("cpu")
        buf = torch.zeros((maxc, ncol), dtype=torch.float64, device=dev)
        if local.shape[0]:
            buf[: local.shape[0]] = torch.from_numpy(local.reshape(local.shape[0], ncol)).to(dev)
        out = torch.empty((self.world, maxc, ncol), dtype=torch.float64, device=dev)
        self._dist.all_gather_into_tensor(out.view(-1), buf.view(-1), group=self.group)
        out = out.cpu().numpy()
        parts = [out[r, : cnts[r]] for r in range(self.world)]
        res = np.concatenate(parts, axis=0) if parts else np.zeros((0, ncol))
        return res if local.ndim == 2 else res[:, 0]

    def barrier(self):
        self._dist.barrier(group=self.group)
