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

    def any_flag(self, flag: bool) -> bool:
        return bool(flag)

    def barrier(self):
        pass


class ShardPool:
    """``torch.distributed`` process group: one rank per GPU.

    The exchange step is an all-gather of per-sim score rows.  With the NCCL backend it runs inside the library
    (csrc/muse_comm.cu): ncclAllGather straight from the device output buffer on the stream the solver kernels were
    launched on, the process group only distributes the communicator id.  With gloo (CPU tests) the rows go
    through ``torch.distributed`` on the host."""

    def __init__(self, group=None, device: int | None = None):
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("ShardPool needs an initialised torch.distributed process group")
        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        if device is None:
            # one rank per GPU: torchrun's LOCAL_RANK, else the rank modulo the visible devices (never "everyone on GPU 0")
            import os
            device = int(os.environ.get("LOCAL_RANK", -1))
            if device < 0:
                try:
                    import torch
                    device = self.rank % max(1, torch.cuda.device_count())
                except Exception:
                    device = 0
        self.device = device
        self._bufs = {}

    def shard(self, n):
        offs, cnts = block_partition(n, self.world)
        return offs[self.rank], cnts[self.rank]

    # ------------------------------------------------------------------ staging buffers
    def _staging(self, maxc: int, ncol: int):
        import torch

        key = (maxc, ncol)
        b = self._bufs.get(key)
        if b is None:
            cuda = self.backend == "nccl"
            dev = torch.device("cuda", self.device) if cuda else torch.device("cpu")
            send_h = torch.zeros((maxc, ncol), dtype=torch.float64)
            recv_h = torch.zeros((self.world, maxc, ncol), dtype=torch.float64)
            if cuda:
                send_h, recv_h = send_h.pin_memory(), recv_h.pin_memory()
            b = dict(send_h=send_h, recv_h=recv_h,
                     send_d=torch.zeros((maxc, ncol), dtype=torch.float64, device=dev) if cuda else send_h,
                     recv_d=torch.zeros((self.world, maxc, ncol), dtype=torch.float64, device=dev) if cuda else recv_h)
            self._bufs[key] = b
        return b

    def _finish(self, b, cnts, ncol):
        import torch

        self._dist.all_gather_into_tensor(b["recv_d"].view(-1), b["send_d"].view(-1), group=self.group)
        if self.backend == "nccl":
            b["recv_h"].copy_(b["recv_d"], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        out = b["recv_h"].numpy()
        parts = [out[r, : cnts[r]] for r in range(self.world)]
        return np.concatenate(parts, axis=0) if parts else np.zeros((0, ncol))

    # ------------------------------------------------------------------ the exchange step
    def allgather_rows(self, local: np.ndarray, n_total: int) -> np.ndarray:
        """Concatenate per-rank row blocks (block_partition order) into the global matrix (host rows in)."""
        import torch

        local = np.ascontiguousarray(local, dtype=np.float64)
        ncol = local.shape[1] if local.ndim == 2 else 1
        _, cnts = block_partition(n_total, self.world)
        if local.shape[0] != cnts[self.rank]:
            raise ValueError("local block does not match the partition")
        b = self._staging(max(max(cnts), 1), ncol)
        n = local.shape[0]
        if n:
            b["send_h"][:n] = torch.from_numpy(local.reshape(n, ncol))
        if self.backend == "nccl":
            b["send_d"].copy_(b["send_h"], non_blocking=True)
        res = self._finish(b, cnts, ncol)
        return res if local.ndim == 2 else res[:, 0]

    # ------------------------------------------------------------------ native exchange (NCCL inside the library)
    def uses_device_gather(self) -> bool:
        return self.backend == "nccl"

    def bind(self, be, nsims: int | None = None, nh_total: int = 0):
        """Give the handle a communicator over this group's ranks (once per handle) and — when the problem size is known
        — the peer-mapped exchange buffers of the one-launch solve (csrc/muse_comm.cu: muse_b200_p2p_*): every rank allocates
        its region, the 64-byte IPC handles go round the process group, every rank maps its peers' regions.  Collective: all
        ranks call it with the same arguments.  MUSE_EXCHANGE=nccl keeps the NCCL all-gather only."""
        import os

        if getattr(be, "comm", None) is None:
            ids = [be.comm_unique_id() if self.rank == 0 else None]
            self._dist.broadcast_object_list(ids, src=self._dist.get_global_rank(self.group, 0) if self.group else 0,
                                             group=self.group)
            be.comm_init(self.world, self.rank, ids[0])
        if nsims is None or not hasattr(be, "p2p_alloc") or os.environ.get("MUSE_EXCHANGE", "auto") == "nccl" or self.world > 16:
            return
        nt = be.ntheta
        maxc = max(block_partition(nsims, self.world)[1])
        maxh = max(block_partition(max(nh_total, 0), self.world)[1]) if nh_total else 0
        need = self.world * max(maxc * nt, maxh * 2 * nt * nt, 1)
        if getattr(be, "_p2p_block", 0) >= need or getattr(be, "_p2p_failed", False):
            return
        # every rank takes the same decisions: the outcome of each step is agreed on before the next
        handle, ok = None, True
        try:
            handle = be.p2p_alloc(self.world, self.rank, need)
        except Exception:
            ok = False
        got = [None] * self.world
        self._dist.all_gather_object(got, handle if ok else None, group=self.group)
        if any(g is None for g in got):
            be._p2p_failed = True
            return
        try:
            be.p2p_connect(got)
        except Exception:
            ok = False
        flags = [None] * self.world
        self._dist.all_gather_object(flags, ok, group=self.group)
        if not all(flags):
            be._p2p_failed = True       # NOTE: ranks that did connect keep p2p_ready; block size 0 on the others stops them using it
            be._p2p_block = 0
            try:
                be.p2p_alloc(self.world, self.rank, 1)      # drop the mapping state: a 1-double region can never fit a solve
            except Exception:
                pass
            return
        be._p2p_block = need

    def allgather_device_scores(self, be, first_row: int, n_total: int) -> np.ndarray:
        """All-gather this rank's sim rows [first_row, first_row + count) of the score matrix the last
        ``map_score_async`` left on the device (NCCL on the launch stream, inside the library); returns the
        global N × nθ matrix on the host."""
        _, cnts = block_partition(n_total, self.world)
        self.bind(be)
        return be.allgather_scores(first_row, cnts)

    def allgather_host_rows(self, be, local: np.ndarray, n_total: int) -> np.ndarray:
        _, cnts = block_partition(n_total, self.world)
        if local.shape[0] != cnts[self.rank]:
            raise ValueError("local block does not match the partition")
        self.bind(be)
        return be.allgather_rows(local, cnts)

    def any_flag(self, flag: bool) -> bool:
        """True on every rank iff any rank passed True (the collective form of a per-rank error test)."""
        rows = self.allgather_rows(np.array([[1.0 if flag else 0.0]]), self.world)
        return bool(rows.any())

    def barrier(self):
        self._dist.barrier(group=self.group)
