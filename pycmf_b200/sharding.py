"""Row sharding of X / U across the GPUs of one box and the collectives the fit loop needs.

One process per GPU.  Rank g owns rows [row_range(n, g, G)) of X and U; V, Z and Y are replicated.
The only per-iteration exchange is an all-reduce (sum) of the V-side partial products
([X^T U ; U^T U] for MU, the gradient / Hessian partials for Newton) and of the objective partial
(SURVEY 8e).  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the plumbing.
"""
import numpy as np


def row_range(n, rank, world):
    """Contiguous, balanced block of rows owned by `rank`."""
    base, rem = divmod(int(n), int(world))
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


class Comm:
    """Single-process communicator: every collective is the identity."""
    rank = 0
    world = 1

    def all_reduce_sum(self, t):
        return t

    def all_gather_rows(self, t, n_total):
        return t

    def barrier(self):
        pass


class TorchComm(Comm):
    """torch.distributed-backed communicator (process group must already be initialised)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_reduce_sum(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather_rows(self, t, n_total):
        """Concatenate the row blocks of every rank (blocks follow row_range)."""
        if self.world == 1:
            return t
        import torch
        k = t.shape[1]
        sizes = [row_range(n_total, r, self.world) for r in range(self.world)]
        maxr = max(b - a for a, b in sizes)
        pad = torch.zeros(maxr, k, dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        out = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(out, pad, group=self.group)
        return torch.cat([o[:b - a] for o, (a, b) in zip(out, sizes)], 0)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.group)


def default_comm():
    """TorchComm if a process group is initialised, else the single-process Comm."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return TorchComm()
    except Exception:
        pass
    return Comm()


def localize_indices(idx, r0, r1):
    """Global sample indices -> shard-local ones; entries outside [r0, r1) become -1 (skipped by the kernels)."""
    idx = np.asarray(idx)
    inside = (idx >= r0) & (idx < r1)
    return np.where(inside, idx - r0, -1).astype(np.int32)
