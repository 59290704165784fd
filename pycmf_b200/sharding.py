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

    def all_gather_rows(self, t, n_total, counts=None):
        return t

    def reduce_scatter_rows(self, t, out=None):
        """Sum over ranks of a (rows x k) tensor whose rows split evenly; returns this rank's row block of the sum
        (written into `out` when given)."""
        if out is not None:
            out.copy_(t)
            return out
        return t

    def all_gather_into(self, out, block):
        """out (rows x k, rows split evenly over the ranks) <- concatenation of every rank's block."""
        if out.data_ptr() != block.data_ptr():
            out.copy_(block)
        return out

    def all_to_all_chunks(self, t, send_counts, recv_counts):
        """t = the chunks for rank 0, 1, ... back to back (send_counts[g] leading-dimension entries for rank g); returns
        the chunks received from rank 0, 1, ... back to back (recv_counts[g] entries from rank g)."""
        return t

    def all_gather_ints(self, values, device=None):
        """(world x len(values)) nested list: every rank's host integers (exchanged through a tensor on `device`)."""
        return [[int(v) for v in values]]

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
        # NCCL collectives can be captured into the CUDA graph of one solver iteration
        import os
        self.graph_capturable = (dist.get_backend(group) == "nccl" and
                                 os.environ.get("PYCMF_B200_GRAPH_COLLECTIVES", "1") != "0")
        # collectives can run on a communication stream next to the compute stream (MUSolver._step_v_overlapped)
        self.overlap_capable = (dist.get_backend(group) == "nccl" and
                                os.environ.get("PYCMF_B200_OVERLAP", "1") != "0")

    def all_reduce_sum(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather_rows(self, t, n_total, counts=None):
        """Concatenate the row blocks of every rank (blocks follow row_range, or have `counts[r]` rows when the caller
        handed in its own blocks: sharded_input)."""
        if self.world == 1:
            return t
        import torch
        k = t.shape[1]
        if counts is None:
            sizes = [row_range(n_total, r, self.world) for r in range(self.world)]
        else:
            ends = np.cumsum([int(c) for c in counts])
            sizes = [(int(e - c), int(e)) for e, c in zip(ends, counts)]
        if len({b - a for a, b in sizes}) == 1 and self.dist.get_backend(self.group) == "nccl":
            # equal blocks: one all-gather straight into the result (the call all_gather_into makes), no padding, no list
            out = torch.empty(n_total, k, dtype=t.dtype, device=t.device)
            self.dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
            return out
        maxr = max(b - a for a, b in sizes)
        pad = torch.zeros(maxr, k, dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        out = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(out, pad, group=self.group)
        return torch.cat([o[:b - a] for o, (a, b) in zip(out, sizes)], 0)

    def reduce_scatter_rows(self, t, out=None):
        """Row block `rank` of the sum over ranks of t (rows x k, rows % world == 0), written into `out` when given.
        NCCL: one reduce-scatter (half the traffic of an all-reduce); other backends (gloo in the CPU tests):
        all-reduce + slice."""
        if self.world == 1:
            return Comm.reduce_scatter_rows(self, t, out)
        import torch
        rows = t.shape[0] // self.world
        assert rows * self.world == t.shape[0]
        if self.dist.get_backend(self.group) == "nccl":
            if out is None:
                out = torch.empty((rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            self.dist.reduce_scatter_tensor(out, t.contiguous(), op=self.dist.ReduceOp.SUM, group=self.group)
            return out
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        mine = t[self.rank * rows:(self.rank + 1) * rows]
        if out is not None:
            out.copy_(mine)
            return out
        return mine

    def all_gather_into(self, out, block):
        if self.world == 1:
            return Comm.all_gather_into(self, out, block)
        if self.dist.get_backend(self.group) == "nccl":
            self.dist.all_gather_into_tensor(out, block.contiguous(), group=self.group)
            return out
        parts = [block.new_empty(block.shape) for _ in range(self.world)]
        self.dist.all_gather(parts, block.contiguous(), group=self.group)
        rows = block.shape[0]
        for r, p in enumerate(parts):
            out[r * rows:(r + 1) * rows] = p
        return out

    def all_to_all_chunks(self, t, send_counts, recv_counts):
        if self.world == 1:
            return t
        import torch
        t = t.contiguous()
        out = torch.empty((int(sum(recv_counts)),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        self.dist.all_to_all_single(out, t, [int(c) for c in recv_counts], [int(c) for c in send_counts],
                                    group=self.group)
        return out

    def all_gather_ints(self, values, device=None):
        if self.world == 1:
            return Comm.all_gather_ints(self, values)
        import torch
        mine = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device)
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(parts, mine, group=self.group)
        return [[int(v) for v in p.tolist()] for p in parts]

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
