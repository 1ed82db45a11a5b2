"""Multi-GPU: atoms are independent, so a trajectory shards by contiguous atom blocks
(SURVEY.md section 8e).  One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch);
every rank runs the jump scan and the correlation kernels on its own slab, and the small
per-(origin, lag) partial sums -- moments and histogram counts packed into ONE float64 buffer --
are summed with a single all-reduce.  The division by the GLOBAL atom count, ratios and the
origin statistics happen after the reduction, on every rank.
"""
from __future__ import annotations

import numpy as np


def atom_block(n_atoms, rank, world_size, align=4):
    """Contiguous block ``[a0, a1)`` of rank ``rank``; block starts are multiples of ``align`` atoms
    so every shard keeps 16-byte aligned rows."""
    per = -(-n_atoms // world_size)
    per = -(-per // align) * align
    a0 = min(n_atoms, rank * per)
    a1 = min(n_atoms, a0 + per)
    return a0, a1


def shard_atoms(coords, group=None, n_atoms_global=None):
    """Mark ``coords`` (this rank's slab of atoms) as one shard of a distributed trajectory.

    ``n_atoms_global`` defaults to the sum of the shard sizes over the group.
    """
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError('torch.distributed is not initialised')
    if n_atoms_global is None:
        local = int(np.count_nonzero(coords.atom_filter)) if coords.atom_filter.dtype == bool else len(
            coords.atom_filter)
        dev = 'cuda' if dist.get_backend(group) == 'nccl' else 'cpu'
        t = torch.tensor([local], dtype=torch.int64, device=dev)
        dist.all_reduce(t, group=group)
        n_atoms_global = int(t.item())
    coords.comm = (group, int(n_atoms_global))
    return coords


def pack_partials(sums, hist):
    """One flat float64 buffer ``[sums | hist]``; counts are exact in float64 up to 2^53."""
    import torch
    if hist is None:
        return sums.reshape(-1)
    return torch.cat([sums.reshape(-1), hist.reshape(-1).to(torch.float64)])


def unpack_partials(buf, sums_shape, hist_shape):
    import torch
    n = int(np.prod(sums_shape))
    sums = buf[:n].reshape(sums_shape)
    hist = None
    if hist_shape is not None:
        hist = buf[n:].round().to(torch.int64).reshape(hist_shape)
    return sums, hist


def allreduce_counts(counts, comm, device):
    """Sub-ensemble sizes per origin (host int array of any shape) summed over the shards."""
    import torch
    import torch.distributed as dist
    group, _ = comm
    t = torch.from_numpy(np.ascontiguousarray(counts, dtype=np.int64)).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy().reshape(np.shape(counts))


def allreduce_partials(sums, hist, n_atoms_local, comm):
    """Sum the per-shard partials over the group: ONE collective per correlation call."""
    import torch.distributed as dist
    group, n_global = comm
    buf = pack_partials(sums, hist)
    if hist is None and not buf.is_contiguous():
        buf = buf.contiguous()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    sums, hist = unpack_partials(buf, tuple(sums.shape), None if hist is None else tuple(hist.shape))
    return sums, hist, n_global
