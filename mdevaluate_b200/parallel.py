"""Multi-GPU: atoms are independent, so a trajectory shards by contiguous atom blocks
(SURVEY.md section 8e).  One process per GPU (torch.distributed for the rendezvous); every rank runs
the jump scan and the correlation kernels on its own slab.  The one collective of the path -- the sum
of the small per-(origin, lag) partials (moments and histogram counts) over the shards -- is done by the
finish kernel of every call itself over NVLink peer memory (``engine.Comm`` / ``mdb_comm``: CUDA-IPC mapped
inboxes, no NCCL call on the data path); ``allreduce_partials`` (ONE all-reduce of the packed buffer through
torch.distributed) is the fallback where peer memory is not available (gloo on CPU, IPC refused).  The division
by the GLOBAL atom count, ratios and the origin statistics happen after the reduction, on every rank.
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


class ShardComm(tuple):
    """``(process group, global atom count)`` plus the peer-memory communicator ``p2p`` (``engine.Comm`` or None)."""

    def __new__(cls, group, n_atoms_global, p2p=None):
        obj = tuple.__new__(cls, (group, int(n_atoms_global)))
        obj.p2p = p2p
        return obj


_p2p_comms = {}          # process group -> engine.Comm (one set of inboxes per group and process)
P2P_CAPACITY_WORDS = 1 << 20     # 8-byte words per rank slot: 8 MiB, e.g. 10 x 82 x 500 histogram counts + sums


def peer_comm(group=None, capacity_words=None, timeout_ms=20000):
    """The peer-memory communicator of ``group`` (created on first use; None when the group is not NCCL-backed or
    CUDA IPC is refused -- every rank then agrees on the fallback)."""
    import torch
    import torch.distributed as dist
    key = group
    if key in _p2p_comms:
        return _p2p_comms[key]
    comm = None
    if dist.get_backend(group) == 'nccl' and torch.cuda.is_available():
        from ._cuda import engine
        world, rank = dist.get_world_size(group), dist.get_rank(group)

        def exchange(mine):
            everyone = [None] * world
            dist.all_gather_object(everyone, mine, group=group)
            return everyone
        ok = 1
        try:
            comm = engine.Comm(rank, world, exchange, capacity_words or P2P_CAPACITY_WORDS, timeout_ms)
        except Exception as exc:          # e.g. IPC not permitted in this container
            from .logging import logger
            logger.warning('peer-memory communicator unavailable (%s); falling back to all_reduce', exc)
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device='cuda')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            if comm is not None:
                comm.close()
            comm = None
    _p2p_comms[key] = comm
    return comm


def release_peer_comms():
    """Close the peer-memory communicators (call on every rank, before ``destroy_process_group``)."""
    import torch.distributed as dist
    for group, comm in list(_p2p_comms.items()):
        if comm is not None:
            if dist.is_initialized():
                dist.barrier(group=group)         # nobody may still be writing into an inbox that goes away
            comm.close()
    _p2p_comms.clear()


def shard_atoms(coords, group=None, n_atoms_global=None, p2p=True):
    """Mark ``coords`` (this rank's slab of atoms) as one shard of a distributed trajectory.

    ``n_atoms_global`` defaults to the sum of the shard sizes over the group.  ``p2p=False`` keeps the partial sums
    on the torch.distributed all-reduce.
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
    coords.comm = ShardComm(group, n_atoms_global, peer_comm(group) if p2p else None)
    return coords


def backend_of(comm):
    import torch.distributed as dist
    return dist.get_backend(comm[0])


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
