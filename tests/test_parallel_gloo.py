"""The N>1 path on CPU: two gloo ranks, atoms sharded, ONE all-reduce of the packed partials.

There is no GPU here, so each rank's per-(origin, lag) partial sums come from the oracle applied to
its atom block (the oracle is the checker, and here also the stand-in for the kernel output); what
is under test is the host logic of mdevaluate_b200.parallel + correlation finalisation:
block partition, packing moments and integer histogram counts into one float64 buffer, the
collective, division by the GLOBAL atom count and the per-origin statistics afterwards.
"""
import os
import socket
from functools import partial

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import md_oracle as orc
from mdevaluate_b200 import correlation as corr, parallel
from mdevaluate_b200.synthetic import random_walk

WORLD = 2
N, T, SEG, WIN = 61, 120, 6, 0.5
BINS = np.linspace(0, 2.0, 41)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _partials(unw, starts, idx, a0, a1):
    S, L = len(starts), len(idx)
    sums = np.zeros((S, L, 3))
    hist = np.zeros((S, L, len(BINS) - 1), dtype=np.int64)
    for si, s in enumerate(starts):
        origin = np.asarray(unw[s + idx[0]])[a0:a1]
        for li, l in enumerate(idx):
            d2 = ((origin - np.asarray(unw[s + l])[a0:a1]) ** 2).sum(axis=1)
            sums[si, li] = [d2.sum(), (d2 ** 2).sum(), np.sinc(np.sqrt(d2) * 7.0 / np.pi).sum()]
            hist[si, li] = orc.histogram(np.sqrt(d2), BINS)[0]
    return sums, hist


def _worker(rank, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=WORLD)
    try:
        pos, box, times = random_walk(N, T, seed=5, sigma=0.1, box=2.0)
        unw = orc.UnwrappedTrajectory(pos, box, times)
        starts, num = orc.schedule(T, SEG, WIN, 0)
        idx = orc.log_indices(0, num)
        a0, a1 = parallel.atom_block(N, rank, WORLD)
        sums, hist = _partials(unw, starts, idx, a0, a1)
        comm = (None, N)
        s_all, h_all, n_glob = parallel.allreduce_partials(torch.from_numpy(sums), torch.from_numpy(hist), a1 - a0,
                                                           comm)
        assert n_glob == N
        s_all, h_all = s_all.numpy(), h_all.numpy()
        res = {}
        for name, spec in (('msd', corr._Spec('msd')), ('isf', corr._Spec('isf', q=7.0)),
                           ('ng', corr._Spec('non_gaussian')), ('vh', corr._Spec('van_hove_self', bins=BINS))):
            res[name] = spec.finalize(s_all, h_all, n_glob, False)
        # no histogram: the buffer is the sums alone
        s_only, h_none, _ = parallel.allreduce_partials(torch.from_numpy(sums), None, a1 - a0, comm)
        assert h_none is None
        np.testing.assert_array_equal(s_only.numpy(), s_all)
        if rank == 0:
            np.savez(out, **res)
    finally:
        dist.destroy_process_group()


def test_two_rank_atom_sharding(tmp_path):
    out = str(tmp_path / 'rank0.npz')
    mp.spawn(_worker, args=(_free_port(), out), nprocs=WORLD, join=True)
    got = np.load(out)
    pos, box, times = random_walk(N, T, seed=5, sigma=0.1, box=2.0)
    unw = orc.UnwrappedTrajectory(pos, box, times)
    with np.errstate(all='ignore'):
        for name, fn in (('msd', orc.msd), ('isf', partial(orc.isf, q=7.0)), ('ng', orc.non_gaussian),
                         ('vh', partial(orc.van_hove_self, bins=BINS))):
            _, ref = orc.shifted_correlation(fn, unw, segments=SEG, window=WIN)
            if name == 'vh':
                np.testing.assert_array_equal(got[name], ref)          # integer counts: exact for any sharding
            else:
                np.testing.assert_allclose(got[name], ref, rtol=1e-12, atol=1e-14, equal_nan=True)


def test_pack_unpack_roundtrip():
    sums = torch.arange(24, dtype=torch.float64).reshape(2, 4, 3)
    hist = torch.arange(40, dtype=torch.int64).reshape(2, 4, 5) * 1000003
    buf = parallel.pack_partials(sums, hist)
    assert buf.dtype == torch.float64 and buf.numel() == 64
    s2, h2 = parallel.unpack_partials(buf, (2, 4, 3), (2, 4, 5))
    assert torch.equal(s2, sums) and torch.equal(h2, hist)
