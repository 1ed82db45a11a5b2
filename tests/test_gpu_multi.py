"""Two ranks on GPUs (NCCL): atom-sharded shifted_correlation equals the single-GPU result.

Skipped unless at least 2 CUDA devices are visible (the 1-GPU box runs the gloo CPU variant).
"""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys
from functools import partial
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
import mdevaluate_b200 as md
from mdevaluate_b200 import parallel
from mdevaluate_b200.synthetic import random_walk
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl')
pos, box, times = random_walk(3001, 300, seed=3, sigma=0.05, box=3.0)
a0, a1 = parallel.atom_block(pos.shape[1], rank, world)
c = md.correlation
shard = parallel.shard_atoms(md.Coordinates.from_arrays(np.ascontiguousarray(pos[:, a0:a1]), box, times))
assert shard.comm[1] == pos.shape[1]
bins = np.linspace(0, 2.0, 101)
out = {}
for name, fn in (('msd', c.msd), ('isf', partial(c.isf, q=22.7)), ('vh', partial(c.van_hove_self, bins=bins))):
    t, r = c.shifted_correlation(fn, shard.nojump, segments=9, window=0.4)
    out[name] = r
t, x4 = c.chi4(shard.nojump, 22.7)
out['chi4'] = x4
L = float(box[0])
def slab(frame):
    z = np.asarray(frame)[:, 2] %% L
    return np.where(z < 0.4 * L)[0]
def slabs(frame):
    z = np.asarray(frame)[:, 2] %% L
    return np.array([z < 0.3 * L, (z >= 0.3 * L) & (z < 0.7 * L), z >= 0.7 * L])
t, out['sub'] = c.shifted_correlation(c.msd, shard.nojump, segments=9, window=0.4,
                                      correlation=c.subensemble_correlation(slab))
dense = lambda a, b: np.arange(a, b)
t, out['dense'] = c.shifted_correlation(partial(c.isf, q=22.7), shard.nojump, segments=150, window=0.5, average=True,
                                        index_distribution=dense)
t, out['multi'], out['multi_n'] = c.shifted_correlation(partial(c.isf, q=22.7), shard.nojump, segments=9, window=0.4,
                                                        average=True,
                                                        correlation=c.multi_subensemble_correlation(slabs))
if rank == 0:
    full = md.Coordinates.from_arrays(pos, box, times)
    for name, fn in (('msd', c.msd), ('isf', partial(c.isf, q=22.7)), ('vh', partial(c.van_hove_self, bins=bins))):
        t, r = c.shifted_correlation(fn, full.nojump, segments=9, window=0.4)
        if name == 'vh':
            assert np.array_equal(out[name], r), name
        else:
            assert np.allclose(out[name], r, rtol=1e-12, atol=1e-14), name
    t, x4 = c.chi4(full.nojump, 22.7)
    assert np.allclose(out['chi4'], x4, rtol=1e-9, atol=1e-12)
    t, r = c.shifted_correlation(c.msd, full.nojump, segments=9, window=0.4, correlation=c.subensemble_correlation(slab))
    assert np.allclose(out['sub'], r, rtol=1e-12, atol=1e-14)
    t, r, n = c.shifted_correlation(partial(c.isf, q=22.7), full.nojump, segments=9, window=0.4, average=True,
                                    correlation=c.multi_subensemble_correlation(slabs))
    assert np.array_equal(out['multi_n'], n) and np.allclose(out['multi'], r, rtol=1e-9, atol=1e-12)
    t, rd = c.shifted_correlation(partial(c.isf, q=22.7), full.nojump, segments=150, window=0.5, average=True,
                                  index_distribution=dense)
    assert np.allclose(out['dense'], rd, rtol=1e-6, atol=1e-9)      # float32 offsets, another tiling per shard
    print('MULTI_GPU_OK')
assert shard.comm.p2p is not None, 'peer-memory communicator not in use'
parallel.release_peer_comms()
dist.destroy_process_group()
'''


def test_two_gpu_sharding(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    script = tmp_path / 'multi.py'
    script.write_text(SCRIPT % {'root': ROOT})
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert 'MULTI_GPU_OK' in res.stdout
