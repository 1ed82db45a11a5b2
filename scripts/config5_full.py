"""BASELINE config 5 at FULL size: dense all-lag MSD / ISF (every frame an origin, linear lags), 200 000 atoms x
50 000 frames = 120 GB of float32 frames in page-locked HOST memory, streamed to the GPU in atom slabs
(correlation._dense_average); 25 000 origins x 25 000 lags x 200 000 atoms = 1.25e14 evaluations per function.

    python scripts/config5_full.py [atoms] [frames] [out.json]
"""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from functools import partial
import numpy as np
import torch
import mdevaluate_b200 as md
from mdevaluate_b200 import correlation as c
from mdevaluate_b200._cuda import engine
from mdevaluate_b200.synthetic import RandomWalkStream, box_length

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
out_path = sys.argv[3] if len(sys.argv) > 3 else 'gpurun_out/r2_config5_full.json'
dev = torch.device('cuda', 0)
t0 = time.time()
host = torch.empty((T, N, 3), dtype=torch.float32, pin_memory=True)
t_pin = time.time() - t0
block = 25000
flat = host.view(T, 3 * N)
t0 = time.time()
for b in range(N // block):
    stream = RandomWalkStream(T, box_length(N), b, 1, block, seed=20240, device=dev, frame_block=512)
    for f0, f1, seg in stream.device_segments():
        flat[f0:f1, 3 * b * block:3 * (b + 1) * block].copy_(seg[:, :3 * block], non_blocking=True)
torch.cuda.synchronize()
t_gen = time.time() - t0
box = np.array([np.float32(box_length(N))] * 3, dtype=np.float32)
times = np.arange(T, dtype=np.float64)
print('pinned %.0f GB in %.1f s, generated in %.1f s' % (host.numel() * 4 / 1e9, t_pin, t_gen), flush=True)
dense = lambda a, b: np.arange(a, b)
result = {'workload': 'config5: dense all-lag MSD / ISF, %d atoms x %d frames from page-locked host memory in atom slabs, '
                      'segments=%d, window=0.5, linear lags, average=True' % (N, T, T // 2),
          'atoms': N, 'frames': T, 'host_GB': host.numel() * 4 / 1e9, 'pin_s': t_pin, 'generate_s': t_gen}
for name, fn in (('msd', c.msd), ('isf', partial(c.isf, q=22.7))):
    coords = md.Coordinates.from_arrays(host.numpy(), box, times)
    launches = engine.launches()
    torch.cuda.synchronize()
    t0 = time.time()
    t, r = c.shifted_correlation(fn, coords.nojump, segments=T // 2, window=0.5, average=True, index_distribution=dense)
    torch.cuda.synchronize()
    dt = time.time() - t0
    S, L = T // 2, len(r)
    evals = float(N) * S * L
    # physics of the Gaussian walk: MSD = 3 sigma^2 tau, F = exp(-q^2 sigma^2 tau / 2)
    tau = np.arange(L)
    if name == 'msd':
        check = float(np.abs(r[1:] / (3 * 0.01 ** 2 * tau[1:]) - 1).max())
    else:
        check = float(np.abs(r - np.exp(-22.7 ** 2 * 0.01 ** 2 * tau / 2)).max())
    result[name] = {'seconds': dt, 'evals': evals, 'evals_per_s': evals / dt, 'origins': S, 'lags': L,
                    'kernel_launches': engine.launches() - launches,
                    'h2d_GB': N * T * 12 / 1e9, 'deviation_from_the_gaussian_walk': check}
    print(name, json.dumps(result[name]), flush=True)
    del coords
with open(out_path, 'w') as fh:
    json.dump(result, fh, indent=1)
