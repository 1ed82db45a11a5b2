import sys, time, json
sys.path.insert(0, '/root/repo')
from functools import partial
import numpy as np, torch
import mdevaluate_b200 as md
from mdevaluate_b200._cuda import engine
from mdevaluate_b200.synthetic import random_walk_device, box_length
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
S = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device('cuda', 0)
block, box, times = random_walk_device(N, T, seed=1, box=box_length(N), device=dev, pitch=engine.row_pitch(N))
coords = md.Coordinates.from_arrays(block, box, times, n_atoms=N)
c = md.correlation
bins = np.linspace(0, 3.0, 501)
fns = {'vh': partial(c.van_hove_self, bins=bins), 'ng': c.non_gaussian, 'msd': c.msd, 'isf': partial(c.isf, q=22.7)}
engine.profile_kernels(True)
frames = coords.nojump
for name, fn in fns.items():
    for _ in range(2):
        t, r = c.shifted_correlation(fn, frames, segments=S)
    ms = []
    for _ in range(5):
        t, r = c.shifted_correlation(fn, frames, segments=S)
        ms.append(engine.last_kernel_ms())
    Sn, L = r.shape[:2]
    k = min(ms)
    print('%-4s N=%d S=%d L=%d kernel %.3f ms  %.3e evals/s  alg %.0f GB/s (frac %.3f)' % (
        name, N, Sn, L, k, N * Sn * L / k * 1e3, 12.0 * N * Sn * (L + 1) / k / 1e6, 12.0 * N * Sn * (L + 1) / k / 1e6 / 6453.1))
for _ in range(2):
    t, out = c.shifted_correlation_multi({'vh': fns['vh'], 'ng': fns['ng']}, frames, segments=S)
ms = []
for _ in range(5):
    t, out = c.shifted_correlation_multi({'vh': fns['vh'], 'ng': fns['ng']}, frames, segments=S)
    ms.append(engine.last_kernel_ms())
k = min(ms)
print('vh+ng fused kernel %.3f ms  %.3e evals/s (frac %.3f)' % (k, N * Sn * L / k * 1e3, 12.0 * N * Sn * (L + 1) / k / 1e6 / 6453.1))
print('hist total', out['vh'].sum(axis=-1).min(), out['vh'].sum(axis=-1).max())
