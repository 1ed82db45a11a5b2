"""Run one correlation function a few times on the config-2 shape (for ncu captures)."""
import sys
sys.path.insert(0, '/root/repo')
from functools import partial
import numpy as np, torch
import mdevaluate_b200 as md
from mdevaluate_b200._cuda import engine
from mdevaluate_b200.synthetic import random_walk_device, box_length
name = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
T = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
S = int(sys.argv[4]) if len(sys.argv) > 4 else 10
dev = torch.device('cuda', 0)
block, box, times = random_walk_device(N, T, seed=1, box=box_length(N), device=dev, pitch=engine.row_pitch(N))
coords = md.Coordinates.from_arrays(block, box, times, n_atoms=N)
c = md.correlation
bins = np.linspace(0, 3.0, 501)
fns = {'vh': partial(c.van_hove_self, bins=bins), 'ng': c.non_gaussian, 'msd': c.msd, 'isf': partial(c.isf, q=22.7)}
engine.profile_kernels(True)
for _ in range(3):
    if name == 'fused':
        t, r = c.shifted_correlation_multi({'vh': fns['vh'], 'a2': fns['ng']}, coords.nojump, segments=S)
    else:
        t, r = c.shifted_correlation(fns[name], coords.nojump, segments=S)
    print(name, engine.last_kernel_ms())
