"""Where the time of one END-TO-END step goes (trajectory in pinned host memory, config-2 shape).

Prints the PCIe floor (the same 12 GB copied host -> HBM in double-buffered 512 MB pieces and nothing
else), the wall time of the bench's e2e step, and a cProfile of the host side of that step.
"""
import cProfile
import pstats
import sys
import time
sys.path.insert(0, '/root/repo')
from functools import partial
import numpy as np, torch
import mdevaluate_b200 as md
from mdevaluate_b200 import correlation as c
from mdevaluate_b200._cuda import engine
from mdevaluate_b200.synthetic import random_walk_device, box_length

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
S = int(sys.argv[3]) if len(sys.argv) > 3 else 100
block, box, times = random_walk_device(N, T, seed=1, box=box_length(N), pitch=engine.row_pitch(N))
host = torch.empty((T, N, 3), dtype=torch.float32, pin_memory=True)
host.view(T, 3 * N).copy_(block[:, :3 * N])
torch.cuda.synchronize()
del block
torch.cuda.empty_cache()
host_np = host.numpy()
fn = partial(c.isf, q=22.7)

# ---- PCIe floor ----------------------------------------------------------------------------------
flat = host.view(T, 3 * N)
seg = max(1, (512 << 20) // (3 * N * 4))
slots = [torch.empty((seg, 3 * N), dtype=torch.float32, device='cuda') for _ in range(2)]
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k, f0 in enumerate(range(0, T, seg)):
        f1 = min(T, f0 + seg)
        slots[k % 2][:f1 - f0].copy_(flat[f0:f1], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('floor: %.1f ms for %.2f GB = %.1f GB/s' % (dt * 1e3, T * N * 12 / 1e9, T * N * 12 / 1e9 / dt))
del slots


def step():
    hc = md.Coordinates.from_arrays(host_np, box, times)
    hc._resident().force_compact = True
    return c.shifted_correlation(fn, hc.nojump, segments=S)


step()
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    step()
    torch.cuda.synchronize()
    print('e2e step: %.1f ms' % ((time.perf_counter() - t0) * 1e3))

prof = cProfile.Profile()
prof.enable()
step()
torch.cuda.synchronize()
prof.disable()
pstats.Stats(prof, stream=sys.stdout).sort_stats('tottime').print_stats(18)
