"""Synthetic trajectories: a random walk in a periodic cubic box.

``x0 ~ U[0, L)^3``, i.i.d. Gaussian steps (sigma per component and frame),
accumulated in float64, wrapped into the box and stored as float32
``(T, N, 3)`` frame-major / xyz-interleaved -- the layout a trajectory reader
hands to mdevaluate.  ``L = (N / density)^(1/3)`` (33.4 nm^-3 = water oxygens).

The host generator is used by tests, golden-vector generation and the CPU
baseline; the device generator (torch Philox) fills HBM directly for the
bench sizes and can hand any atom sub-block back to the host for parity.
"""
from __future__ import annotations

import numpy as np

WATER_DENSITY = 33.4  # atoms / nm^3


def box_length(n_atoms, density=WATER_DENSITY):
    return float((n_atoms / density) ** (1.0 / 3.0))


def random_walk(n_atoms, n_frames, seed=20240, sigma=0.01, box=None,
                density=WATER_DENSITY, atom_block=65536):
    """Host generator.  Returns ``(positions f32 (T,N,3), box_diag f32 (3,), times f64 (T,))``."""
    L = np.float32(box_length(n_atoms, density) if box is None else box)
    Ld = float(L)
    rng = np.random.default_rng(seed)
    out = np.empty((n_frames, n_atoms, 3), dtype=np.float32)
    for a0 in range(0, n_atoms, atom_block):
        a1 = min(n_atoms, a0 + atom_block)
        x = rng.uniform(0.0, Ld, size=(a1 - a0, 3))
        for t in range(n_frames):
            if t:
                x += rng.normal(0.0, sigma, size=x.shape)
            w = np.mod(x, Ld).astype(np.float32)
            # float32 rounding can land exactly on L; fold it back like a reader would
            w[w >= L] -= L
            out[t, a0:a1] = w
    box_diag = np.array([L, L, L], dtype=np.float32)
    times = np.arange(n_frames, dtype=np.float64)
    return out, box_diag, times


def random_walk_device(n_atoms, n_frames, seed=20240, sigma=0.01, box=None,
                       density=WATER_DENSITY, device="cuda", frame_block=256,
                       out=None, pitch=None):
    """Device generator (torch Philox).  Fills / returns a float32 ``(T, pitch)`` buffer.

    ``pitch`` is the row length in floats (>= 3*N, a multiple of 4 keeps every
    frame row 16-byte aligned for the float4 streaming kernels).  The walk is
    accumulated in float64 on the device, ``frame_block`` frames at a time.
    """
    import torch

    L = np.float32(box_length(n_atoms, density) if box is None else box)
    Ld = float(L)
    if pitch is None:
        pitch = (3 * n_atoms + 3) // 4 * 4
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    if out is None:
        out = torch.zeros((n_frames, pitch), dtype=torch.float32, device=device)
    x = torch.rand((n_atoms * 3,), generator=gen, device=device, dtype=torch.float64) * Ld
    t = 0
    while t < n_frames:
        nb = min(frame_block, n_frames - t)
        steps = torch.randn((nb, n_atoms * 3), generator=gen, device=device,
                            dtype=torch.float32).to(torch.float64) * sigma
        if t == 0:
            steps[0].zero_()
        steps[0] += x
        walk = torch.cumsum(steps, dim=0)
        x = walk[-1].clone()
        w = torch.remainder(walk, Ld).to(torch.float32)
        w = torch.where(w >= float(L), w - float(L), w)
        out[t:t + nb, :3 * n_atoms] = w
        t += nb
        del steps, walk, w
    box_diag = np.array([L, L, L], dtype=np.float32)
    times = np.arange(n_frames, dtype=np.float64)
    return out, box_diag, times
