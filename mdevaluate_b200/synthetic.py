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


class RandomWalkStream:
    """Reader protocol over a random walk that is GENERATED on the device, time segment by time segment, and never
    resident as a whole (BASELINE config 4: 10^6 atoms x 20 000 frames = 240 GB of float32 frames).

    The atoms come in blocks of ``block_atoms`` with one Philox stream each (``seed + global block number``), so that
    any contiguous run of blocks -- the shard of one rank -- is the same trajectory whatever the number of ranks.
    ``device_segments(device)`` yields ``(t0, t1, float32 tensor (t1 - t0, >= 3N))`` in time order, which is what
    ``_cuda/resident.py`` streams through the jump scan; every call regenerates the same walk.  ``capture=(a0, n)``
    keeps the frames of ``n`` atoms from local atom ``a0`` on the host (``captured``, float32 ``(T, n, 3)``) while a
    pass runs -- the sub-block the CPU oracle checks.  Only frame 0 can be indexed.
    """

    def __init__(self, n_frames, box, first_block, n_blocks, block_atoms, seed=20240, sigma=0.01, device='cuda',
                 frame_block=128, capture=None, filename='<generated>'):
        import torch
        self.filename = filename
        self.n_frames = int(n_frames)
        self.block_atoms = int(block_atoms)
        self.blocks = list(range(int(first_block), int(first_block) + int(n_blocks)))
        self.n_atoms = self.block_atoms * len(self.blocks)
        self.dtype = np.dtype('float32')
        self.L = np.float32(box)
        self.box_constant = True
        self.box_diag = np.array([self.L] * 3, dtype=np.float32)
        self.times = np.arange(self.n_frames, dtype=np.float64)
        self.seed = int(seed)
        self.sigma = float(sigma)
        self.device = torch.device(device)
        self.frame_block = int(frame_block)
        self.capture = capture
        self.captured = None
        self._nojump_matrixes = None
        self.device_block = None
        self.host_block = None
        self.passes = 0

    def __len__(self):
        return self.n_frames

    def box_of(self, i):
        return self.box_diag

    def _start(self):
        import torch
        gens, xs = [], []
        for b in self.blocks:
            gen = torch.Generator(device=self.device)
            gen.manual_seed(self.seed + b)
            gens.append(gen)
            xs.append(torch.rand((3 * self.block_atoms,), generator=gen, device=self.device, dtype=torch.float64)
                      * float(self.L))
        return gens, xs

    def _wrap(self, walk):
        import torch
        w = torch.remainder(walk, float(self.L)).to(torch.float32)
        return torch.where(w >= float(self.L), w - float(self.L), w)

    def __getitem__(self, i):
        from .reader import MemoryFrame
        if int(i) != 0:
            raise IndexError('a generated stream delivers its frames in order (device_segments); only frame 0 can be '
                             'indexed')
        _, xs = self._start()
        pos = np.concatenate([self._wrap(x).cpu().numpy() for x in xs]).reshape(self.n_atoms, 3)
        return MemoryFrame(pos, np.diag(self.box_diag), self.times[0], 0)

    @property
    def nojump_matrixes(self):
        from .reader import NojumpError
        if self._nojump_matrixes is None:
            raise NojumpError('Nojump Data not available: {}'.format(self.filename))
        return self._nojump_matrixes

    @nojump_matrixes.setter
    def nojump_matrixes(self, mats):
        self._nojump_matrixes = mats

    def device_segments(self, device=None):
        import torch
        gens, xs = self._start()
        n3b = 3 * self.block_atoms
        pitch = (3 * self.n_atoms + 31) // 32 * 32
        out = torch.zeros((self.frame_block, pitch), dtype=torch.float32, device=self.device)
        if self.capture is not None:
            a0, n = self.capture
            self.captured = np.empty((self.n_frames, n, 3), dtype=np.float32)
        self.passes += 1
        t = 0
        while t < self.n_frames:
            nb = min(self.frame_block, self.n_frames - t)
            for k, gen in enumerate(gens):
                steps = torch.randn((nb, n3b), generator=gen, device=self.device,
                                    dtype=torch.float32).to(torch.float64) * self.sigma
                if t == 0:
                    steps[0].zero_()
                steps[0] += xs[k]
                walk = torch.cumsum(steps, dim=0)
                xs[k] = walk[-1].clone()
                out[:nb, k * n3b:(k + 1) * n3b] = self._wrap(walk)
                del steps, walk
            if self.capture is not None:
                a0, n = self.capture
                self.captured[t:t + nb] = out[:nb, 3 * a0:3 * (a0 + n)].cpu().numpy().reshape(nb, n, 3)
            yield t, t + nb, out[:nb]
            t += nb
