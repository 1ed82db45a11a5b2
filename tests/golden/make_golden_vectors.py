"""Golden vectors for coordinates.vectors + rotational_autocorrelation from the REAL reference (build container only).

    python tests/golden/make_golden_vectors.py

Same loader as make_golden.py (stub modules for the reference's absent I/O dependencies).  Writes
tests/golden/vectors.npz: a small trajectory of three-site molecules, the vectors of a few frames and the
rotational autocorrelation of the mapped coordinates, for nojump (float64) and raw (float32, minimum image) frames.
"""
import os
import sys
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference, reference_coordinates          # noqa: E402


def main():
    md = load_reference()
    rng = np.random.default_rng(77)
    n_mol, n_frames, box_len = 50, 140, 2.0
    centre = rng.uniform(0, box_len, size=(n_mol, 3))
    axis = rng.standard_normal((n_mol, 3))
    pos = np.empty((n_frames, 3 * n_mol, 3), dtype=np.float32)
    for t in range(n_frames):
        centre = centre + rng.normal(0, 0.03, size=centre.shape)
        axis = axis + 0.2 * rng.standard_normal(axis.shape)
        axis /= np.linalg.norm(axis, axis=1, keepdims=True)
        side = np.cross(axis, rng.standard_normal(axis.shape))
        side /= np.linalg.norm(side, axis=1, keepdims=True)
        sites = np.stack([centre, centre + 0.06 * axis + 0.08 * side, centre + 0.06 * axis - 0.08 * side], axis=1)
        w = np.mod(sites.reshape(-1, 3), box_len).astype(np.float32)
        w[w >= np.float32(box_len)] -= np.float32(box_len)
        pos[t] = w
    box = np.array([box_len] * 3, dtype=np.float32)
    times = np.arange(n_frames) * 0.5
    coords, reader = reference_coordinates(md, pos, box, times)
    oxygen = np.arange(0, 3 * n_mol, 3)
    hydrogens = np.stack([oxygen + 1, oxygen + 2])
    out = {'pos': pos, 'box': box, 'times': times, 'oxygen': oxygen, 'hydrogens': hydrogens}
    corr = md.correlation
    # nojump frames (float64): O -> midpoint of the hydrogens, normed, no minimum image needed
    vec_nj = md.coordinates.vectors(coords.nojump, atoms_a=oxygen, atoms_b=hydrogens, normed=True)
    out['vec_nojump'] = np.stack([np.asarray(vec_nj[i]) for i in (0, 17, 139)])
    for order in (1, 2):
        t, r = corr.shifted_correlation(partial(corr.rotational_autocorrelation, order=order), vec_nj, segments=9,
                                        window=0.4)
        out['rot%d_nojump' % order] = r
    out['t'] = t
    # raw frames (float32): O -> H1 with the minimum image convention, not normed / normed
    vec_raw = md.coordinates.vectors(coords, atoms_a=oxygen, atoms_b=oxygen + 1, normed=True, box=box)
    out['vec_raw'] = np.stack([np.asarray(vec_raw[i]) for i in (0, 17, 139)])
    t, r = corr.shifted_correlation(partial(corr.rotational_autocorrelation, order=2), vec_raw, segments=9, window=0.4)
    out['rot2_raw'] = r
    vec_len = md.coordinates.vectors(coords, atoms_a=oxygen, atoms_b=oxygen + 2, box=box)
    out['vec_raw_unnormed'] = np.stack([np.asarray(vec_len[i]) for i in (5, 100)])
    np.savez_compressed(os.path.join(HERE, 'vectors.npz'), **out)
    print('wrote vectors.npz', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
