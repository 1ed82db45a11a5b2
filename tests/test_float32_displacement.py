"""The numerical argument behind the ISF kernel's float32 displacement path, checked on the CPU.

csrc/corr_kernels.cu::process_tile_sinc_f32 computes  d = (x0 - n0 L) - (xt - nt L)  from the float32
wrapped positions and the image counts WITHOUT float64:  k = n0 - nt;  the operand on the far side of the
periodic boundary is moved first, by one FMA --  k > 0: (x0 - k L) - xt,   k < 0: x0 - (xt + k L).  This file
replays that arithmetic with NumPy float32 operations (the FMA as a float64 product-sum rounded once to
float32) against the float64 expression of the reference (pbc.py:235-247 then correlation.py:218) and
checks the claim DESIGN.md makes: every component is good to ~2^-23 relative, where the naive orders
(subtract first, or always move the same operand) lose ulp(L).
"""
import numpy as np
import pytest

from oracle import md_oracle as orc
from mdevaluate_b200.synthetic import random_walk


def fma32(a, b, c):
    """float32 fused multiply-add: the product of two float32 is exact in float64; one rounding of the sum."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def kernel_displacement(x0, n0, xt, nt, box):
    """process_tile_sinc_f32, component-wise, in float32."""
    k = (n0 - nt).astype(np.float32)                  # f0 - ft: exact small integers
    up = np.maximum(k, np.float32(0))                 # images the origin is moved by
    dw = up - k                                       # images the lag frame is moved by
    nb = (-box).astype(np.float32)
    return fma32(up, nb, x0) - fma32(dw, nb, xt)


def naive_displacement(x0, n0, xt, nt, box):
    """(x0 - xt) - k L in float32: the subtraction of the two wrapped positions rounds at ulp(L) when k != 0."""
    k = (n0 - nt).astype(np.float32)
    return fma32(k, (-box).astype(np.float32), x0 - xt)


def reference_displacement(x0, n0, xt, nt, box):
    b = box.astype(np.float64)
    return (x0.astype(np.float64) - n0 * b) - (xt.astype(np.float64) - nt * b)


@pytest.mark.parametrize('n_atoms,box,sigma', [(400, 1.3, 0.12), (400, 14.4, 0.4), (300, 31.1, 1.5), (200, 100.0, 6.0)])
def test_float32_displacement_is_accurate_to_the_last_bits(n_atoms, box, sigma):
    T = 200
    pos, bx, times = random_walk(n_atoms, T, seed=int(box * 10), sigma=sigma, box=box)
    unw = orc.UnwrappedTrajectory(pos, bx, times)
    counts = unw.counts.astype(np.int64)                              # (T, N, 3)
    worst_kernel = worst_naive = 0.0
    crossed = multi = 0
    for t0, t1 in ((0, 1), (0, 7), (3, 60), (10, 199), (150, 199), (0, 199)):
        x0, xt = pos[t0], pos[t1]
        n0, nt = counts[t0], counts[t1]
        ref = reference_displacement(x0, n0, xt, nt, bx)
        got = kernel_displacement(x0, n0, xt, nt, bx).astype(np.float64)
        naive = naive_displacement(x0, n0, xt, nt, bx).astype(np.float64)
        k = n0 - nt
        crossed += int(np.count_nonzero(k))
        multi += int(np.count_nonzero(np.abs(k) >= 2))
        scale = np.maximum(np.abs(ref), 1e-30)
        worst_kernel = max(worst_kernel, float((np.abs(got - ref) / scale).max()))
        moved = k != 0
        if moved.any():
            worst_naive = max(worst_naive, float((np.abs(naive - ref)[moved]).max()))
        # same-image components: one correctly rounded subtraction
        same = ~moved
        assert (np.abs(got - ref)[same] <= 2.0 ** -24 * np.abs(ref)[same] + 1e-300).all()
        # absolute error of the kernel's order never exceeds one ulp of the displacement itself
        assert (np.abs(got - ref) <= 2.0 ** -22 * np.abs(ref) + 2.0 ** -149).all()
    assert crossed > 100, 'the walk must cross the boundary'
    if box < 2:
        assert multi > 100, 'and several images at once in the small box'
    assert worst_kernel <= 2.0 ** -22
    # the naive order is off by up to half an ulp of the BOX for the crossing components
    assert worst_naive > 4 * worst_kernel * 1e-3 or box < 2
    assert worst_naive <= np.spacing(np.float32(box)) * (1 + np.abs(counts).max())


def test_float32_isf_matches_the_float64_isf_at_small_n():
    """End of the argument: with displacements good to 2^-23 the ISF of even a few dozen atoms stays within the
    tolerance the GPU parity tests use (1e-5 |ref| + 1e-7), crossing atoms included."""
    q = 22.7
    pos, bx, times = random_walk(48, 240, seed=5, sigma=0.08, box=1.5)
    unw = orc.UnwrappedTrajectory(pos, bx, times)
    counts = unw.counts.astype(np.int64)
    worst = 0.0
    for t0 in range(0, 120, 7):
        for lag in (1, 2, 5, 13, 40, 119):
            x0, xt, n0, nt = pos[t0], pos[t0 + lag], counts[t0], counts[t0 + lag]
            d32 = kernel_displacement(x0, n0, xt, nt, bx)
            r2 = (d32[:, 0] * d32[:, 0] + d32[:, 1] * d32[:, 1] + d32[:, 2] * d32[:, 2]).astype(np.float32)
            x = np.sqrt((np.float32(q * q) * r2).astype(np.float64))
            fast = np.sinc(x / np.pi).mean()
            ref = orc.isf(unw[t0], unw[t0 + lag], q)
            assert abs(fast - ref) <= 1e-5 * abs(ref) + 1e-7
            worst = max(worst, abs(fast - ref))
    assert worst < 5e-8
