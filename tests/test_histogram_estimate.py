"""The float32 bin estimate of the van Hove kernel, replayed on the CPU against the reference's rule.

csrc/corr_kernels.cu::hist_fast4 decides the bin of a sample from a float32 estimate th = r / dx - 1/2 (one FMA)
whose nearest integer is floor(r / dx) whenever r / dx lies farther from every bin edge than a margin, and sends the
rest through the exact float64 path (utils.histogram's floor_divide rule on the reference's own arithmetic).  The counts are exact only if the margin really bounds the
error of the estimate.  Here the estimate is recomputed with NumPy float32 operations -- MUFU.RSQ modelled as the
correctly rounded 1/sqrt perturbed by up to 2^-22 relative, its documented bound -- on samples placed within
1e-8 .. 1e-2 bin widths of bin edges, and every sample the estimate would decide is checked against
np.floor_divide(r, dx) of the reference's float64 r.  Two variants:

* 'f64': r^2 from the float64 displacement chain (tuning.hist_float64, per-frame boxes), margin 8e-7 t + 1e-6;
* 'f32': r^2 from the float32 displacements (process_tile_f32d, tests/test_float32_displacement.py), the default
  for van Hove alone and van Hove + alpha_2 on a constant float32 box, which needs the wider margin 1.6e-6 t + 1e-6.
"""
import numpy as np
import pytest

from test_float32_displacement import fma32, kernel_displacement, reference_displacement

MAGIC = np.float32(12582912.0)          # 1.5 * 2^23


def estimate_bins(r2f, inv_dx_f, slope, rng):
    """(bin, decided) of the kernel's branch-free estimate for float32 r^2 (hist_fast4; the constants as
    corr_host.cu::device_args sets them)."""
    exact_rs = (1.0 / np.sqrt(r2f.astype(np.float64)))
    rs = (exact_rs * (1.0 + rng.uniform(-2.0 ** -22, 2.0 ** -22, r2f.shape))).astype(np.float32)
    r = r2f * rs                                           # float32 product
    th = fma32(r, np.full_like(r, inv_dx_f), np.full_like(r, np.float32(-0.5)))
    tb = th + MAGIC
    dd = th - (tb - MAGIC)                                 # frac(r / dx) - 1/2
    c1 = np.float32(-slope)
    lim0 = np.float32(0.5 - 1e-6 - 0.5 * slope)
    lim = fma32(th, np.full_like(th, c1), np.full_like(th, lim0))
    bins = tb.view(np.int32).astype(np.int64) - 0x4B400000
    return bins, np.abs(dd) < lim


def samples_near_edges(n, box, dx, n_bins, rng):
    """Origin / lag positions (float32, wrapped, with image counts) whose distance is a bin edge +- tiny."""
    k = rng.integers(1, n_bins, n).astype(np.float64)
    eps = rng.choice([-1.0, 1.0], n) * 10.0 ** rng.uniform(-8, -2, n)
    r = (k + eps) * dx
    direction = rng.normal(size=(n, 3))
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    x0 = rng.uniform(0, box, (n, 3))
    u_t = x0 - direction * r[:, None]                      # unwrapped lag position, origin in image 0
    n_t = np.floor(u_t / box)
    x_t = u_t - n_t * box
    x0f, xtf = x0.astype(np.float32), x_t.astype(np.float32)
    bf = np.float32(box)
    xtf = np.where(xtf >= bf, xtf - bf, xtf)               # float32 rounding can land on the box length
    return x0f, np.zeros((n, 3), np.int64), xtf, (-n_t).astype(np.int64)   # u = x - n L  =>  n = -floor(u / L)


@pytest.mark.parametrize('variant,slope', [('f64', 8.0e-7), ('f32', 1.6e-6)])
@pytest.mark.parametrize('box,r_max,n_bins', [(14.4, 3.0, 500), (3.1, 1.0, 500), (31.1, 8.0, 500), (14.4, 6.0, 60)])
def test_estimate_never_decides_a_wrong_bin(variant, slope, box, r_max, n_bins):
    rng = np.random.default_rng(int(box * 100) + n_bins)
    edges = np.linspace(0, r_max, n_bins + 1)
    dx = float(np.diff(edges).mean())                                       # utils.py:333
    inv_dx_f = np.float32(1.0 / dx)
    bx = np.full(3, box, dtype=np.float32)
    x0, n0, xt, nt = samples_near_edges(400000, float(np.float32(box)), dx, n_bins, rng)
    d64 = reference_displacement(x0, n0, xt, nt, bx)
    r_ref = np.sqrt((d64[:, 0] ** 2 + d64[:, 1] ** 2) + d64[:, 2] ** 2)     # correlation.py:246-247 in float64
    ref_bin = np.floor_divide(r_ref, dx).astype(np.int64)                   # utils.py:334
    if variant == 'f64':
        r2 = d64[:, 2] * d64[:, 2] + (d64[:, 1] * d64[:, 1] + d64[:, 0] * d64[:, 0])
        r2f = r2.astype(np.float32)
    else:
        d32 = kernel_displacement(x0, n0, xt, nt, bx)
        r2f = fma32(d32[:, 2], d32[:, 2], fma32(d32[:, 1], d32[:, 1], d32[:, 0] * d32[:, 0]))
    bins, decided = estimate_bins(r2f, inv_dx_f, slope, rng)
    # the estimate decides a good part of the samples even this close to the edges, and never wrongly
    assert decided.mean() > 0.25
    wrong = decided & (bins != ref_bin)
    assert not wrong.any(), 'estimate decided %d samples wrongly (first: t=%r)' % (wrong.sum(), r_ref[wrong][:3] / dx)
    # how much of the margin the worst sample used: the error of t over the margin, for the decided ones
    t_true = r_ref / dx
    t_est = (r2f.astype(np.float64) ** 0.5) / dx
    used = (np.abs(t_est - t_true) / (slope * t_true + 1e-6))[decided].max()
    assert used < 0.75, 'estimate error reaches %.2f of the margin' % used


def test_margin_of_the_float32_variant_is_not_generous():
    """With the float64 chain's margin (8e-7 t) the float32 displacements WOULD decide wrongly now and then -- the
    reason the candidate widens it."""
    rng = np.random.default_rng(7)
    box, dx, n_bins = 31.1, 8.0 / 500, 500
    bx = np.full(3, box, dtype=np.float32)
    x0, n0, xt, nt = samples_near_edges(2000000, float(np.float32(box)), dx, n_bins, rng)
    d64 = reference_displacement(x0, n0, xt, nt, bx)
    r_ref = np.sqrt((d64[:, 0] ** 2 + d64[:, 1] ** 2) + d64[:, 2] ** 2)
    d32 = kernel_displacement(x0, n0, xt, nt, bx)
    r2f = fma32(d32[:, 2], d32[:, 2], fma32(d32[:, 1], d32[:, 1], d32[:, 0] * d32[:, 0]))
    t_true = r_ref / dx
    t_est = (r2f.astype(np.float64) ** 0.5) / dx
    worst = (np.abs(t_est - t_true) / t_true).max()
    assert 1.0e-7 < worst < 4.0e-7          # relative error of r from float32 displacements: ~2^-22
