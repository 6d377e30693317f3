"""Seeded synthetic cases shared by the oracle tests and the GPU parity tests."""
import itertools

import numpy as np


def smooth_field(shape, rng):
    """prod_d cos(2 pi i_d / N_d - pi) (interpolation-speed-test.cpp:84-89) plus a little noise."""
    grids = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
    f = np.ones(shape)
    for g, n in zip(grids, shape):
        f = f * np.cos(2 * np.pi * g / n - np.pi)
    return f + 0.05 * rng.standard_normal(shape)


def axis_ranges(dim, rng):
    lo = rng.uniform(-2.0, 0.5, dim)
    hi = lo + rng.uniform(0.7, 3.0, dim)
    return lo, hi


def queries(lo, hi, periodic, n, rng, mode="inside"):
    """[n][dim] query points: uniform inside the range; 'wild' adds out-of-range /
    multi-period points (extrapolation on non-periodic axes, wrap on periodic)."""
    dim = len(lo)
    u = rng.uniform(0, 1, (n, dim))
    pts = lo + u * (hi - lo)
    if mode == "wild":
        k = n // 4
        span = hi - lo
        pts[:k] = lo - 2.5 * span + rng.uniform(0, 6, (k, dim)) * span
        for d in range(dim):
            if not periodic[d]:  # keep extrapolation mild: it amplifies rounding
                pts[:k, d] = np.clip(pts[:k, d], lo[d] - 0.05 * span[d], hi[d] + 0.05 * span[d])
    return pts


def adversarial_points(knots, lo, hi, periodic, rng, per_axis=400):
    """Points on knots, +-1 ulp around them, on the range ends and far outside."""
    cols = []
    for d, t in enumerate(knots):
        pick = rng.choice(t, size=per_axis, replace=True)
        around = np.concatenate([pick, np.nextafter(pick, -np.inf), np.nextafter(pick, np.inf),
                                 [lo[d], hi[d], np.nextafter(hi[d], -np.inf), np.nextafter(lo[d], np.inf)],
                                 [lo[d] - 3.7 * (hi[d] - lo[d]), hi[d] + 5.2 * (hi[d] - lo[d])]])
        if periodic[d]:
            period = hi[d] - lo[d]
            around = np.concatenate([around, pick + period, pick - 2 * period, pick + 7 * period])
        cols.append(around)
    m = min(len(c) for c in cols)
    return np.stack([rng.permutation(c)[:m] for c in cols], axis=1)


def small_shapes(dim, order, periodic):
    base = {1: (37,), 2: (19, 23), 3: (11, 9, 13), 4: (8, 7, 9, 10)}[dim]
    return tuple(max(n, order + 2) for n in base)


def all_combos(dims=(1, 2, 3), orders=range(6)):
    for dim in dims:
        for order in orders:
            for per in itertools.product([False, True], repeat=dim):
                yield dim, order, per


def extended_combos():
    """(dim, order, periodic) beyond the everyday set: orders 6 and 7 in 1-3 dimensions, and 4-D splines -- what the
    reference's templates accept (Interpolation.hpp:17) and this build serves with its generic kernels."""
    for dim in (1, 2, 3):
        for order in (6, 7):
            for per in ([False] * dim, [True] * dim, [d % 2 == 0 for d in range(dim)]):
                yield dim, order, tuple(per)
    for order in (1, 2, 3, 5, 6):
        for per in ((False, False, False, False), (True, False, True, False), (True, True, True, True)):
            yield 4, order, per


def long_axis_field(n):
    """Exactly representable pseudo-random data in [-1, 1) (an integer hash of the index): identical
    on every platform, so tests/golden/ref_outputs_long.npz stores only the reference's outputs."""
    i = np.arange(n, dtype=np.uint64)
    h = (i * np.uint64(2654435761) + np.uint64(12345)) & np.uint64(0xFFFFFFFF)
    return (h.astype(np.float64) - 2147483648.0) / 2147483648.0
