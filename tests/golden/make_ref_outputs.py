#!/usr/bin/env python3
"""Generate tests/golden/ref_outputs.npz from the UNMODIFIED reference headers
(oracle/_ref, built by `make -f oracle/Makefile` in the build container):
control points, spans, values and one mixed derivative on seeded inputs for a
spread of (dim, order, periodicity).  The fixture is committed; tests replay it
without the reference."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from cases import all_combos, axis_ranges, queries, small_shapes, smooth_field  # noqa: E402
from oracle.pyoracle import RefSpline, build  # noqa: E402


def main():
    build()
    out = {}
    c = 0
    for dim, order, per in all_combos():
        if (dim * 7 + order * 3 + sum(per)) % 3 != 0 and not (dim == 3 and order == 3):
            continue  # keep the fixture small
        rng = np.random.default_rng(4242 + c)
        shape = small_shapes(dim, order, per)
        lo, hi = axis_ranges(dim, rng)
        f = smooth_field(shape, rng)
        r = RefSpline(order, f, per, lo=lo, hi=hi, kind="cell")
        rp = RefSpline(order, f, per, lo=lo, hi=hi, kind="plain")
        rlo = np.array([r.range(d)[0] for d in range(dim)])
        rhi = np.array([r.range(d)[1] for d in range(dim)])
        pts = queries(rlo, rhi, per, 64, rng, mode="wild")
        dv = [(1 + d) % (order + 1) for d in range(dim)]
        out.update({"c%d_order" % c: order, "c%d_periodic" % c: np.array(per), "c%d_f" % c: f,
                    "c%d_lo" % c: lo, "c%d_hi" % c: hi, "c%d_ctrl" % c: rp.control_points(),
                    "c%d_pts" % c: pts, "c%d_spans" % c: r.spans(pts), "c%d_vals" % c: r.eval(pts),
                    "c%d_dv" % c: np.array(dv), "c%d_dvals" % c: r.deriv(pts, dv)})
        c += 1
    out["n_cases"] = c
    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)
    print("wrote ref_outputs.npz with", c, "cases")


if __name__ == "__main__":
    main()
