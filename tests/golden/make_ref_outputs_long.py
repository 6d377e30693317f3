#!/usr/bin/env python3
"""Generate tests/golden/ref_outputs_long.npz from the UNMODIFIED reference (oracle/_ref, plain
layout): control points, spans and values of 1-D splines on a LONG uniform axis (16 411 points: the
library factors such axes in compact form, bspl_host.h).  The data are an exact integer hash of the
index, so the fixture stores only the reference's outputs.  Build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from cases import long_axis_field  # noqa: E402
from oracle.pyoracle import RefSpline, build  # noqa: E402

N = 16411


def main():
    build()
    out = {}
    cases = [(3, False), (3, True), (5, False), (5, True), (2, True), (4, False)]
    for c, (order, per) in enumerate(cases):
        f = long_axis_field(N)
        r = RefSpline(order, f, [per], lo=[-1.5], hi=[2.25], kind="plain")
        rng = np.random.default_rng(900 + c)
        pts = rng.uniform(-1.5 - 0.5 * per, 2.25 + 0.5 * per, 256)
        out.update({"c%d_order" % c: order, "c%d_periodic" % c: per, "c%d_ctrl" % c: r.control_points(),
                    "c%d_pts" % c: pts, "c%d_spans" % c: r.spans(pts), "c%d_vals" % c: r.eval(pts)})
    out["n_cases"] = len(cases)
    out["n"] = N
    np.savez_compressed(os.path.join(HERE, "ref_outputs_long.npz"), **out)
    print("wrote ref_outputs_long.npz,", len(cases), "cases")


if __name__ == "__main__":
    main()
