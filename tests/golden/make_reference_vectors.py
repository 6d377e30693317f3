#!/usr/bin/env python3
"""Extract the reference's own golden vectors (Mathematica-computed expected
values and their inputs) from its test sources into a JSON fixture.

Run in the BUILD container only (needs the read-only checkout):
    python tests/golden/make_reference_vectors.py [/root/reference]
Writes tests/golden/reference_vectors.json, which is committed; the tests never
read /root/reference.  Sources: test/src/interpolation-test.cpp and
test/src/bspline-test.cpp (array names below are the C++ variable names).
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def grab(src, name, occurrence=0):
    """Return the brace-initialised numbers of variable `name` as a flat list."""
    hits = [m for m in re.finditer(r"\b%s\b\s*(=\s*)?\{" % re.escape(name), src)]
    m = hits[occurrence]
    i = m.end() - 1
    depth, j = 0, i
    while True:
        if src[j] == "{":
            depth += 1
        elif src[j] == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    body = src[i:j + 1].replace("{", " ").replace("}", " ")
    body = re.sub(r"//[^\n]*", "", body)
    vals = []
    for tok in body.split(","):
        tok = tok.strip()
        if tok:
            vals.append(float(eval(tok, {"__builtins__": {}})))  # "16. / 3" etc.
    return vals


def main():
    it = open(os.path.join(REF, "test/src/interpolation-test.cpp")).read()
    bs = open(os.path.join(REF, "test/src/bspline-test.cpp")).read()
    out = {"_source": "12ff54e/BSplineInterpolation test/src/{interpolation,bspline}-test.cpp",
           "interpolation": {}, "bspline": {}}
    for name in ["f", "coords_1d_half", "coords_1d", "vals_1d", "f2", "coords_2d", "vals_2d", "f3",
                 "coords_3d", "vals_3d", "vals_1d_periodic", "vals_2d_periodic",
                 "vals_1d_derivative_1", "vals_1d_derivative_periodic", "vals_2d_derivative_x2_y1",
                 "vals_3d_derivative_x1_y0_z3", "input_coords_1d", "vals_1d_nonuniform",
                 "vals_1d_nonuniform_periodic", "nonuniform_coord_for_2d",
                 "vals_2d_X_periodic_Y_nonuniform"]:
        out["interpolation"][name] = grab(it, name)
    # scalar known answers (interpolation-test.cpp:92-101)
    out["interpolation"]["extrapolate_left"] = [-0.5, -6.3167718907512755]
    out["interpolation"]["extrapolate_right"] = [6.5, -4.508470210464194]
    for name in ["knots", "cp", "coords_1d", "vals_1d", "cp2", "coords_2d", "vals_2d", "cp3",
                 "coords_3d", "vals_3d", "knots2", "vals_2d_periodic", "vals_1d_derivative_1",
                 "vals_1d_derivative_2", "vals_2d_derivative_x2_y0", "vals_2d_derivative_x1_y1",
                 "vals_2d_periodic_derivative_x1_y1"]:
        out["bspline"][name] = grab(bs, name)
    shapes = {("interpolation", "f"): 13, ("interpolation", "f2"): 25, ("interpolation", "f3"): 210,
              ("interpolation", "coords_2d"): 20, ("interpolation", "coords_3d"): 30,
              ("bspline", "cp3"): 125, ("bspline", "coords_3d"): 60, ("bspline", "knots"): 9,
              ("bspline", "knots2"): 12}
    for (grp, name), n in shapes.items():
        assert len(out[grp][name]) == n, (grp, name, len(out[grp][name]))
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote reference_vectors.json")


if __name__ == "__main__":
    main()
