"""bsplineinterpolation_b200 -- B200 (sm_100a) implementation of the
BSplineInterpolation hot path: batched spline evaluation and the separable
control-point solve.  The numerical work lives in csrc/ (hand-written CUDA
behind the C ABI of include/bspline_b200.h); this package is the thin host
mirror of the reference API used by tests and benchmarks."""
from ._capi import BsplError, lib  # noqa: F401
from .interpolation import (BSpline, InterpolationFunction, InterpolationFunction1D,  # noqa: F401
                            InterpolationFunctionTemplate, InterpolationFunctionTemplate1D,
                            band_solve, band_solve_rows, bspline, last_kernel_ms, launch_count, reset_launch_count,
                            set_eval_path, set_fields_path, set_sweep_path)

__all__ = ["InterpolationFunction", "InterpolationFunctionTemplate", "InterpolationFunction1D",
           "InterpolationFunctionTemplate1D", "BSpline", "band_solve", "band_solve_rows", "bspline", "lib",
           "BsplError", "launch_count", "reset_launch_count", "last_kernel_ms", "set_eval_path", "set_fields_path", "set_sweep_path"]
