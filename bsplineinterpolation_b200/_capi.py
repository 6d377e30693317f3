"""ctypes view of include/bspline_b200.h.  Loading fails loudly when the CUDA
library has not been built -- there is no fallback implementation."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BSPL_B200_LIB", os.path.join(HERE, "libbspline_b200.so"))

OK, ERR_INVALID, ERR_CUDA, ERR_DOMAIN, ERR_ALLOC, ERR_UNSUPPORTED = range(6)
F64, F32 = 0, 1

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p
_vpp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/bspline_b200.h declares
SIGNATURES = {
    "bspl_template_create": (C.c_int, [C.c_int, C.c_int, C.c_int, _i64p, _ip, _dp, _dp, C.POINTER(_dp),
                                       C.c_int, _vpp]),
    "bspl_template_destroy": (None, [_vp]),
    "bspl_template_interpolate": (C.c_int, [_vp, _vp, C.c_int64, C.c_int, _vp, _vpp]),
    "bspl_template_interpolate_into": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int, _vp]),
    "bspl_template_sweep_axis": (C.c_int, [_vp, C.c_int, _vp, _i64p, _i64p, C.c_int64, _vp]),
    "bspl_template_sweep_axis_exchange": (C.c_int, [_vp, C.c_int, _vp, _i64p, _i64p, C.c_int64, C.c_int, _i64p,
                                                    _vpp, _ip, _i64p, _i64p, _vp]),
    "bspl_sharded_solve_create": (C.c_int, [_vp, C.c_int, C.c_int, _vpp]),
    "bspl_sharded_solve_destroy": (None, [_vp]),
    "bspl_sharded_solve_layout": (C.c_int, [_vp, _i64p, _i64p]),
    "bspl_sharded_solve_handle": (C.c_int, [_vp, C.POINTER(C.c_ubyte)]),
    "bspl_sharded_solve_connect": (C.c_int, [_vp, C.POINTER(C.c_ubyte)]),
    "bspl_sharded_solve_run": (C.c_int, [_vp, _vp, _vpp, _vp]),
    "bspl_sharded_solve_pack": (C.c_int, [_vp, _vp, _vpp, _vpp, _i64p, _i64p, _vp]),
    "bspl_sharded_solve_finish": (C.c_int, [_vp, _vpp, _vp]),
    "bspl_sharded_solve_status": (C.c_int, [_vp, _ip]),
    "bspl_ipc_alloc": (C.c_int, [C.c_int, C.c_int64, _vpp, C.POINTER(C.c_ubyte)]),
    "bspl_ipc_open": (C.c_int, [C.c_int, C.POINTER(C.c_ubyte), _vpp]),
    "bspl_ipc_close": (C.c_int, [C.c_int, _vp]),
    "bspl_ipc_free": (C.c_int, [C.c_int, _vp]),
    "bspl_template_function_from_ctrl": (C.c_int, [_vp, _vp, C.c_int64, C.c_int, _vp, _vpp]),
    "bspl_function_from_control_points": (C.c_int, [C.c_int, C.c_int, C.c_int, _i64p, _ip, C.POINTER(_dp),
                                                    _i64p, _vp, C.c_int64, C.c_int, _vpp]),
    "bspl_function_clone": (C.c_int, [_vp, _vpp]),
    "bspl_function_destroy": (None, [_vp]),
    "bspl_function_info": (C.c_int, [_vp, _ip, _ip, _ip, _i64p, _i64p, _ip, _ip, _i64p, _dp, _dp]),
    "bspl_function_device": (C.c_int, [_vp, _ip]),
    "bspl_function_knots": (C.c_int, [_vp, C.c_int, _dp, C.c_int64]),
    "bspl_function_control_points": (C.c_int, [_vp, C.c_int64, _vp]),
    "bspl_evaluate": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, _ip, _vp, C.c_int, _vp]),
    "bspl_evaluate_at": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, _ip, _vp, _i64p]),
    "bspl_evaluate_value_grad": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, _vp, C.c_int, _vp]),
    "bspl_evaluate_fields": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_int, _vp]),
    "bspl_evaluate_fields_query_major": (C.c_int, [_vp, _vp, C.c_int64, _ip, _vp, C.c_int, _vp]),
    "bspl_query_plan_create": (C.c_int, [_vp, _vp, C.c_int64, C.c_int, _vp, _vpp]),
    "bspl_template_query_plan_create": (C.c_int, [_vp, _vp, C.c_int64, C.c_int, _vp, _vpp]),
    "bspl_query_plan_evaluate": (C.c_int, [_vp, _vp, C.c_int64, _ip, C.c_int, _vp, C.c_int, _vp]),
    "bspl_query_plan_destroy": (None, [_vp]),
    "bspl_locate": (C.c_int, [_vp, _vp, C.c_int64, C.POINTER(C.c_int32), C.c_int, _vp]),
    "bspl_band_solve": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int, _dp, _dp, C.c_int64, C.c_int]),
    "bspl_band_solve_rows": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int, _dp, _dp, C.c_int64, C.c_int]),
    "bspl_host_axis_knots": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_double, _dp, _dp,
                                       C.c_int64, _i64p, _dp]),
    "bspl_host_axis_factor": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_double, C.c_double, _dp, _ip,
                                        _dp, _dp, _dp, _dp, _dp]),
    "bspl_set_eval_path": (C.c_int, [C.c_int]),
    "bspl_template_axis_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bspl_set_fields_path": (C.c_int, [C.c_int]),
    "bspl_set_sweep_path": (C.c_int, [C.c_int]),
    "bspl_launch_count": (C.c_int64, []),
    "bspl_reset_launch_count": (None, []),
    "bspl_last_kernel_ms": (C.c_double, []),
    "bspl_last_error": (C.c_char_p, []),
    "bspl_version": (C.c_char_p, []),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "bsplineinterpolation_b200: %s is missing; build it with "
                "`python -m bsplineinterpolation_b200.build` (nvcc, sm_100a). "
                "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class BsplError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("bspline_b200 error %d: %s" % (code, msg))
        self.code = code


def check(rc):
    if rc == OK:
        return
    msg = lib().bspl_last_error().decode(errors="replace")
    if rc == ERR_DOMAIN:
        raise ValueError(msg)  # std::domain_error in the reference
    if rc == ERR_ALLOC:
        raise MemoryError(msg)
    raise BsplError(rc, msg)
