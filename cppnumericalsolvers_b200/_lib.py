"""ctypes binding of libcno.so's C ABI (include/cno.h).

There is no CPU fallback: if the CUDA library has not been built, loading fails
loudly (ImportError) instead of degrading.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# CNO_LIB selects another build of the same library (kernel-variant experiments, e.g. libcno_w20.so)
LIB_PATH = os.path.join(HERE, os.environ.get("CNO_LIB", "libcno.so"))

# enums of include/cno.h
LBFGS, BFGS, NEWTON, GRADIENT_DESCENT, CONJUGATED_GRADIENT_DESCENT = 0, 1, 2, 3, 4
LBFGS_HAGER_ZHANG, BFGS_HAGER_ZHANG, GRADIENT_DESCENT_HAGER_ZHANG = 5, 6, 7
F64, F32 = 0, 1
FN_ROSENBROCK, FN_DIAG_QUADRATIC, FN_HALF_SQUARED_NORM, FN_LOGISTIC, FN_DENSE_QUADRATIC = range(5)
POLICY_WARP_TREE, POLICY_EIGEN_SSE2, POLICY_DMMA_TREE, POLICY_DMMA_LU = 0, 1, 2, 3
OK, ERR_INVALID_ARGUMENT, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_CUDA, ERR_WORKSPACE = 0, -1, -2, -3, -4, -5


class Stop(C.Structure):  # cno_stop_t
    _fields_ = [
        ("num_iterations", C.c_uint64), ("x_delta", C.c_double),
        ("x_delta_violations", C.c_int32), ("f_delta", C.c_double),
        ("f_delta_violations", C.c_int32), ("f_delta_relative", C.c_int32),
        ("gradient_norm", C.c_double), ("gradient_norm_relative", C.c_int32),
        ("condition_hessian", C.c_double), ("past", C.c_int32),
        ("past_delta", C.c_double),
    ]


class Problem(C.Structure):  # cno_problem_t
    _fields_ = [
        ("family", C.c_int32), ("dtype", C.c_int32), ("d", C.c_int32), ("n", C.c_int32),
        ("param", C.c_double), ("data", C.c_void_p), ("data_stride", C.c_int64),
        ("policy", C.c_int32), ("mode", C.c_int32), ("lbfgs_m", C.c_int32), ("reserved_", C.c_int32),
    ]


class BatchOut(C.Structure):  # cno_batch_out_t
    _fields_ = [(n, C.c_void_p) for n in (
        "x", "value", "gradient", "num_iterations", "status", "nfev",
        "x_delta", "f_delta", "gradient_norm")]


class LaunchInfo(C.Structure):  # cno_launch_info_t
    _fields_ = [
        ("kernel_launches", C.c_int32), ("grid", C.c_int32), ("block", C.c_int32),
        ("warps_per_cta", C.c_int32), ("dynamic_smem", C.c_int64),
        ("kernel_ms", C.c_float), ("total_ms", C.c_float),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
    ]


class Bounds(C.Structure):  # cno_bounds_t
    _fields_ = [("lower", C.c_void_p), ("upper", C.c_void_p), ("stride", C.c_int64)]


class Constraints(C.Structure):  # cno_constraints_t (include/cno_al.h)
    _fields_ = [("n_eq", C.c_int32), ("n_ineq", C.c_int32), ("kinds", C.c_void_p),
                ("data", C.c_void_p), ("data_stride", C.c_int64)]


class AlConfig(C.Structure):  # cno_al_config_t
    _fields_ = [("penalty_growth_factor", C.c_double), ("violation_shrink_ratio", C.c_double),
                ("auto_scale_initial_penalty", C.c_int32), ("penalty_auto_objective_scale", C.c_double),
                ("penalty_auto_min", C.c_double), ("penalty_auto_max", C.c_double),
                ("warmup_max_inner_iterations", C.c_int32),
                ("warmup_inner_gradient_tolerance", C.c_double), ("multiplier_max", C.c_double),
                ("kkt_gradient_tolerance", C.c_double)]


class AlStop(C.Structure):  # cno_al_stop_t
    _fields_ = [("num_iterations", C.c_uint64), ("constraint_threshold", C.c_double),
                ("kkt_stationarity_threshold", C.c_double)]


class AlOut(C.Structure):  # cno_al_out_t
    _fields_ = [(n, C.c_void_p) for n in (
        "x", "equality_multipliers", "inequality_multipliers", "penalty", "max_violation",
        "max_lagrangian_gradient", "num_iterations", "status", "nfev", "x_delta", "f_delta",
        "gradient_norm")]


CON_AFFINE, CON_SQNORM = 0, 1

# every symbol include/*.h declares
EXPORTS = (
    "cno_al_default_config", "cno_al_default_stop", "cno_al_supported", "cno_al_workspace_bytes",
    "cno_al_minimize",
    "cno_version", "cno_error_string", "cno_last_cuda_error", "cno_default_stop",
    "cno_conservative_stop", "cno_supported", "cno_workspace_bytes", "cno_minimize",
    "cno_state_bytes", "cno_minimize_steps",
    "cno_minimize_host", "cno_release_host_arena", "cno_evaluate", "cno_fill_uniform", "cno_done_bitmap", "cno_device_cstep",
    "cno_device_div_check", "cno_condition_hessian",
    "cno_allgather_done", "cno_count_done",
    "cno_lbfgsb_default_stop", "cno_lbfgsb_supported", "cno_lbfgsb_minimize",
)

_lib = None


class CnoError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = lib().cno_error_string(code).decode()
        if code == ERR_CUDA:
            s = C.c_char_p()
            lib().cno_last_cuda_error(C.byref(s))
            msg += f": {s.value.decode() if s.value else '?'}"
        super().__init__(f"{where}: {msg} ({code})")


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m cppnumericalsolvers_b200.build` "
                "(or __graft_entry__.build()). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.cno_error_string.restype = C.c_char_p
        L.cno_error_string.argtypes = [C.c_int]
        L.cno_last_cuda_error.argtypes = [C.POINTER(C.c_char_p)]
        L.cno_default_stop.argtypes = [C.POINTER(Stop)]
        L.cno_default_stop.restype = None
        L.cno_conservative_stop.argtypes = [C.POINTER(Stop)]
        L.cno_conservative_stop.restype = None
        L.cno_supported.argtypes = [C.c_int, C.POINTER(Problem)]
        L.cno_workspace_bytes.argtypes = [C.c_int, C.POINTER(Problem), C.c_int64, C.POINTER(C.c_size_t)]
        L.cno_minimize.argtypes = [
            C.c_int, C.POINTER(Problem), C.c_int64, C.c_void_p, C.POINTER(Stop),
            C.POINTER(BatchOut), C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(LaunchInfo)]
        L.cno_state_bytes.argtypes = [C.c_int, C.POINTER(Problem), C.c_int64, C.POINTER(C.c_size_t)]
        L.cno_minimize_steps.argtypes = [
            C.c_int, C.POINTER(Problem), C.c_int64, C.c_void_p, C.POINTER(Stop),
            C.POINTER(BatchOut), C.c_void_p, C.c_size_t, C.c_int32, C.c_int32, C.c_void_p,
            C.c_size_t, C.c_void_p, C.POINTER(LaunchInfo)]
        L.cno_minimize_host.argtypes = [
            C.c_int, C.POINTER(Problem), C.c_int64, C.c_void_p, C.POINTER(Stop),
            C.POINTER(BatchOut), C.POINTER(LaunchInfo)]
        L.cno_lbfgsb_default_stop.argtypes = [C.POINTER(Stop)]
        L.cno_lbfgsb_default_stop.restype = None
        L.cno_lbfgsb_supported.argtypes = [C.POINTER(Problem)]
        L.cno_lbfgsb_minimize.argtypes = [
            C.POINTER(Problem), C.POINTER(Bounds), C.c_int64, C.c_void_p, C.POINTER(Stop), C.POINTER(BatchOut),
            C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(LaunchInfo)]
        L.cno_evaluate.argtypes = [C.POINTER(Problem), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cno_fill_uniform.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_uint64,
                                       C.c_double, C.c_double, C.c_void_p]
        L.cno_done_bitmap.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.cno_device_cstep.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_int),
                                       C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.cno_condition_hessian.argtypes = [C.POINTER(Problem), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_size_t, C.c_void_p]
        L.cno_device_div_check.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cno_al_default_config.argtypes = [C.POINTER(AlConfig)]
        L.cno_al_default_config.restype = None
        L.cno_al_default_stop.argtypes = [C.POINTER(AlStop)]
        L.cno_al_default_stop.restype = None
        L.cno_al_supported.argtypes = [C.POINTER(Problem), C.POINTER(Constraints)]
        L.cno_al_workspace_bytes.argtypes = [C.POINTER(Problem), C.POINTER(Constraints), C.c_int64,
                                             C.POINTER(C.c_size_t)]
        L.cno_al_minimize.argtypes = [
            C.POINTER(Problem), C.POINTER(Constraints), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p, C.POINTER(Stop), C.POINTER(AlStop), C.POINTER(AlConfig), C.POINTER(AlOut),
            C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(LaunchInfo)]
        _lib = L
    return _lib


def check(code: int, where: str) -> None:
    if code != OK:
        raise CnoError(code, where)
