"""ctypes binding of the CPU ORACLE (oracle/libcno_oracle.so, oracle/_ref).

TEST INFRASTRUCTURE: import this only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never from the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(HERE, "libcno_oracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libcno_ref.so")

LBFGS, BFGS, NEWTON, GRADIENT_DESCENT, CONJUGATED_GRADIENT_DESCENT = 0, 1, 2, 3, 4
FN_ROSENBROCK, FN_DIAG_QUADRATIC, FN_HALF_SQUARED_NORM, FN_LOGISTIC, FN_DENSE_QUADRATIC = range(5)
POLICY_WARP_TREE, POLICY_EIGEN_SSE2, POLICY_DMMA_TREE = 0, 1, 2


def device_policy(dtype) -> int:
    """The reduction policy the CUDA kernels of this dtype implement."""
    return POLICY_DMMA_TREE if np.dtype(dtype) == np.float64 else POLICY_WARP_TREE


class Stop(C.Structure):
    _fields_ = [
        ("num_iterations", C.c_uint64), ("x_delta", C.c_double),
        ("x_delta_violations", C.c_int32), ("f_delta", C.c_double),
        ("f_delta_violations", C.c_int32), ("f_delta_relative", C.c_int32),
        ("gradient_norm", C.c_double), ("gradient_norm_relative", C.c_int32),
        ("condition_hessian", C.c_double), ("past", C.c_int32),
        ("past_delta", C.c_double),
    ]


def default_stop() -> "Stop":
    """solver/progress.h:353-431 (DefaultStoppingSolverProgress)."""
    return Stop(10000, 1e-9, 1, 0.0, 1, 0, 1e-5, 1, 0.0, 3, 1e-6)


def conservative_stop() -> "Stop":
    """solver/progress.h:456-464 (ConservativeStoppingSolverProgress)."""
    s = default_stop()
    s.gradient_norm, s.past, s.past_delta = 5e-6, 5, 1e-10
    return s


class Problem(C.Structure):
    _fields_ = [
        ("family", C.c_int32), ("dtype", C.c_int32), ("d", C.c_int32), ("n", C.c_int32),
        ("param", C.c_double), ("data", C.c_void_p), ("data_stride", C.c_int64),
        ("policy", C.c_int32), ("mode", C.c_int32),
    ]


class BatchOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "x", "value", "gradient", "num_iterations", "status", "nfev",
        "x_delta", "f_delta", "gradient_norm")]


def build(reference: str = "/root/reference") -> None:
    """Compiles the oracle (and oracle/_ref when the reference tree is present)."""
    subprocess.run(["make", "-s", "-C", HERE, "all", f"REFERENCE={reference}"], check=True)


_oracle = None
_ref = None


def oracle_lib() -> C.CDLL:
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_LIB):
            build()
        _oracle = C.CDLL(ORACLE_LIB, mode=C.RTLD_GLOBAL)
        _oracle.cno_oracle_reduce_sum_f64.restype = C.c_double
        _oracle.cno_oracle_reduce_sum_f32.restype = C.c_float
    return _oracle


def ref_available() -> bool:
    return os.path.exists(REF_LIB)


def ref_lib() -> C.CDLL:
    global _ref
    if _ref is None:
        oracle_lib()
        _ref = C.CDLL(REF_LIB)
    return _ref


def _np_dtype(a):
    return 0 if a.dtype == np.float64 else 1


def minimize(solver: int, family: int, x0: np.ndarray, *, policy: int | None = None, stop: Stop | None = None,
             data: np.ndarray | None = None, n: int = 0, param: float = 0.0, threads: int = 0,
             impl: str = "oracle", mode: int = 0) -> dict:
    """Runs the CPU oracle ("oracle") or the reference-headers build ("ref")."""
    x0 = np.ascontiguousarray(x0)
    assert x0.dtype in (np.float64, np.float32) and x0.ndim == 2
    B, d = x0.shape
    dt = x0.dtype
    if policy is None:
        policy = device_policy(dt)
    if data is not None:
        data = np.ascontiguousarray(data, dtype=dt)
    p = Problem(family, _np_dtype(x0), d, n, param,
                data.ctypes.data if data is not None else None,
                data.shape[1] if data is not None else 0, policy, mode)
    r = dict(x=np.zeros_like(x0), value=np.zeros(B, dt), gradient=np.zeros_like(x0),
             num_iterations=np.zeros(B, np.uint32), status=np.zeros(B, np.int8),
             nfev=np.zeros(B, np.uint32), x_delta=np.zeros(B, dt), f_delta=np.zeros(B, dt),
             gradient_norm=np.zeros(B, dt))
    o = BatchOut(*[r[k].ctypes.data for k in (
        "x", "value", "gradient", "num_iterations", "status", "nfev", "x_delta", "f_delta",
        "gradient_norm")])
    fn = oracle_lib().cno_oracle_minimize if impl == "oracle" else ref_lib().cno_ref_minimize
    t0 = time.perf_counter()
    rc = fn(solver, C.byref(p), C.c_int64(B), C.c_void_p(x0.ctypes.data),
            C.byref(stop) if stop is not None else None, C.byref(o), threads)
    r["seconds"] = time.perf_counter() - t0
    if rc != 0:
        raise RuntimeError(f"oracle minimize failed: {rc}")
    return r


def evaluate(family: int, x: np.ndarray, *, policy: int | None = None, data=None, n: int = 0,
             param: float = 0.0, hessian: bool = False):
    x = np.ascontiguousarray(x)
    B, d = x.shape
    dt = x.dtype
    if policy is None:
        policy = device_policy(dt)
    if data is not None:
        data = np.ascontiguousarray(data, dtype=dt)
    p = Problem(family, _np_dtype(x), d, n, param,
                data.ctypes.data if data is not None else None,
                data.shape[1] if data is not None else 0, policy, 0)
    f = np.zeros(B, dt)
    g = np.zeros_like(x)
    H = np.zeros((B, d, d), dt) if hessian else None
    rc = oracle_lib().cno_oracle_evaluate(C.byref(p), C.c_int64(B), C.c_void_p(x.ctypes.data),
                                          C.c_void_p(f.ctypes.data), C.c_void_p(g.ctypes.data),
                                          C.c_void_p(H.ctypes.data) if hessian else None)
    if rc != 0:
        raise RuntimeError(f"oracle evaluate failed: {rc}")
    return (f, g, H) if hessian else (f, g)


def fill_uniform(shape, first: int, seed: int, lo: float, hi: float, dtype=np.float64) -> np.ndarray:
    a = np.zeros(shape, dtype)
    rc = oracle_lib().cno_oracle_fill_uniform(0 if dtype == np.float64 else 1,
                                              C.c_void_p(a.ctypes.data), C.c_int64(first),
                                              C.c_int64(a.size), C.c_uint64(seed),
                                              C.c_double(lo), C.c_double(hi))
    assert rc == 0
    return a


def cstep(io, brackt: int, info: int = 0, impl: str = "oracle"):
    """MoreThuente::cstep; io = [stx, fx, dx, sty, fy, dy, stp, fp, dp, stpmin, stpmax]."""
    arr = (C.c_double * 11)(*io)
    b, i, r = C.c_int(brackt), C.c_int(info), C.c_int(0)
    fn = oracle_lib().cno_oracle_cstep if impl == "oracle" else ref_lib().cno_ref_cstep
    fn(arr, C.byref(b), C.byref(i), C.byref(r))
    return list(arr), b.value, i.value, r.value


def num_threads() -> int:
    return oracle_lib().cno_oracle_num_threads()
