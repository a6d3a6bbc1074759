"""ctypes binding of tests/cpp/build/libcno_usertest.so (tests/cpp/user_functions.cu): user functors and
composites of functors compiled for the device from the public headers.  TEST INFRASTRUCTURE."""
import ctypes as C
import os

import numpy as np
import torch

from cppnumericalsolvers_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "cpp", "build", "libcno_usertest.so")
OP_MINIMIZE, OP_STEPS, OP_STATE_BYTES, OP_EVALUATE, OP_CONDITION = range(5)


class TestCall(C.Structure):
    _fields_ = [("op", C.c_int), ("solver", C.c_int), ("mode", C.c_int), ("batch", C.c_int64), ("x0", C.c_void_p),
                ("stop", C.POINTER(_lib.Stop)), ("out", C.POINTER(_lib.BatchOut)), ("state", C.c_void_p),
                ("state_bytes", C.c_size_t), ("max_iterations", C.c_int32), ("first_call", C.c_int32),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("stream", C.c_void_p),
                ("info", C.POINTER(_lib.LaunchInfo)), ("value", C.c_void_p), ("gradient", C.c_void_p),
                ("bytes", C.POINTER(C.c_size_t)), ("lbfgs_m", C.c_int)]


_handle = None


def lib():
    global _handle
    if _handle is None:
        if not os.path.exists(LIB):
            from cppnumericalsolvers_b200 import build
            build.build_cpp_tests()
        _lib.lib()
        _handle = C.CDLL(LIB)
        _handle.cno_test_expr.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, C.POINTER(TestCall)]
    return _handle


KEYS = ("x", "value", "gradient", "num_iterations", "status", "nfev", "x_delta", "f_delta", "gradient_norm")


def _stop_ptr(stop):
    """cno_stop_t* from any ctypes mirror of the struct (oracle_binding.Stop and _lib.Stop have one layout)."""
    if stop is None:
        return None
    assert C.sizeof(stop) == C.sizeof(_lib.Stop)
    return C.cast(C.pointer(stop), C.POINTER(_lib.Stop))


def _outputs(B, d, dt, dev):
    t = dict(x=torch.empty(B, d, dtype=dt, device=dev), value=torch.empty(B, dtype=dt, device=dev),
             gradient=torch.empty(B, d, dtype=dt, device=dev),
             num_iterations=torch.zeros(B, dtype=torch.int32, device=dev),
             status=torch.zeros(B, dtype=torch.int8, device=dev), nfev=torch.zeros(B, dtype=torch.int32, device=dev),
             x_delta=torch.empty(B, dtype=dt, device=dev), f_delta=torch.empty(B, dtype=dt, device=dev),
             gradient_norm=torch.empty(B, dtype=dt, device=dev))
    return t, _lib.BatchOut(*[t[k].data_ptr() for k in KEYS])


def _numpy(t):
    r = {k: v.cpu().numpy() for k, v in t.items()}
    r["num_iterations"] = r["num_iterations"].astype(np.uint32)
    r["nfev"] = r["nfev"].astype(np.uint32)
    return r


def minimize(expr, solver, x0_np, *, param=0.0, mode=1, stop=None, dev="cuda:0", lbfgs_m=0):
    """cno_<tag>_minimize of the composite `expr` (ids = oracle_binding.EXPR_*)."""
    x0 = torch.from_numpy(np.ascontiguousarray(x0_np)).to(dev)
    B, d = x0.shape
    t, out = _outputs(B, d, x0.dtype, dev)
    ws = torch.zeros(256, dtype=torch.uint8, device=dev)
    info = _lib.LaunchInfo()
    call = TestCall(OP_MINIMIZE, solver, mode, B, x0.data_ptr(), _stop_ptr(stop),
                    C.pointer(out), None, 0, 0, 0, ws.data_ptr(), 256, None, C.pointer(info), None, None, None, lbfgs_m)
    rc = lib().cno_test_expr(expr, param, 0 if x0.dtype == torch.float64 else 1, d, C.byref(call))
    torch.cuda.synchronize()
    if rc != 0:
        raise _lib.CnoError(rc, "cno_test_expr(minimize)")
    r = _numpy(t)
    r["launch"] = info
    return r


def minimize_steps(expr, solver, x0_np, every, *, param=0.0, mode=1, stop=None, dev="cuda:0", callback=None):
    """Rounds of `every` iterations through cno_<tag>_minimize_steps until every instance has stopped."""
    x0 = torch.from_numpy(np.ascontiguousarray(x0_np)).to(dev)
    B, d = x0.shape
    t, out = _outputs(B, d, x0.dtype, dev)
    ws = torch.zeros(256, dtype=torch.uint8, device=dev)
    nbytes = C.c_size_t(0)
    dt = 0 if x0.dtype == torch.float64 else 1
    call = TestCall(OP_STATE_BYTES, solver, mode, B, None, None, None, None, 0, 0, 0, None, 0, None, None, None, None,
                    C.pointer(nbytes))
    rc = lib().cno_test_expr(expr, param, dt, d, C.byref(call))
    if rc != 0:
        raise _lib.CnoError(rc, "cno_test_expr(state_bytes)")
    state = torch.zeros(max(nbytes.value, 16), dtype=torch.uint8, device=dev)
    rounds, first = 0, 1
    while True:
        call = TestCall(OP_STEPS, solver, mode, B, x0.data_ptr(), _stop_ptr(stop),
                        C.pointer(out), state.data_ptr(), state.numel(), every, first, ws.data_ptr(), 256, None, None,
                        None, None, None)
        rc = lib().cno_test_expr(expr, param, dt, d, C.byref(call))
        torch.cuda.synchronize()
        if rc != 0:
            raise _lib.CnoError(rc, "cno_test_expr(steps)")
        rounds, first = rounds + 1, 0
        if callback is not None:
            callback(t)
        if bool((t["status"] != 0).all().item()):
            break
    r = _numpy(t)
    r["rounds"] = rounds
    return r


def evaluate(expr, x_np, *, param=0.0, dev="cuda:0"):
    x = torch.from_numpy(np.ascontiguousarray(x_np)).to(dev)
    B, d = x.shape
    f = torch.empty(B, dtype=x.dtype, device=dev)
    g = torch.empty_like(x)
    call = TestCall(OP_EVALUATE, 0, 1, B, x.data_ptr(), None, None, None, 0, 0, 0, None, 0, None, None, f.data_ptr(),
                    g.data_ptr(), None)
    rc = lib().cno_test_expr(expr, param, 0 if x.dtype == torch.float64 else 1, d, C.byref(call))
    torch.cuda.synchronize()
    if rc != 0:
        raise _lib.CnoError(rc, "cno_test_expr(evaluate)")
    return f.cpu().numpy(), g.cpu().numpy()


def condition_hessian(expr, x_np, *, param=0.0, dev="cuda:0"):
    """cno_<tag>_condition_hessian of the composite: Progress::condition_hessian (progress.h:203-210) per row of x."""
    x = torch.from_numpy(np.ascontiguousarray(x_np)).to(dev)
    B, d = x.shape
    c = torch.empty(B, dtype=x.dtype, device=dev)
    ws = torch.zeros(256, dtype=torch.uint8, device=dev)
    call = TestCall(OP_CONDITION, 0, 2, B, x.data_ptr(), None, None, None, 0, 0, 0, ws.data_ptr(), ws.numel(), None, None,
                    c.data_ptr(), None, None)
    rc = lib().cno_test_expr(expr, param, 0 if x.dtype == torch.float64 else 1, d, C.byref(call))
    torch.cuda.synchronize()
    if rc != 0:
        raise _lib.CnoError(rc, "cno_test_expr(condition_hessian)")
    return c.cpu().numpy()
