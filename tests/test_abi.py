"""C-ABI checks that need no GPU: libcno.so loads, exports every symbol that
include/cno.h and include/cno_al.h declare, presets match the reference, and compute entry points
fail loudly (CNO_ERR_NO_DEVICE) instead of falling back to the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from cppnumericalsolvers_b200 import _lib
import cppnumericalsolvers_b200 as cn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_exported():
    hdr = "".join(open(os.path.join(ROOT, "include", h)).read() for h in ("cno.h", "cno_al.h"))
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)  # declarations only, not comments
    declared = set(re.findall(r"\b(cno_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name


def test_version_and_error_strings():
    L = _lib.lib()
    a, b = C.c_int(), C.c_int()
    L.cno_version(C.byref(a), C.byref(b))
    assert (a.value, b.value) == (0, 2)
    assert b"no CPU fallback" in L.cno_error_string(_lib.ERR_NO_DEVICE)


def test_default_preset_matches_reference():  # solver/progress.h:353-431
    p = cn.DefaultStoppingSolverProgress()
    assert (p.num_iterations, p.x_delta, p.x_delta_violations) == (10000, 1e-9, 1)
    assert (p.f_delta, p.f_delta_violations, p.f_delta_relative) == (0.0, 1, False)
    assert (p.gradient_norm, p.gradient_norm_relative) == (1e-5, True)
    assert (p.condition_hessian, p.past, p.past_delta) == (0.0, 3, 1e-6)
    q = cn.ConservativeStoppingSolverProgress()  # :456-464
    assert (q.gradient_norm, q.past, q.past_delta) == (5e-6, 5, 1e-10)


def test_status_enum_values():  # solver/progress.h:37-47
    assert [int(s) for s in cn.Status] == [-1, 0, 1, 2, 3, 4, 5, 6]


def test_supported_table():
    assert cn.Lbfgs().supported(cn.Rosenbrock(128))
    assert cn.Lbfgs().supported(cn.Rosenbrock(2))
    assert cn.Lbfgs().supported(cn.Rosenbrock(128, torch.float32))
    assert not cn.Lbfgs().supported(cn.Rosenbrock(129))


def test_struct_layouts_match_oracle_binding():
    from oracle import oracle_binding as ob
    for a, b in ((ob.Stop, _lib.Stop), (ob.Problem, _lib.Problem), (ob.BatchOut, _lib.BatchOut),
                 (ob.Constraints, _lib.Constraints), (ob.AlConfig, _lib.AlConfig), (ob.AlStop, _lib.AlStop),
                 (ob.AlOut, _lib.AlOut)):
        assert C.sizeof(a) == C.sizeof(b)
        assert [f[0] for f in a._fields_] == [f[0] for f in b._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_device_fails_loudly_no_cpu_fallback():
    L = _lib.lib()
    x0 = np.zeros((4, 128))
    out = _lib.BatchOut(x0.ctypes.data, None, None, None, None, None, None, None, None)
    prob = cn.Rosenbrock(128).problem()
    rc = L.cno_minimize_host(_lib.LBFGS, C.byref(prob), 4, x0.ctypes.data, None, C.byref(out), None)
    assert rc == _lib.ERR_NO_DEVICE
    ws = np.zeros(64, np.uint64)
    rc = L.cno_minimize(_lib.LBFGS, C.byref(prob), 4, x0.ctypes.data, None, C.byref(out),
                        ws.ctypes.data, 512, None, None)
    assert rc == _lib.ERR_NO_DEVICE
    with pytest.raises(RuntimeError):
        cn.Lbfgs().Minimize(cn.Rosenbrock(128), cn.BatchedFunctionState(torch.zeros(4, 128, dtype=torch.float64)))


def test_invalid_arguments():
    L = _lib.lib()
    prob = cn.Rosenbrock(128).problem()
    assert L.cno_minimize_host(99, C.byref(prob), 4, None, None, None, None) == _lib.ERR_INVALID_ARGUMENT
    bad = cn.Rosenbrock(129).problem()
    assert L.cno_supported(_lib.LBFGS, C.byref(bad)) == _lib.ERR_UNSUPPORTED
    n = C.c_size_t()
    assert L.cno_workspace_bytes(_lib.LBFGS, C.byref(prob), 10, C.byref(n)) == 0 and n.value >= 8


def test_al_entry_points_without_a_device():
    """include/cno_al.h: argument checking and the scratch-size contract need no GPU; the solve itself
    fails loudly (CNO_ERR_NO_DEVICE) instead of falling back to the CPU."""
    L = _lib.lib()
    prob = cn.Rosenbrock(8).problem()
    kinds = np.array([0, 1], np.int32)
    rows = np.zeros((2, 9))
    k = _lib.Constraints(1, 1, kinds.ctypes.data, rows.ctypes.data, 0)
    assert L.cno_al_supported(C.byref(prob), C.byref(k)) == _lib.OK
    assert L.cno_al_supported(C.byref(cn.Rosenbrock(64).problem()), C.byref(k)) == _lib.ERR_UNSUPPORTED
    assert L.cno_al_supported(C.byref(cn.RosenbrockFull(2).problem()), C.byref(k)) == _lib.ERR_UNSUPPORTED  # First mode only
    too_many = _lib.Constraints(33, 0, kinds.ctypes.data, rows.ctypes.data, 0)
    assert L.cno_al_supported(C.byref(prob), C.byref(too_many)) == _lib.ERR_UNSUPPORTED
    no_rows = _lib.Constraints(1, 0, kinds.ctypes.data, None, 0)
    assert L.cno_al_supported(C.byref(prob), C.byref(no_rows)) == _lib.ERR_INVALID_ARGUMENT
    sizes = []
    for B in (1, 1000, 4096):
        n = C.c_size_t(0)
        assert L.cno_al_workspace_bytes(C.byref(prob), C.byref(k), C.c_int64(B), C.byref(n)) == _lib.OK
        # x_work + best_x (2 B d) + 5 scalars per instance + best multipliers, every region 256-byte aligned
        assert n.value % 256 == 0 and n.value >= B * 8 * (2 * 8 + 5 + 2)
        sizes.append(n.value)
    assert sizes == sorted(sizes)
    cfg, stop = _lib.AlConfig(), _lib.AlStop()
    L.cno_al_default_config(C.byref(cfg))
    L.cno_al_default_stop(C.byref(stop))
    assert (cfg.penalty_growth_factor, cfg.violation_shrink_ratio, cfg.auto_scale_initial_penalty) == (10.0, 0.25, 1)
    assert (cfg.warmup_max_inner_iterations, cfg.warmup_inner_gradient_tolerance, cfg.multiplier_max) == (10, 1e-2, 1e20)
    assert (stop.num_iterations, stop.constraint_threshold, stop.kkt_stationarity_threshold) == (10000, 1e-5, 1e-4)
    out = _lib.AlOut()
    assert L.cno_al_minimize(C.byref(prob), C.byref(k), C.c_int64(4), None, None, None, None, None, None, None,
                             C.byref(out), None, C.c_size_t(0), None, None) == _lib.ERR_INVALID_ARGUMENT
    if not torch.cuda.is_available():
        assert L.cno_al_minimize(C.byref(prob), C.byref(k), C.c_int64(0), None, None, None, None, None, None, None,
                                 C.byref(out), None, C.c_size_t(0), None, None) == _lib.ERR_NO_DEVICE


def test_python_mirror_of_the_constrained_interface_reaches_the_abi():
    """cn.AugmentedLagrangian builds the C structs, the scratch buffer and the argument list and calls
    cno_al_minimize; without a device that call must come back as CNO_ERR_NO_DEVICE (no CPU fallback) --
    which also exercises every line of the Python glue before the first GPU run of the path."""
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    rows = torch.zeros(4, 2, 9, dtype=torch.float64)
    rows[:, 1, 8] = 2.0
    problem = cn.ConstrainedOptimizationProblem(cn.Rosenbrock(8), [_lib.CON_AFFINE, _lib.CON_SQNORM], rows, 1)
    solver = cn.AugmentedLagrangian(problem, cn.Lbfgs(cn.ConservativeStoppingSolverProgress()),
                                    cn.AugmentedLagrangianConfig(multiplier_max=5.0))
    assert solver.supported()
    solver.stopping_progress.num_iterations = 7
    state = cn.AugmentedLagrangeState(torch.zeros(4, 8, dtype=torch.float64), equality_multipliers=0.25, penalty=2.0)
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        solver.Minimize(state)
    with pytest.raises(_lib.CnoError) as e:
        solver._minimize(state, 0)
    assert e.value.code == _lib.ERR_NO_DEVICE
    with pytest.raises(ValueError):  # rows of the wrong scalar type
        cn.AugmentedLagrangian(cn.ConstrainedOptimizationProblem(cn.Rosenbrock(8), [0], rows[:, :1].float(), 1))._minimize(state, 0)
