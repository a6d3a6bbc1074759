"""HagerZhang as the LineSearch policy of Lbfgs / Bfgs / GradientDescent on the device
(linesearch/hager_zhang.h:54-552 -> `hzls` in csrc/cno_linesearch.cuh; C ABI solver ids
CNO_*_HAGER_ZHANG): bit for bit against the CPU oracle and the reference-headers fixtures."""
import glob
import os

import numpy as np
import pytest
import torch

import cppnumericalsolvers_b200 as cn
from oracle import oracle_binding as ob

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"
TDT = {np.float64: torch.float64, np.float32: torch.float32}
SOLVERS = {ob.LBFGS: cn.Lbfgs, ob.BFGS: cn.Bfgs, ob.GRADIENT_DESCENT: cn.GradientDescent}
KEYS = ("num_iterations", "status", "nfev", "x", "value", "gradient", "x_delta", "f_delta", "gradient_norm")


def _gpu(solver, fn, x0_np, progress=None):
    x0 = torch.from_numpy(x0_np).to(DEV)
    s = SOLVERS[solver](progress, linesearch=cn.HagerZhang)
    assert s.supported(fn)
    state, prog = s.Minimize(fn, cn.BatchedFunctionState(x0))
    torch.cuda.synchronize()
    return dict(x=state.x.cpu().numpy(), value=state.value.cpu().numpy(), gradient=state.gradient.cpu().numpy(),
                num_iterations=prog.num_iterations.cpu().numpy().astype(np.uint32),
                status=prog.status.cpu().numpy(), nfev=prog.nfev.cpu().numpy().astype(np.uint32),
                x_delta=prog.x_delta.cpu().numpy(), f_delta=prog.f_delta.cpu().numpy(),
                gradient_norm=prog.gradient_norm.cpu().numpy())


@pytest.mark.parametrize("solver,dtype,d,B,limit", [
    (ob.LBFGS, np.float64, 2, 128, 10000), (ob.LBFGS, np.float64, 37, 96, 10000),
    (ob.LBFGS, np.float64, 128, 160, 10000), (ob.LBFGS, np.float32, 37, 96, 10000),
    (ob.BFGS, np.float64, 8, 96, 10000), (ob.BFGS, np.float64, 32, 128, 10000),
    (ob.GRADIENT_DESCENT, np.float64, 8, 64, 400), (ob.GRADIENT_DESCENT, np.float64, 37, 64, 300)])
def test_hager_zhang_bitwise_equals_oracle(solver, dtype, d, B, limit):
    x0 = ob.fill_uniform((B, d), 0, 4242 + d, -2.0, 2.0, dtype)
    prog = cn.DefaultStoppingSolverProgress()
    prog.num_iterations = limit
    stop = ob.default_stop()
    stop.num_iterations = limit
    r = _gpu(solver, cn.Rosenbrock(d, TDT[dtype]), x0, prog)
    o = ob.minimize(solver, ob.FN_ROSENBROCK, x0, stop=stop, linesearch=ob.LS_HAGER_ZHANG)
    for k in KEYS:
        assert np.array_equal(r[k].view(np.uint8), o[k].view(np.uint8)), f"{k} differs"


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "hz_*.npz"))))
def test_hager_zhang_matches_reference_fixtures(path):
    z = np.load(path)
    d = z["x0"].shape[1]
    r = _gpu(int(z["solver"]), cn.Rosenbrock(d, TDT[z["x0"].dtype.type]), z["x0"])
    for k in ("num_iterations", "status", "nfev", "x", "value", "gradient"):
        assert np.array_equal(r[k].view(np.uint8), z[k].view(np.uint8)), k


def test_hager_zhang_reference_test_starts_and_edges():
    x0 = np.array([[15.0, 8.0], [-1.0, 2.0], [1.0, 1.0]])  # verify.cc Far / Near, and the minimiser (g = 0)
    r = _gpu(ob.LBFGS, cn.Rosenbrock(2), x0)
    o = ob.minimize(ob.LBFGS, ob.FN_ROSENBROCK, x0, linesearch=ob.LS_HAGER_ZHANG)
    for k in KEYS:
        assert np.array_equal(r[k].view(np.uint8), o[k].view(np.uint8)), k
    for x in r["x"]:
        assert (1 - x[0]) ** 2 + 100 * (x[1] - x[0] ** 2) ** 2 < 1e-4
    st, pr = cn.Lbfgs(linesearch=cn.HagerZhang).Minimize(cn.Rosenbrock(2), cn.BatchedFunctionState(
        torch.empty(0, 2, dtype=torch.float64, device=DEV)))
    assert st.x.shape[0] == 0
    assert not cn.Lbfgs(linesearch=cn.HagerZhang).supported(cn.Rosenbrock(64))
