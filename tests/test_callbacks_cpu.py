"""Host-side callback helpers (no GPU): the batched PrintProgressCallback prints the reference's
per-iteration block (solver/solver.h:59-130: label width 18, numeric width 15, fixed 6 decimals;
vectors through Eigen's default IOFormat) for one instance of the batch."""
import io

import torch

import cppnumericalsolvers_b200 as cn


def test_print_progress_callback_layout():
    state = cn.BatchedFunctionState(x=torch.tensor([[1.0, -2.5], [0.125, 1e-7]], dtype=torch.float64),
                                    value=torch.tensor([3.0, 0.5], dtype=torch.float64),
                                    gradient=torch.tensor([[0.5, 100.0], [-1.0, 2.0]], dtype=torch.float64))
    progress = cn.BatchedProgress(num_iterations=torch.tensor([7, 12], dtype=torch.int32),
                                  status=torch.tensor([0, 4], dtype=torch.int8),
                                  nfev=torch.tensor([9, 15], dtype=torch.int32),
                                  x_delta=torch.tensor([0.25, 1e-3], dtype=torch.float64),
                                  f_delta=torch.tensor([1.5, 2e-6], dtype=torch.float64),
                                  gradient_norm=torch.tensor([100.0, 2.0], dtype=torch.float64))
    out = io.StringIO()
    cn.PrintProgressCallback(out, instance=1)(cn.Rosenbrock(2), state, progress)
    assert out.getvalue().splitlines() == [
        "--- Iteration:    12 ---",
        "  Value:                 0.500000",
        "  X:               0.125 1e-07",
        "  Gradient:        -1  2",
        "  Gradient Norm:         2.000000",
        "  X Delta:               0.001000",
        "  F Delta:               0.000002",
        "  Batch:           1 of 2 instances still running",
        "-------------------------",
    ]
    out = io.StringIO()
    cn.PrintProgressCallback(out)(cn.RosenbrockFull(2), state, progress)  # Second mode: the extra row
    lines = out.getvalue().splitlines()
    assert lines[0] == "--- Iteration:     7 ---"
    assert lines[-3] == "  Hessian Cond.:" + " " * 2 + "N/A".rjust(15)
