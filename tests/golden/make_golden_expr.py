"""Generates tests/golden/expr_*.npz from oracle/_ref: the reference's own operators (function_expressions.h)
over FunctionCRTP functors, minimised by the reference's own solvers.  Run in the container that has /root/reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ("num_iterations", "status", "nfev", "x", "value", "gradient", "x_delta", "f_delta", "gradient_norm")
CASES = [  # name, expr, ref solver, device solver id, d, param, mode
    ("rosen_plus_half_d37_lbfgs", ob.EXPR_ROSEN_PLUS_HALF, ob.LBFGS, 0, 37, 0.0, 1),
    ("prod_d8_bfgs", ob.EXPR_PROD, ob.BFGS, 1, 8, 0.0, 1),
    ("penalty_d8_lbfgs", ob.EXPR_PENALTY, ob.LBFGS, 0, 8, 0.0, 1),
    ("second_sum_d8_newton", ob.EXPR_SECOND_SUM, ob.NEWTON, 2, 8, 0.0, 2),
    ("second_prod_d2_newton", ob.EXPR_SECOND_PROD, ob.NEWTON, 2, 2, 0.0, 2),
    ("bowl_d16_bfgs", ob.EXPR_BOWL, ob.BFGS, 1, 16, -2.5, 1),
]
for name, expr, solver, dev_solver, d, param, mode in CASES:
    x0 = ob.fill_uniform((12, d), 0, 4242 + d, -1.5, 1.5)
    r = ob.ref_minimize_expr(expr, solver, x0, param=param)
    np.savez_compressed(os.path.join(HERE, f"expr_{name}.npz"), expr=expr, device_solver=dev_solver, param=param, mode=mode,
                        x0=x0, **{k: r[k] for k in KEYS})
    print(name, r["num_iterations"])
