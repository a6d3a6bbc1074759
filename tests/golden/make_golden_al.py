"""Generates tests/golden/al_*.npz from oracle/_ref = the REFERENCE'S OWN
augmented_lagrangian.h / function_penalty.h / function_expressions.h / lbfgs.h
(/root/reference/include, compiled against oracle/ref_shim) run in this container:

    python tests/golden/make_golden_al.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ("x", "equality_multipliers", "inequality_multipliers", "penalty", "max_violation",
        "max_lagrangian_gradient", "num_iterations", "status", "nfev", "x_delta", "f_delta",
        "gradient_norm")


def case(name, family, d, B, dtype, n_eq, kinds, seed, outer_limit):
    rng = np.random.default_rng(seed)
    x0 = ob.fill_uniform((B, d), 0, seed, -1.5, 1.5, dtype)
    rows = rng.uniform(-1, 1, (B, len(kinds), d + 1)).astype(dtype)
    for i, k in enumerate(kinds):
        if k == ob.CON_SQNORM:
            rows[:, i, d] = 2.0 + rng.uniform(0, 1, B)
    stop = ob.al_default_stop()
    stop.num_iterations = outer_limit
    r = ob.al_minimize(family, x0, kinds, rows, n_eq, outer_stop=stop, impl="ref")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x0=x0, kinds=np.array(kinds, np.int32), rows=rows,
                        n_eq=n_eq, family=family, outer_limit=outer_limit, policy=ob.device_policy(dtype),
                        **{k: r[k] for k in KEYS})
    print(name, "outer iterations", r["num_iterations"], "status", r["status"], "max violation",
          r["max_violation"].max())


def main():
    assert ob.ref_available(), "oracle/_ref is not built (needs /root/reference)"
    case("al_halfnorm_d8_f64_affine_eq_ball_ineq", ob.FN_HALF_SQUARED_NORM, 8, 12, np.float64, 1,
         [ob.CON_AFFINE, ob.CON_SQNORM], 5, 40)
    case("al_rosenbrock_d8_f64_ball_eq_affine_ineq", ob.FN_ROSENBROCK, 8, 8, np.float64, 1,
         [ob.CON_SQNORM, ob.CON_AFFINE], 7, 25)
    case("al_rosenbrock_d37_f64_two_affine_ineq", ob.FN_ROSENBROCK, 37, 6, np.float64, 0,
         [ob.CON_AFFINE, ob.CON_AFFINE], 9, 25)


if __name__ == "__main__":
    main()
