"""Generates tests/golden/lbfgsb_*.npz from oracle/_ref: the reference's own solver/lbfgsb.h (compiled from
/root/reference against the Eigen-API shim), Lbfgsb<F, 5> with SetBounds.  Run where /root/reference exists."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ("num_iterations", "status", "nfev", "x", "value", "gradient", "x_delta", "f_delta", "gradient_norm")
rng = np.random.default_rng(2024)
CASES = [  # name, d, B, lower, upper
    ("rosenbrock_d2_unbounded_verify_starts", 2, None, None, None),   # src/test/verify.cc:190 (Far, Near)
    ("rosenbrock_d8_box", 8, 16, np.full(8, -0.5), np.full(8, 0.8)),
    ("rosenbrock_d37_per_instance_boxes", 37, 8, "rand", "rand"),
    ("rosenbrock_d128_box", 128, 6, np.full(128, -0.5), np.full(128, 0.8)),
    ("rosenbrock_d128_unbounded", 128, 4, None, None),
]
for name, d, B, lo, hi in CASES:
    if B is None:
        x0 = np.array([[15.0, 8.0], [-1.0, 2.0]])
        B = 2
    else:
        x0 = ob.fill_uniform((B, d), 0, 99 + d, -2.0, 2.0)
    if isinstance(lo, str):
        lo = rng.uniform(-1.5, -0.2, (B, d))
        hi = lo + rng.uniform(0.3, 2.0, (B, d))
    r = ob.ref_lbfgsb_minimize(ob.FN_ROSENBROCK, x0, lo, hi)
    extra = {}
    if lo is not None:
        extra = dict(lower=lo, upper=hi)
    np.savez_compressed(os.path.join(HERE, f"lbfgsb_{name}.npz"), x0=x0, **extra, **{k: r[k] for k in KEYS})
    print(name, r["num_iterations"], r["status"], r["value"][:3])
