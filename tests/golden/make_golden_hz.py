"""Generates tests/golden/hz_*.npz from oracle/_ref = the REFERENCE'S OWN lbfgs.h / bfgs.h with
LineSearch = linesearch::HagerZhang (/root/reference/include, compiled against oracle/ref_shim):

    python tests/golden/make_golden_hz.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    assert ob.ref_available(), "oracle/_ref is not built (needs /root/reference)"
    for name, solver, d, dtype, B in [("hz_lbfgs_rosenbrock_d128_f64", ob.LBFGS, 128, np.float64, 16),
                                      ("hz_lbfgs_rosenbrock_d37_f32", ob.LBFGS, 37, np.float32, 16),
                                      ("hz_bfgs_rosenbrock_d32_f64", ob.BFGS, 32, np.float64, 16)]:
        x0 = ob.fill_uniform((B, d), 0, 12345, -2.0, 2.0, dtype)
        r = ob.minimize(solver, ob.FN_ROSENBROCK, x0, impl="ref", linesearch=ob.LS_HAGER_ZHANG)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x0=x0, x=r["x"], value=r["value"],
                            gradient=r["gradient"], num_iterations=r["num_iterations"], status=r["status"],
                            nfev=r["nfev"], solver=solver, family=ob.FN_ROSENBROCK,
                            policy=ob.device_policy(dtype))
        print(name, "mean iterations", r["num_iterations"].mean(), "mean nfev", r["nfev"].mean())


if __name__ == "__main__":
    main()
