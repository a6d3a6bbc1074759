"""Generates tests/golden/*.npz from oracle/_ref = the REFERENCE'S OWN headers
(/root/reference/include, compiled against oracle/ref_shim) run in this
container.  /root/reference does not exist on the GPU box, so the vectors are
committed; rerun this script here to regenerate them:

    python tests/golden/make_golden.py            # everything
    python tests/golden/make_golden.py gd_ cg_    # only the cases whose name starts with gd_ / cg_
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 12345

CASES = [
    # name, solver, family, d, dtype, B  (reduction policy = the kernels' default for the dtype,
    # except names ending in _eigen_sse2: CNO_POLICY_EIGEN_SSE2, the "parity mode")
    ("lbfgs_rosenbrock_d128_f64", ob.LBFGS, ob.FN_ROSENBROCK, 128, np.float64, 48),
    ("lbfgs_rosenbrock_d37_f64", ob.LBFGS, ob.FN_ROSENBROCK, 37, np.float64, 32),
    ("lbfgs_rosenbrock_d2_f64", ob.LBFGS, ob.FN_ROSENBROCK, 2, np.float64, 64),
    ("lbfgs_rosenbrock_d128_f32", ob.LBFGS, ob.FN_ROSENBROCK, 128, np.float32, 32),
    ("lbfgs_rosenbrock_d128_f64_eigen_sse2", ob.LBFGS, ob.FN_ROSENBROCK, 128, np.float64, 24),
    ("bfgs_rosenbrock_d32_f64", ob.BFGS, ob.FN_ROSENBROCK, 32, np.float64, 48),
    ("bfgs_rosenbrock_d2_f64", ob.BFGS, ob.FN_ROSENBROCK, 2, np.float64, 32),
    ("gd_rosenbrock_d8_f64", ob.GRADIENT_DESCENT, ob.FN_ROSENBROCK, 8, np.float64, 16),
    ("gd_rosenbrock_d37_f32", ob.GRADIENT_DESCENT, ob.FN_ROSENBROCK, 37, np.float32, 8),
    ("cg_rosenbrock_d8_f64", ob.CONJUGATED_GRADIENT_DESCENT, ob.FN_ROSENBROCK, 8, np.float64, 16),
    ("cg_rosenbrock_d2_f64", ob.CONJUGATED_GRADIENT_DESCENT, ob.FN_ROSENBROCK, 2, np.float64, 32),
]


def main():
    assert ob.ref_available(), "oracle/_ref is not built (needs /root/reference)"
    only = tuple(sys.argv[1:])
    for name, solver, family, d, dtype, B in CASES:
        if only and not name.startswith(only):
            continue
        x0 = ob.fill_uniform((B, d), 0, SEED, -2.0, 2.0, dtype)
        policy = ob.POLICY_EIGEN_SSE2 if name.endswith("_eigen_sse2") else None
        r = ob.minimize(solver, family, x0, impl="ref", policy=policy)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x0=x0, x=r["x"], value=r["value"],
                            gradient=r["gradient"], num_iterations=r["num_iterations"],
                            status=r["status"], nfev=r["nfev"], solver=solver, family=family,
                            policy=(policy if policy is not None else ob.device_policy(dtype)))
        print(name, "mean iters", r["num_iterations"].mean())
    if not only or "newton_" in only:
        # NewtonDescent on per-instance dense quadratics, d = 64, under CNO_POLICY_DMMA_LU (every multiply-subtract
        # of lu().solve() fused): what the tensor-core kernel csrc/cno_newton_dmma.cuh must reproduce
        rng = np.random.default_rng(SEED)
        B, d = 12, 64
        M = rng.uniform(-1, 1, (B, d, d))
        A = np.einsum("bki,bkj->bij", M, M) / d + np.eye(d)
        A = (A + A.transpose(0, 2, 1)) / 2
        A[3] = (M[3] + M[3].T) / 2   # symmetric indefinite: the pivots move
        bvec = rng.uniform(-1, 1, (B, d))
        data = np.ascontiguousarray(np.concatenate([A.transpose(0, 2, 1).reshape(B, -1), bvec], axis=1))
        x0 = ob.fill_uniform((B, d), 0, SEED, -2.0, 2.0, np.float64)
        r = ob.minimize(ob.NEWTON, ob.FN_DENSE_QUADRATIC, x0, data=data, impl="ref", policy=ob.POLICY_DMMA_LU)
        np.savez_compressed(os.path.join(HERE, "newton_dense_quadratic_d64_dmma_lu.npz"), x0=x0, data=data, x=r["x"],
                            value=r["value"], gradient=r["gradient"], num_iterations=r["num_iterations"],
                            status=r["status"], nfev=r["nfev"], solver=ob.NEWTON, family=ob.FN_DENSE_QUADRATIC,
                            policy=ob.POLICY_DMMA_LU)
        print("newton_dense_quadratic_d64_dmma_lu", "iters", r["num_iterations"])
    if only and "pins" not in only:
        return
    # the two verify.cc starts + Dockerfile.test + AL-test half norm (reference code, d = 2)
    pins = {}
    for tag, solver, family, x0 in [
        ("lbfgs_far", ob.LBFGS, ob.FN_ROSENBROCK, [15.0, 8.0]),
        ("lbfgs_near", ob.LBFGS, ob.FN_ROSENBROCK, [-1.0, 2.0]),
        ("bfgs_far", ob.BFGS, ob.FN_ROSENBROCK, [15.0, 8.0]),
        ("bfgs_near", ob.BFGS, ob.FN_ROSENBROCK, [-1.0, 2.0]),
        ("newton_far", ob.NEWTON, ob.FN_ROSENBROCK, [15.0, 8.0]),
        ("newton_near", ob.NEWTON, ob.FN_ROSENBROCK, [-1.0, 2.0]),
        ("lbfgs_quadratic", ob.LBFGS, ob.FN_DIAG_QUADRATIC, [-10.0, 2.0]),
        ("lbfgs_halfnorm", ob.LBFGS, ob.FN_HALF_SQUARED_NORM, [5.0, 5.0]),
        ("gd_far", ob.GRADIENT_DESCENT, ob.FN_ROSENBROCK, [15.0, 8.0]),
        ("gd_near", ob.GRADIENT_DESCENT, ob.FN_ROSENBROCK, [-1.0, 2.0]),
        ("cg_far", ob.CONJUGATED_GRADIENT_DESCENT, ob.FN_ROSENBROCK, [15.0, 8.0]),
        ("cg_near", ob.CONJUGATED_GRADIENT_DESCENT, ob.FN_ROSENBROCK, [-1.0, 2.0]),
    ]:
        r = ob.minimize(solver, family, np.array([x0]), impl="ref")
        pins[tag + "_x"] = r["x"][0]
        pins[tag + "_f"] = r["value"][0]
        pins[tag + "_it"] = r["num_iterations"][0]
        pins[tag + "_status"] = r["status"][0]
        print(tag, r["x"][0], r["value"][0], r["num_iterations"][0], r["status"][0])
    np.savez_compressed(os.path.join(HERE, "reference_pins_d2.npz"), **pins)


if __name__ == "__main__":
    main()
