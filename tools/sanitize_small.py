"""tools/sanitize_small.py -- tiny run of every kernel family, for compute-sanitizer
(memcheck / racecheck / synccheck).  Usage on the GPU box:
   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cppnumericalsolvers_b200 as cn  # noqa: E402
from cppnumericalsolvers_b200 import _lib  # noqa: E402

dev = "cuda"
p = cn.Progress(num_iterations=12, gradient_norm=1e-5, x_delta=1e-9, x_delta_violations=1, past=3, past_delta=1e-6)
x0 = torch.empty(40, 128, dtype=torch.float64, device=dev)
cn.fill_uniform(x0, 0, 1, -2.0, 2.0)
cn.Lbfgs(p).Minimize(cn.Rosenbrock(128), cn.BatchedFunctionState(x0))
x0 = torch.empty(40, 37, dtype=torch.float32, device=dev)
cn.fill_uniform(x0, 0, 1, -2.0, 2.0)
cn.Lbfgs(p).Minimize(cn.Rosenbrock(37, torch.float32), cn.BatchedFunctionState(x0))
x0 = torch.empty(40, 32, dtype=torch.float64, device=dev)
cn.fill_uniform(x0, 0, 1, -2.0, 2.0)
cn.Bfgs(p).Minimize(cn.Rosenbrock(32), cn.BatchedFunctionState(x0))
rng = np.random.default_rng(0)
d = 64
M = rng.uniform(-1, 1, (20, d, d))
A = np.einsum("bki,bkj->bij", M, M) / d + np.eye(d)
A = (A + A.transpose(0, 2, 1)) / 2
data = np.concatenate([A.transpose(0, 2, 1).reshape(20, -1), rng.uniform(-1, 1, (20, d))], 1)
x0 = torch.empty(20, d, dtype=torch.float64, device=dev)
cn.fill_uniform(x0, 0, 1, -2.0, 2.0)
cn.NewtonDescent(p).Minimize(cn.DenseQuadratic(torch.from_numpy(data).to(dev), d), cn.BatchedFunctionState(x0))
n = 256
X = rng.uniform(-1, 1, (12, n, d)).astype(np.float32)
y = np.sign(rng.normal(size=(12, n))).astype(np.float32)
data = np.ascontiguousarray(np.concatenate([X.transpose(0, 2, 1).reshape(12, -1), y], 1))
cn.Lbfgs(p).Minimize(cn.Logistic(torch.from_numpy(data).to(dev), n, d, 1e-2),
                     cn.BatchedFunctionState(torch.zeros(12, d, device=dev)))
torch.cuda.synchronize()
print("sanitize_small done")
