#!/bin/bash
cd /root/repo
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/r02_gputests_e.log
tail -12 $O/r02_gputests_e.log
python bench_configs.py c5 > $O/r02_configs_d.jsonl 2> $O/r02_configs_d.err; cut -c1-200 $O/r02_configs_d.jsonl
python - <<'PY' 2>&1 | tail -5
# Lbfgsb throughput: Rosenbrock d=128 fp64, box [-0.5, 0.8]^d, B = 2^16
import numpy as np, torch, json
import cppnumericalsolvers_b200 as cn
B, d = 1 << 16, 128
x0 = torch.empty(B, d, dtype=torch.float64, device="cuda")
cn.fill_uniform(x0, 0, 12345, -2.0, 2.0)
s = cn.Lbfgsb()
s.SetBounds(torch.full((d,), -0.5, dtype=torch.float64, device="cuda"), torch.full((d,), 0.8, dtype=torch.float64, device="cuda"))
for _ in range(2):
    st, pr = s.Minimize(cn.Rosenbrock(d), cn.BatchedFunctionState(x0), timed=True)
print(json.dumps({"config": "lbfgsb rosenbrock d128 f64 box[-0.5,0.8]", "batch": B, "kernel_ms": pr.launch.kernel_ms,
                  "instances_per_s": B / pr.launch.kernel_ms * 1e3, "mean_iterations": float(pr.num_iterations.float().mean()),
                  "warps_per_cta": pr.launch.warps_per_cta}))
PY
