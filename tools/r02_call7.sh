#!/bin/bash
cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests/test_full_size_gpu.py -m gpu -q 2>&1 | tail -15 > $O/r02_gputests_f.log
tail -8 $O/r02_gputests_f.log
ncu --set full --clock-control none --import-source on -k regex:lbfgs_minimize -c 1 -o $O/prof_c3_r2 python bench_configs.py c3 --scale 3 > $O/ncu_c3_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bfgs_minimize -c 1 -o $O/prof_c4_r2 python bench_configs.py c4 --scale 3 > $O/ncu_c4_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lbfgsb_minimize -c 1 -o $O/prof_lbfgsb_r2 python - > $O/ncu_lbfgsb_r2.log 2>&1 <<'PY'
import torch, cppnumericalsolvers_b200 as cn
B, d = 1 << 13, 128
x0 = torch.empty(B, d, dtype=torch.float64, device="cuda")
cn.fill_uniform(x0, 0, 12345, -2.0, 2.0)
s = cn.Lbfgsb()
s.SetBounds(torch.full((d,), -0.5, dtype=torch.float64, device="cuda"), torch.full((d,), 0.8, dtype=torch.float64, device="cuda"))
st, pr = s.Minimize(cn.Rosenbrock(d), cn.BatchedFunctionState(x0))
torch.cuda.synchronize()
print(float(pr.num_iterations.float().mean()))
PY
ls -la $O/*.ncu-rep | tail -5
