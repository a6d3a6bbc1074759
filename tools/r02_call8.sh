#!/bin/bash
# re-capture the C3 / C4 / Lbfgsb profiles; summarise ON the box (the .ncu-rep files exceed the 64 MiB copy-back limit together)
cd /root/repo
O=gpurun_out
T=/tmp/prof; mkdir -p $T
ncu --set full --clock-control none --import-source on -k regex:lbfgs_minimize -c 1 -o $T/prof_c3_r2 python bench_configs.py c3 --scale 3 > $O/ncu_c3_r2.log 2>&1
python tools/ncu_summary.py $T/prof_c3_r2.ncu-rep > $O/r02_c3_ncu_full.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:bfgs_minimize -c 1 -o $T/prof_c4_r2 python bench_configs.py c4 --scale 3 > $O/ncu_c4_r2.log 2>&1
python tools/ncu_summary.py $T/prof_c4_r2.ncu-rep > $O/r02_c4_ncu_full.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:lbfgsb_minimize -c 1 -o $T/prof_lbfgsb_r2 python - > $O/ncu_lbfgsb_r2.log 2>&1 <<'PY'
import torch, cppnumericalsolvers_b200 as cn
B, d = 1 << 13, 128
x0 = torch.empty(B, d, dtype=torch.float64, device="cuda")
cn.fill_uniform(x0, 0, 12345, -2.0, 2.0)
s = cn.Lbfgsb()
s.SetBounds(torch.full((d,), -0.5, dtype=torch.float64, device="cuda"), torch.full((d,), 0.8, dtype=torch.float64, device="cuda"))
st, pr = s.Minimize(cn.Rosenbrock(d), cn.BatchedFunctionState(x0))
torch.cuda.synchronize()
print(float(pr.num_iterations.float().sum()))
PY
IT=$(tail -1 $O/ncu_lbfgsb_r2.log)
python tools/ncu_summary.py $T/prof_lbfgsb_r2.ncu-rep $IT > $O/r02_lbfgsb_ncu_full.txt 2>&1
# launch list of the headline bench (share of the step per kernel)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench_steps2_warmup1.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $O/r02_launches.log 2>&1
head -20 $O/r02_c3_ncu_full.txt | cut -c1-120
