#!/bin/bash
# round 2: what compute-sanitizer synccheck reports on the logistic kernel, and on the headline kernel (no helper warps) for comparison
cd /root/repo
O=gpurun_out; mkdir -p $O
S=$O/r02_synccheck.txt; rm -f $S
echo "## synccheck, logistic (two warps per instance)" >> $S
timeout 200 compute-sanitizer --tool synccheck --print-limit 3 python tools/sanitize_logistic.py 2>&1 | grep "=========" | head -30 >> $S
cat > /tmp/headline_small.py <<'PY'
import sys, torch
sys.path.insert(0, '/root/repo')
import cppnumericalsolvers_b200 as cn
p = cn.Progress(num_iterations=8, gradient_norm=1e-5, x_delta=1e-9, x_delta_violations=1, past=3, past_delta=1e-6)
x0 = torch.empty(40, 128, dtype=torch.float64, device='cuda')
cn.fill_uniform(x0, 0, 1, -2.0, 2.0)
cn.Lbfgs(p).Minimize(cn.Rosenbrock(128), cn.BatchedFunctionState(x0))
torch.cuda.synchronize(); print('headline_small done')
PY
echo "## synccheck, headline kernel (one warp per instance, Tensor Memory y-history)" >> $S
timeout 200 compute-sanitizer --tool synccheck --print-limit 3 python /tmp/headline_small.py 2>&1 | grep "=========\|done" | head -20 >> $S
cat $S | cut -c1-220
