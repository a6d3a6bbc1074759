#!/bin/bash
# tools/gpu_next_session.sh -- the first GPU call of a following session, to be run through gpurun:
#   gpurun --timeout 900 -- 'bash tools/gpu_next_session.sh'
# 1. the parity suite as the driver runs it (includes the two GradientDescent-HagerZhang and the five NewtonDescent
#    kernels that changed after their last GPU run);
# 2. the AugmentedLagrangian device path's first run (tests/test_al_gpu_pending.py);
# 3. timings of what has not been timed yet (HagerZhang, AugmentedLagrangian) and of NewtonDescent after the pivot fix;
# 4. one ncu capture of the AugmentedLagrangian inner kernel.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
CNO_RUN_PENDING=1 timeout 300 python -m pytest tests/test_al_gpu_pending.py -x -q 2>&1 | tail -15
timeout 200 python bench_configs.py c5 hz al 2>&1 | tee gpurun_out/next_session_configs.jsonl | cut -c1-400
timeout 200 ncu --set full --clock-control none --import-source on -k regex:lbfgs_minimize_kernel.*AugLagFn -c 1 \
  -o gpurun_out/prof_al_inner python bench_configs.py al --scale 3 > gpurun_out/ncu_al.log 2>&1
tail -2 gpurun_out/ncu_al.log | cut -c1-200
