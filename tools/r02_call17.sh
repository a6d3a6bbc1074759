#!/bin/bash
# round 2, GPU call 17: the default-policy NewtonDescent kernel after porting the pivot fast path / shared-reciprocal divisions / column prefetch / L2 prefetch
cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r02_call17_gputests.log
tail -3 $O/r02_call17_gputests.log
python bench_configs.py c5 c5t 2>&1 | cut -c1-200
