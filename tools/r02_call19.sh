#!/bin/bash
# round 2, GPU call 19: Lbfgsb<F, 10>, the logistic functor with the transposing butterfly, full suite
cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r02_call19_gputests.log
tail -3 $O/r02_call19_gputests.log
python bench_configs.py c3 2>&1 | cut -c1-220
