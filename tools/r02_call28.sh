#!/bin/bash
# round 2: logistic functor with a helper warp per instance (5 teams of 2 warps): parity + timing
cd /root/repo
O=gpurun_out; mkdir -p $O
L=$O/r02_call28.log; rm -f $L
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_full_size_gpu.py -m gpu -x -q -k "logistic" 2>&1 | tail -3 >> $L
timeout 300 python bench_configs.py c3 2>&1 | cut -c1-400 >> $L
timeout 300 python bench_configs.py c3 2>&1 | cut -c1-200 >> $L
cat $L
