#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_full_size_gpu.py -m gpu -x -q -k "logistic or c3" 2>&1 | tail -2
python bench_configs.py c3 2>&1 | cut -c1-190
