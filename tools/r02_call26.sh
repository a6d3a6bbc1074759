#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "newton or condition" 2>&1 | tail -1
python bench_configs.py c5 c5t 2>&1 | cut -c1-190
