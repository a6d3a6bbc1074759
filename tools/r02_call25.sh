#!/bin/bash
cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench_configs.py c4 2>&1 | cut -c1-200
