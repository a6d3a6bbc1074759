#!/bin/bash
cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests/test_expressions_gpu.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -60 > $O/r02_gputests_c.log
tail -40 $O/r02_gputests_c.log
python bench_configs.py c4 c5 gd cg > $O/r02_configs_c.jsonl 2> $O/r02_configs_c.err; cut -c1-220 $O/r02_configs_c.jsonl
