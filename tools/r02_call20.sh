#!/bin/bash
# round 2, GPU call 20: full suite (condition numbers of composites), refreshed ncu of the default-policy NewtonDescent kernel
cd /root/repo
O=gpurun_out; mkdir -p $O
T=/tmp/prof; mkdir -p $T
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/r02_call20_gputests.log
tail -4 $O/r02_call20_gputests.log
ncu --set full --clock-control none --import-source on -k regex:newton_minimize -c 1 -o $T/prof_c5 python bench_configs.py c5 --scale 1 > $O/ncu_c5.log 2>&1
python tools/ncu_summary.py $T/prof_c5.ncu-rep > $O/r02_c5_ncu_full.txt 2>&1
head -24 $O/r02_c5_ncu_full.txt | cut -c1-130
