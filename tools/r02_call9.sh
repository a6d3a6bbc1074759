#!/bin/bash
# round 2, GPU call 9: full -m gpu suite at HEAD + the tensor-core NewtonDescent kernel (parity, timing, ncu)
cd /root/repo
O=gpurun_out; mkdir -p $O
T=/tmp/prof; mkdir -p $T
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r02_call9_gputests.log
tail -6 $O/r02_call9_gputests.log
python bench_configs.py c5 c5t > $O/r02_call9_c5.jsonl 2> $O/r02_call9_c5.err
cat $O/r02_call9_c5.jsonl | cut -c1-600
ncu --set full --clock-control none --import-source on -k regex:newton_dmma -c 1 -o $T/prof_c5t python bench_configs.py c5t --scale 3 > $O/ncu_c5t.log 2>&1
python tools/ncu_summary.py $T/prof_c5t.ncu-rep > $O/r02_c5t_ncu_full.txt 2>&1
cp $T/prof_c5t.ncu-rep $O/ 2>/dev/null
head -40 $O/r02_c5t_ncu_full.txt | cut -c1-150
