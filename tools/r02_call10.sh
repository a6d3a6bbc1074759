#!/bin/bash
# round 2, GPU call 10: tensor-core NewtonDescent after the ILP restructuring (parity, timing, ncu)
cd /root/repo
O=gpurun_out; mkdir -p $O
T=/tmp/prof; mkdir -p $T
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "newton or division" 2>&1 | tail -5 > $O/r02_call10_gputests.log
tail -3 $O/r02_call10_gputests.log
python bench_configs.py c5t c5 > $O/r02_call10_c5.jsonl 2> $O/r02_call10_c5.err
cut -c1-330 $O/r02_call10_c5.jsonl
ncu --set full --clock-control none --import-source on -k regex:newton_dmma -c 1 -o $T/prof_c5t python bench_configs.py c5t --scale 3 > $O/ncu_c5t.log 2>&1
python tools/ncu_summary.py $T/prof_c5t.ncu-rep > $O/r02_c5t_ncu_full.txt 2>&1
cp $T/prof_c5t.ncu-rep $O/ 2>/dev/null
head -22 $O/r02_c5t_ncu_full.txt | cut -c1-150
