#!/bin/bash
# round 2, GPU call 15: full -m gpu suite, racecheck of the new kernels, ncu of the final tensor-core NewtonDescent kernel, bench
cd /root/repo
O=gpurun_out; mkdir -p $O
T=/tmp/prof; mkdir -p $T
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/r02_call15_gputests.log
tail -4 $O/r02_call15_gputests.log
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_newton_dmma.py > $O/r02_call15_racecheck.log 2>&1
tail -4 $O/r02_call15_racecheck.log
python bench_configs.py c5 c5t > $O/r02_call15_c5.jsonl 2>&1
cut -c1-200 $O/r02_call15_c5.jsonl
ncu --set full --clock-control none --import-source on -k regex:newton_dmma -c 1 -o $T/prof_c5t python bench_configs.py c5t --scale 1 > $O/ncu_c5t.log 2>&1
python tools/ncu_summary.py $T/prof_c5t.ncu-rep > $O/r02_c5t_ncu_full.txt 2>&1
head -44 $O/r02_c5t_ncu_full.txt | cut -c1-130
python bench.py > $O/r02_call15_bench.json 2> $O/r02_call15_bench.err
tail -c 1500 $O/r02_call15_bench.json
