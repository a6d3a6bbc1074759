#!/bin/bash
# round 2: ncu --set full + source page of the logistic kernel with a helper warp per instance (where the time goes)
cd /root/repo
O=gpurun_out; mkdir -p $O
T=/tmp/prof; mkdir -p $T
ncu --set full --clock-control none --import-source on -k regex:lbfgs_minimize -c 1 -o $T/prof_c3h python bench_configs.py c3 --scale 4 > $O/ncu_c3h.log 2>&1
python tools/ncu_summary.py $T/prof_c3h.ncu-rep > $O/r02_c3h_ncu_full.txt 2>&1
ncu -i $T/prof_c3h.ncu-rep --page source --csv > $O/r02_c3h_source.csv 2>/dev/null
head -30 $O/r02_c3h_ncu_full.txt | cut -c1-130
ls -la $O/r02_c3h_source.csv
