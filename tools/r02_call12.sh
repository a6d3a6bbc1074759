#!/bin/bash
# round 2, GPU call 12: ncu of the mixed-store tensor-core NewtonDescent kernel (layout 3)
cd /root/repo
O=gpurun_out; mkdir -p $O
T=/tmp/prof; mkdir -p $T
export CNO_NEWTON_DMMA_LAYOUT=${1:-3}
ncu --set full --clock-control none --import-source on -k regex:newton_dmma -c 1 -o $T/prof_c5t python bench_configs.py c5t --scale 3 > $O/ncu_c5t.log 2>&1
python tools/ncu_summary.py $T/prof_c5t.ncu-rep > $O/r02_c5t_ncu_full.txt 2>&1
cp $T/prof_c5t.ncu-rep $O/ 2>/dev/null
head -46 $O/r02_c5t_ncu_full.txt | cut -c1-150
