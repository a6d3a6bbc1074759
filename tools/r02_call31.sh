#!/bin/bash
# round 2, validation of the final build after the two-warp logistic functor: smoke, full -m gpu suite, bench.py, ncu --set full of the logistic kernel
cd /root/repo
O=gpurun_out; mkdir -p $O
T=/tmp/prof; mkdir -p $T
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > $O/r02_final2_bench.json 2> $O/r02_final2_bench.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02_final2_bench.json').read().strip().split('\n') if x.startswith('{')][-1]
d=json.loads(l)
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['parity']['x_bits_equal'], 'compute', d['roofline']['compute']['frac'], 'cpu', d['cpu_baseline']['value'])
for o in d.get('other_configs',[]): print(o['config'][:40], round(o['kernel_ms'],2), round(o['instances_per_s']))
PY
ncu --set full --clock-control none --import-source on -k regex:lbfgs_minimize -c 1 -o $T/prof_c3 python bench_configs.py c3 --scale 3 > $O/ncu_c3.log 2>&1
python tools/ncu_summary.py $T/prof_c3.ncu-rep > $O/r02_c3_ncu_full.txt 2>&1
head -14 $O/r02_c3_ncu_full.txt | cut -c1-120
