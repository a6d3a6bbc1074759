#!/bin/bash
# round 2, GPU call 24: fresh ncu --set full of the headline kernel (final build) -> summary + per-iteration instruction mix; launch list of the bench
cd /root/repo
O=gpurun_out; mkdir -p $O
T=/tmp/prof; mkdir -p $T
ncu --set full --clock-control none --import-source on -k regex:lbfgs_minimize -c 1 -o $T/prof_lbfgs_r2c \
  python bench.py --log2-batch 14 --steps 1 --warmup 0 --no-cpu --no-e2e --no-extra > $O/ncu_r2c.log 2>&1
IT=$(python - <<'PY'
import json
l=[x for x in open('gpurun_out/ncu_r2c.log').read().strip().split('\n') if x.startswith('{')][-1]
d=json.loads(l)
print(d['config']['mean_iterations']*d['config']['batch_per_gpu'])
PY
)
python tools/ncu_summary.py $T/prof_lbfgs_r2c.ncu-rep $IT --mix-json $O/r02_instr_mix.json > $O/r02_lbfgs_rosenbrock_d128_ncu_full.txt 2>&1
head -30 $O/r02_lbfgs_rosenbrock_d128_ncu_full.txt | cut -c1-130; tail -12 $O/r02_lbfgs_rosenbrock_d128_ncu_full.txt | cut -c1-130
cat $O/r02_instr_mix.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench_steps2_warmup1.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > $O/r02_launches.log 2>&1
tail -3 $O/r02_launches.log | cut -c1-200
