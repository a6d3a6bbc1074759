#!/bin/bash
# tools/gpu_check.sh [tag] -- run on the GPU box via gpurun: parity tests, bench, optional ncu.
TAG=${1:-x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 2 --warmup 1 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_$TAG.json
python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['clocks'])"
if [ "$2" = "ncu" ]; then
  ncu --set full --clock-control none --import-source on -k regex:lbfgs_minimize -c 1 -o gpurun_out/prof_lbfgs_$TAG python bench.py --log2-batch 14 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/ncu_$TAG.log 2>&1
  tail -1 gpurun_out/ncu_$TAG.log | cut -c1-120
fi
