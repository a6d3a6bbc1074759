#!/bin/bash
# round 2, final validation on the final build: smoke, full -m gpu suite, bench.py (complete record), the reference arm, the 8(f) configs
cd /root/repo
O=gpurun_out; mkdir -p $O
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > $O/r02_final_bench.json 2> $O/r02_final_bench.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02_final_bench.json').read().strip().split('\n') if x.startswith('{')][-1]
d=json.loads(l)
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['parity']['x_bits_equal'], 'compute', d['roofline']['compute']['frac'], 'cpu', d['cpu_baseline']['value'])
for o in d.get('other_configs',[]): print(o['config'][:40], round(o['kernel_ms'],2), round(o['instances_per_s']))
PY
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_final_bench_reference.json 2>/dev/null; tail -c 400 $O/r02_final_bench_reference.json
python bench_configs.py hz gd cg al > $O/r02_final_configs.jsonl 2>&1; cut -c1-150 $O/r02_final_configs.jsonl
