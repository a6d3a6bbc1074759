#!/bin/bash
# round 2, GPU call 23: L-BFGS main loop with the square-root / division free pre-tests: full suite + headline bench
cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r02_call23_gputests.log
tail -3 $O/r02_call23_gputests.log
python bench.py --no-cpu --no-extra > $O/r02_call23_bench.json 2> $O/r02_call23_bench.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02_call23_bench.json').read().strip().split('\n') if x.startswith('{')][-1]
d=json.loads(l)
print({k:d.get(k) for k in ('value','ms_per_step')}, d['e2e']['value'], d['parity'])
PY
