#!/bin/bash
cd /root/repo
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/r02_gputests_d.log
tail -12 $O/r02_gputests_d.log
./tests/cpp/build/sharded_nccl 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > $O/r02_bench_e.json 2> $O/r02_bench_e.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_e.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "kernel_ms", round(d["roofline"]["kernel_ms"],1), "compute", d["roofline"]["compute"], d["parity"], d["cpu_baseline"])
    for o in d["other_configs"]: print({k:o[k] for k in ("config","kernel_ms","instances_per_s")})
except Exception as e: print("ERR", e, open("gpurun_out/r02_bench_e.err").read()[-800:])
PY
ncu --set full --clock-control none --import-source on -k regex:newton_minimize -c 1 -o $O/prof_c5_r2a python bench_configs.py c5 > $O/ncu_c5_r2a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lbfgs_minimize -c 1 -o $O/prof_lbfgs_r2b \
  python bench.py --log2-batch 14 --steps 1 --warmup 0 --no-cpu --no-e2e --no-extra > $O/ncu_r2b.log 2>&1
