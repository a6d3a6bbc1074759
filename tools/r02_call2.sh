#!/bin/bash
# round-2 GPU call 2: full parity suite, FP64/DMMA pipe probe, headline bench (16- and 20-warp builds),
# the other configs, one ncu capture of the headline kernel.
cd /root/repo
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r02_gputests_a.log
./tools/fp64_pipe_probe > $O/r02_fp64_pipe_probe.txt 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu > $O/r02_bench_a.json 2> $O/r02_bench_a.err
CNO_LIB=libcno_w20.so python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extra > $O/r02_bench_w20.json 2> $O/r02_bench_w20.err
python bench_configs.py hz al > $O/r02_configs_a.jsonl 2> $O/r02_configs_a.err
ncu --set full --clock-control none --import-source on -k regex:lbfgs_minimize -c 1 -o $O/prof_lbfgs_r2a \
  python bench.py --log2-batch 14 --steps 1 --warmup 0 --no-cpu --no-e2e --no-extra > $O/ncu_r2a.log 2>&1
tail -3 $O/r02_gputests_a.log; cat $O/r02_fp64_pipe_probe.txt
python - <<'PY'
import json
for f in ("r02_bench_a","r02_bench_w20"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],1), "e2e", d["e2e"] and round(d["e2e"]["value"]), "kernel_ms", round(d["roofline"]["kernel_ms"],1), d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
cat $O/r02_configs_a.jsonl | cut -c1-230
