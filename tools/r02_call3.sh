#!/bin/bash
# round-2 GPU call 3: parity suite incl. the expression templates / user-functor boundary; kernel variants
cd /root/repo
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/r02_gputests_b.log
tail -8 $O/r02_gputests_b.log
for V in libcno.so libcno_s2.so libcno_w20.so libcno_w20s2.so; do
  CNO_LIB=$V python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-extra > $O/r02_bench_$V.json 2> $O/r02_bench_$V.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_$V.json").read().strip().splitlines()[-1])
    print("$V", "value", round(d["value"]), "kernel_ms", round(d["roofline"]["kernel_ms"],1), "compute", d["roofline"]["compute"] and round(d["roofline"]["compute"]["frac"],3), d["parity"]["x_bits_equal"])
except Exception as e: print("$V", "ERR", e, open("gpurun_out/r02_bench_$V.err").read()[-600:])
PY
done
python bench_configs.py c5 c4 > $O/r02_configs_b.jsonl 2> $O/r02_configs_b.err; cut -c1-200 $O/r02_configs_b.jsonl
