#!/bin/bash
# round 2, GPU call 18 (--gpus N): bench.py under torchrun at N GPUs: the weak record + the strong-scaling sub-record (global 2^20 split)
cd /root/repo
N=${1:-8}
O=gpurun_out; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 > $O/r02_bench_${N}gpu.json 2> $O/r02_bench_${N}gpu.err
tail -c 1800 $O/r02_bench_${N}gpu.json
