#!/bin/bash
# round 2, GPU call 16 (--gpus 2): bench.py under torchrun at N = 2 (weak record + the strong-scaling sub-record), the reference arm at N = 2,
# and the NCCL C++ program
cd /root/repo
O=gpurun_out; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r02_bench_2gpu.json 2> $O/r02_bench_2gpu.err
tail -c 2500 $O/r02_bench_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/r02_bench_2gpu_reference.json 2> $O/r02_bench_2gpu_reference.err
tail -c 800 $O/r02_bench_2gpu_reference.json
timeout 300 tests/cpp/build/sharded_nccl 2>&1 | tail -3
