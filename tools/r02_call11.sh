#!/bin/bash
# round 2: tensor-core NewtonDescent, warp populations (smem / TMEM / mixed): parity + timing; old kernel beside them
cd /root/repo
O=gpurun_out; mkdir -p $O
rm -f $O/r02_call11.log
for L in ${LAYOUTS:-1 3 0}; do
  export CNO_NEWTON_DMMA_LAYOUT=$L
  echo "== layout $L" >> $O/r02_call11.log
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core" 2>&1 | tail -1 >> $O/r02_call11.log
  python bench_configs.py c5t 2>&1 | cut -c100-200 >> $O/r02_call11.log
done
echo "== old kernel (default policy)" >> $O/r02_call11.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_full_size_gpu.py -m gpu -x -q -k "newton" 2>&1 | tail -1 >> $O/r02_call11.log
python bench_configs.py c5 2>&1 | cut -c1-160 >> $O/r02_call11.log
cat $O/r02_call11.log
