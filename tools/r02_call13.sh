#!/bin/bash
cd /root/repo
O=gpurun_out; mkdir -p $O
rm -f $O/r02_call13.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "newton or condition or division" 2>&1 | tail -3 >> $O/r02_call13.log
for L in 3 4 0 1; do
  export CNO_NEWTON_DMMA_LAYOUT=$L
  echo "== layout $L" >> $O/r02_call13.log
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core" 2>&1 | tail -1 >> $O/r02_call13.log
  python bench_configs.py c5t 2>&1 | cut -c100-200 >> $O/r02_call13.log
done
cat $O/r02_call13.log
