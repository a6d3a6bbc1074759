#!/bin/bash
cd /root/repo
O=gpurun_out; mkdir -p $O
rm -f $O/r02_call14.log
for L in 3 0; do
  export CNO_NEWTON_DMMA_LAYOUT=$L
  echo "== layout $L" >> $O/r02_call14.log
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core" 2>&1 | tail -1 >> $O/r02_call14.log
  python bench_configs.py c5t 2>&1 | cut -c100-200 >> $O/r02_call14.log
done
unset CNO_NEWTON_DMMA_LAYOUT
for tool in memcheck racecheck; do
  echo "== compute-sanitizer $tool" >> $O/r02_call14.log
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_newton_dmma.py 2>&1 | tail -6 >> $O/r02_call14.log
done
cat $O/r02_call14.log
