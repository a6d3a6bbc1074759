#!/bin/bash
# round 2: loop-size knobs of the logistic functor (gradient round size, margin batch), one build each
cd /root/repo
O=gpurun_out; mkdir -p $O
L=$O/r02_call30.log; rm -f $L
for V in rf16nf16 rf16; do
  cp cppnumericalsolvers_b200/variants/libcno_$V.so cppnumericalsolvers_b200/libcno.so
  echo "== $V" >> $L
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "logistic" 2>&1 | tail -1 >> $L
  timeout 300 python bench_configs.py c3 2>&1 | cut -c48-175 >> $L
done
cat $L
