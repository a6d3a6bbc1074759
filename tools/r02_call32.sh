#!/bin/bash
# round 2, last validation: full -m gpu suite on the final build (user-functor library rebuilt on the final headers), compute-sanitizer on the two-warp logistic kernel
cd /root/repo
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
S=$O/r02_sanitizer_logistic.txt; rm -f $S
for tool in memcheck racecheck synccheck; do
  echo "## $tool" >> $S
  timeout 400 compute-sanitizer --tool $tool python tools/sanitize_logistic.py 2>&1 | grep -v "^$" | tail -8 >> $S
done
cat $S
