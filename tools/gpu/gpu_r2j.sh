#!/bin/bash
mkdir -p gpurun_out
for sp in 0 1 40 70; do
  echo "== scan_tc_split=$sp"
  SCAN_TC_SPLIT=$sp timeout 300 python tools/query_prof.py 10000 10 4 1 1 gauss check 2>&1 | grep -v "^$" | tail -6
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "search or config5 or index or lire" 2>&1 | tail -3
