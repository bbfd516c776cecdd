#!/bin/bash
mkdir -p gpurun_out
for sp in 76 84 92 100; do
  echo "== scan_tc_split=$sp"
  SCAN_TC_SPLIT=$sp timeout 300 python tools/query_prof.py 10000 10 4 1 1 gauss 2>&1 | grep -v "^$" | tail -3
done
echo "== clustered"
for sp in 0 1 80; do
  echo "== scan_tc_split=$sp"
  SCAN_TC_SPLIT=$sp timeout 300 python tools/query_prof.py 10000 10 4 1 1 clustered 2>&1 | grep -v "^$" | tail -3
done
echo "== 30k queries"
for sp in 0 1 80 100; do
  echo "== scan_tc_split=$sp"
  SCAN_TC_SPLIT=$sp timeout 300 python tools/query_prof.py 30000 10 4 1 1 gauss 2>&1 | grep -v "^$" | tail -3
done
