#!/bin/bash
# the bound pass's two launches of the last query call, each measured alone under ncu (they overlap in a normal run)
mkdir -p gpurun_out
python tools/query_wall.py 10000 10 1 > gpurun_out/plain_qwall.log 2>&1 || { tail -3 gpurun_out/plain_qwall.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 12 -c 4 -o gpurun_out/r02_scan_tc_split_v1 python tools/query_wall.py 10000 10 1 > gpurun_out/ncu_scan_split.log 2>&1
echo "scan rc=$?"; tail -2 gpurun_out/ncu_scan_split.log; ls -la gpurun_out/r02_scan_tc_split_v1.ncu-rep
