#!/bin/bash
mkdir -p gpurun_out
python tools/gist_once.py 100000 > gpurun_out/plain_gist.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:assign_exact_tma_kernel -c 4 -o gpurun_out/r02_assign_exact_tma python tools/gist_once.py 100000 > gpurun_out/ncu_assign_exact.log 2>&1
echo "rc=$?"; ls -la gpurun_out/r02_assign_exact_tma.ncu-rep
