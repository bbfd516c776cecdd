#!/bin/bash
# full ncu capture of the row-staged medoid_kernel (after the plain run has exited 0)
mkdir -p gpurun_out
timeout 60 python tools/update_prof.py > gpurun_out/plain_update2.log 2>&1 || { tail -3 gpurun_out/plain_update2.log; exit 1; }
grep kernels gpurun_out/plain_update2.log
timeout 100 ncu --set full --clock-control none --import-source on -k regex:"medoid_kernel" -c 1 -o gpurun_out/r02_medoid_v2 python tools/update_prof.py > gpurun_out/ncu_medoid.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r02_medoid_v2.ncu-rep
