#!/bin/bash
# round 2: full ncu capture of classify_kernel and finalize_kernel of one bench step
mkdir -p gpurun_out
python tools/step_once.py gauss 3 > gpurun_out/plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"classify_kernel|finalize_kernel" -s 4 -c 2 -o gpurun_out/r02_resolve python tools/step_once.py gauss 3 > gpurun_out/ncu_resolve.log 2>&1
echo "rc=$?"; ls -la gpurun_out/r02_resolve.ncu-rep
