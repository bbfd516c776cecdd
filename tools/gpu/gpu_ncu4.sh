#!/bin/bash
mkdir -p gpurun_out
python tools/exact_variants.py 100000 > gpurun_out/plain_variants.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:assign_exact_tma_kernel -c 6 -o gpurun_out/r02_exact_packed python tools/exact_variants.py 100000 > gpurun_out/ncu_variants.log 2>&1
echo "rc=$?"; cat gpurun_out/plain_variants.log | tail -8; ls -la gpurun_out/r02_exact_packed.ncu-rep
