#!/bin/bash
# fourth session: default bench line with its wall time; full ncu captures of the update_centroids kernels
mkdir -p gpurun_out
( time timeout 600 python bench.py > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err ) 2>&1 | grep real
echo "bench rc=$?"; tail -c 300 gpurun_out/bench_v7.err; cut -c1-200 gpurun_out/bench_v7.json
timeout 120 python tools/update_prof.py > gpurun_out/plain_update.log 2>&1 || { tail -3 gpurun_out/plain_update.log; exit 1; }
cat gpurun_out/plain_update.log
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"cluster_sum_ws_kernel|medoid_kernel" -c 3 -o gpurun_out/r02_update_v1 python tools/update_prof.py > gpurun_out/ncu_update.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_update.log; ls -la gpurun_out/*.ncu-rep
