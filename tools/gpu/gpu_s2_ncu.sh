#!/bin/bash
# round-2 (second session) captures: launch list of 64 k-means++ rounds, full capture of the cluster scan and the pick
mkdir -p gpurun_out
python tools/kmpp_rounds_once.py > gpurun_out/kmpp_once.log 2>&1 || { cat gpurun_out/kmpp_once.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_kmpp.csv python tools/kmpp_rounds_once.py > gpurun_out/ncu_kmpp_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"seq_sum_cluster_kernel|kmpp_pick_kernel|kmpp_update_tiled_kernel" -s 30 -c 6 -o gpurun_out/r02_kmpp python tools/kmpp_rounds_once.py > gpurun_out/ncu_kmpp_full.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/kmpp_once.log; ls -la gpurun_out/r02_kmpp.ncu-rep gpurun_out/r02_launches_kmpp.csv
