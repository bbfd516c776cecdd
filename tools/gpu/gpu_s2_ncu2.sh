#!/bin/bash
# round 2, second session: full captures of the two tcgen05 kernels at their final state + launch list of the bench step
mkdir -p gpurun_out
python tools/step_once.py gauss 3 > gpurun_out/plain_step.log 2>&1 || { tail gpurun_out/plain_step.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"assign_tc_kernel" -s 2 -c 1 -o gpurun_out/r02_assign_tc_v2 python tools/step_once.py gauss 3 > gpurun_out/ncu_assign.log 2>&1
python tools/query_prof.py 10000 10 2 > gpurun_out/plain_query.log 2>&1 || { tail gpurun_out/plain_query.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"scan_tc_kernel" -s 3 -c 1 -o gpurun_out/r02_scan_tc_v3 python tools/query_prof.py 10000 10 2 > gpurun_out/ncu_scan.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_v2.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs > gpurun_out/ncu_bench_list.log 2>&1
echo "rc=$?"; ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_bench_v2.csv; tail -2 gpurun_out/plain_query.log
