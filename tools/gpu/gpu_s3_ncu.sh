#!/bin/bash
# round-2 (third session) captures: launch list of the bench command, full captures of finalize_kernel and of the
# bound pass's two launches (under ncu they run one after the other, each measured alone)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/plain_bench.json 2> gpurun_out/plain_bench.err || { tail -5 gpurun_out/plain_bench.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bench_v3.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/ncu_bench_list.log 2>&1
echo "list rc=$?"
python tools/step_once.py gauss 3 > gpurun_out/plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:finalize_kernel -s 2 -c 1 -o gpurun_out/r02_finalize_v1 python tools/step_once.py gauss 3 > gpurun_out/ncu_finalize.log 2>&1
echo "finalize rc=$?"
python tools/query_wall.py 10000 10 2 > gpurun_out/plain_qwall.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"scan_tc_kernel<0" -s 8 -c 2 -o gpurun_out/r02_scan_tc_split_v1 python tools/query_wall.py 10000 10 1 > gpurun_out/ncu_scan_split.log 2>&1
echo "scan rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_bench_v3.csv
