#!/bin/bash
# round 2: ncu evidence — full capture of assign_tc_kernel and of the L1 / Linf direct-form kernel,
# launch list of the bench step
mkdir -p gpurun_out
python tools/step_once.py gauss 3 > gpurun_out/plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:assign_tc_kernel -s 2 -c 1 -o gpurun_out/r02_assign_tc python tools/step_once.py gauss 3 > gpurun_out/ncu_assign_tc.log 2>&1
echo "assign_tc rc=$?"
python tools/gist_once.py 200000 > gpurun_out/plain_gist.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:assign_exact_kernel -c 6 -o gpurun_out/r02_assign_exact python tools/gist_once.py 200000 > gpurun_out/ncu_assign_exact.log 2>&1
echo "assign_exact rc=$?"
python bench.py --steps 2 --warmup 3 --no-query --no-cpu --no-configs > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-query --no-cpu --no-configs > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_bench.csv
