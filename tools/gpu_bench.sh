#!/bin/bash
# GPU round: parity tests, smoke, bench (plain), then the ncu launch list of the same bench command.
mkdir -p gpurun_out
echo "== gpu tests"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "rc=$?"; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "rc=$?"; cat gpurun_out/bench_ref.json
if [ "$1" == "ncu" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-query --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
fi
