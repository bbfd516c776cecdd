#!/bin/bash
# bench (plain) at N = 1 + the reference arm; output under gpurun_out/
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "rc=$?"; tail -c 400 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-400
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
