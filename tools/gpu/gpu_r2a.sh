#!/bin/bash
# round 2, call A: parity tests (incl. BASELINE-size cases), kernel variants, bench
mkdir -p gpurun_out
echo "== gpu tests"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== variants"
timeout 600 python tools/assign_variants.py 10 2> gpurun_out/variants.err | tee gpurun_out/variants.jsonl
tail -3 gpurun_out/variants.err
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "rc=$?"; tail -c 7000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
