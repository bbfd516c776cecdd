#!/bin/bash
# First GPU contact: exact path parity, then the tcgen05 path in its own process.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== exact-path parity (SPF_FORCE_EXACT=1)"
SPF_FORCE_EXACT=1 timeout 900 python -m pytest tests -m gpu -x -q -k "not tensor" 2>&1 | tail -25 | tee gpurun_out/pytest_exact.log
echo "== tcgen05 diagnostics"
timeout 300 python tools/tc_debug.py 2>&1 | tail -20 | tee gpurun_out/tc_debug.log
echo "== full gpu suite"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_full.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
