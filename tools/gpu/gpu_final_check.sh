#!/bin/bash
# fresh-checkout insurance: smoke + the GPU parity suite on the tree as rebuilt from git
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout 520 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/final_gputests.log 2>&1; echo "pytest rc=$?"; tail -20 gpurun_out/final_gputests.log
