#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kmeans or medoid or mean or config2_assign or update" 2>&1 | tail -4
timeout 900 python tools/kmeans_prof.py 100000000 2> gpurun_out/kmprof.err | tee gpurun_out/kmprof.jsonl
tail -3 gpurun_out/kmprof.err
