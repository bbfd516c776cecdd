#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_cpp.py -m gpu -x -q -k "medoid or kmeans or fit or sharded or update or host" 2>&1 | tail -4
python tools/_mean_micro.py 2>/dev/null
timeout 600 python tools/kmeans_prof.py 25000000 2>gpurun_out/kmprof.err | tee gpurun_out/kmprof.jsonl
tail -3 gpurun_out/kmprof.err
