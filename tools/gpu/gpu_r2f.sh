#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kmeans or config2 or tensor or small_buffers or overflow or lire or host" 2>&1 | tail -3
timeout 900 python tools/kmeans_prof.py 100000000 2> gpurun_out/kmprof.err | cut -c1-900 | tee gpurun_out/kmprof.jsonl
tail -3 gpurun_out/kmprof.err
