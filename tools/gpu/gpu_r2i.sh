#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/pipe_variants.py 10 0,1,0,1 2>gpurun_out/pipe.err | tee gpurun_out/pipe_variants.jsonl
SPF_TC_PIPE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor or config2 or kmeans_session or balanced or fit" 2>&1 | tail -3
