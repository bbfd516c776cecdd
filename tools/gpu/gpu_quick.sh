#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor or small_buffers or overflow or config2_assign or exact_matches or kmeans_session_single" 2>&1 | tail -4
timeout 600 python tools/assign_variants.py 10 2> gpurun_out/variants.err | tee gpurun_out/variants.jsonl
tail -3 gpurun_out/variants.err
