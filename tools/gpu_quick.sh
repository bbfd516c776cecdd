#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/assign_variants.py 10 2> gpurun_out/variants.err | tee gpurun_out/variants.jsonl
tail -3 gpurun_out/variants.err
