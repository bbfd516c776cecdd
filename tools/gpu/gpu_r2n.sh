#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/pipe_variants.py 10 32,16,8,16,8 finalize_lanes 2>gpurun_out/pipe.err
