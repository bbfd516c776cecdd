#!/bin/bash
# medoid_kernel staging variants: parity first, then A/B of the update path, then the whole GPU suite
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "medoid or sharded_build or kmeans" 2>&1 | tail -3
for v in 0 1 0 1; do echo "medoid_direct=$v"; timeout 100 python tools/update_prof.py medoid_direct=$v 2>&1 | grep -E "update_medoids_from|medoid_candidates|kernels:"; done | tee gpurun_out/medoid_ab.log
