#!/bin/bash
# compute_mean producer variants: parity, then A/B of the update path
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cluster_mean or medoid or sharded_build or kmeans_session or kmeans_balanced" 2>&1 | tail -3
for v in 0 1; do echo "sum_fast=$v"; timeout 60 python tools/update_prof.py sum_fast=$v 2>&1 | grep -E "update_medoids_from|cluster_sums|kernels:"; done | tee gpurun_out/sum_fast_ab.log
