#!/bin/bash
# whole GPU suite on the final tree, then the default bench line
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/final_gputests2.log
timeout 300 python bench.py > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err; echo "bench rc=$?"; tail -c 200 gpurun_out/bench_v8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_v8.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'parity',d['parity_ok'],'wall',d['wall_s'])
k=d['sharded_kmeans_iteration']; print('kmeans', k['ms_per_iteration'], k['kernels_ms'])
print('deep', d['configs']['deep_strong']['kmeans_iteration_ms'], 'query', d['query']['qps_e2e'])
PY
