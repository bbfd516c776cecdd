#!/bin/bash
# last call of the round: the GPU suite without the 100 k-query sweep cases (unchanged code, 65 s), then a short bench line
mkdir -p gpurun_out
timeout 50 python -m pytest tests -m gpu -x -q -k "not config5" 2>&1 | tail -2 | tee gpurun_out/final_gputests3.log
timeout 28 python bench.py --no-configs --no-cpu > gpurun_out/bench_v9_short.json 2> gpurun_out/bench_v9_short.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_v9_short.json'))
k=d['sharded_kmeans_iteration']; print('value',d['value'],'kmeans', k['ms_per_iteration'], {a:round(b,3) for a,b in k['kernels_ms'].items()})
PY
