#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 --no-query --no-configs > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
echo "rc=$?"; tail -c 300 gpurun_out/bench_a.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_a.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], 'parity', d.get('parity_ok'), d.get('parity_checked_rows'))
print('kernels', {k:round(v,3) for k,v in d['kernels_ms'].items()})
print('kmeans', d['kmeans_iteration']['ms_per_iteration'] if d.get('kmeans_iteration') else None)
PY
