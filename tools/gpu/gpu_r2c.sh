#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "search or query or scan or sweep or config5 or probe or spann or kmeans or config2" 2>&1 | tail -4
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "rc=$?"; tail -c 300 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], 'parity', d.get('parity_ok'))
print('kernels', {k:round(v,3) for k,v in d['kernels_ms'].items()})
q=d['query']; print('query', q['qps_e2e'], q['scan_ms'], q['probe_ms'], q['scan']['passes_ms'], q['scan']['frac'])
print('kmeans', d['kmeans_iteration']['ms_per_iteration'] if d.get('kmeans_iteration') else None)
c=d['configs']
for k,v in c['sweep']['points'].items(): print(k, round(v['qps_e2e']), v['scan_ms_max'], v.get('parity_ok'))
print('deep', c['deep_strong']['kmeans_iteration_ms'], c['deep_strong']['rank0_kernels_ms'])
for k in ('gist_manhattan','gist_chebyshev'): print(k, c[k]['assign_exact_kernel_ms'], c[k]['roofline']['frac_at_measured_clock'], c[k]['parity_ok'])
PY
