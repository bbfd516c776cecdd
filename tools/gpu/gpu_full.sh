#!/bin/bash
# full GPU suite + bench (N = 1) + reference arm + smoke
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], 'parity', d.get('parity_ok'), d.get('parity_checked_rows'), 'roof', d['roofline']['frac'])
print('kernels', {k:round(v,3) for k,v in d['kernels_ms'].items()})
q=d['query']; print('query', q['qps_e2e'], q['scan_ms'], q['probe_ms'], q['scan']['passes_ms'], q['scan']['frac'])
for k in d:
    if 'kmeans' in k: print(k, d[k].get('ms_per_iteration'))
c=d['configs']
for k,v in c['sweep']['points'].items(): print(k, round(v['qps_e2e']), v['scan_ms_max'], v.get('parity_ok'))
print('deep', c['deep_strong']['kmeans_iteration_ms'], {k:round(v,2) for k,v in c['deep_strong']['rank0_kernels_ms'].items()})
for k in ('gist_manhattan','gist_chebyshev'): print(k, c[k]['assign_exact_kernel_ms'], c[k]['roofline']['frac_at_measured_clock'], c[k]['parity_ok'])
PY
