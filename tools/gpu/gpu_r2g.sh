#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 5 --warmup 3 --no-query --no-cpu > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
echo "rc=$?"; tail -c 300 gpurun_out/bench_c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
c=d['configs']
print('deep', c['deep_strong']['kmeans_iteration_ms'], c['deep_strong']['iteration_ms_all'], {k:round(v,2) for k,v in c['deep_strong']['rank0_kernels_ms'].items()})
for k in ('gist_manhattan','gist_chebyshev'): print(k, c[k]['assign_exact_kernel_ms'], c[k]['call_ms'], c[k]['other_kernels_ms'])
PY
