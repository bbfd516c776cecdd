#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "search or query or scan or sweep or config5 or probe or spann or tensor or config2 or exact_matches or overflow" 2>&1 | tail -4
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
echo "rc=$?"; tail -c 300 gpurun_out/bench_q.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_q.json'))
q=d['query']; print('query', q['qps_e2e'], q['scan_ms'], q['probe_ms'], q['scan']['passes_ms'], q['scan']['frac'], q['scan'].get('subgroups_refined_per_query'))
PY
