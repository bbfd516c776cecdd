#!/bin/bash
# two-rank bench (driver form) + the two-GPU NCCL parity test
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "rc=$?"; tail -c 400 gpurun_out/bench_2gpu.err
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus" 2>&1 | tail -2
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_2gpu.json') if l.startswith('{')][0])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'parity',d.get('parity_ok'))
print('sharded_check', d['sharded_check'])
k=d['sharded_kmeans_iteration']; print('kmeans', k['ms_per_iteration'], k['collective_ms'])
c=d['configs']
for n,v in c['sweep']['points'].items(): print(n, round(v['qps_e2e']), v['scan_ms_max'], v['exchange_ms_max'], v.get('replicated_lists',{}).get('qps_e2e'), v.get('replicated_lists',{}).get('identical_to_list_sharded'), v.get('parity_ok'))
print('deep', c['deep_strong']['kmeans_iteration_ms'])
PY
