"""Wall time per search call of the bench batch (10 k queries, top-10, nprobe 10), profiling off.
usage: python tools/query_wall.py [nq] [nprobe] [reps] [param=value ...]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import spfresh_b200 as s  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nprobe = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
rows = bench.make_rows(0)
ctx = s.Context(0)
for kv in sys.argv[4:]:                       # name=value context parameters
    name, value = kv.split("=")
    ctx.set_param(name, int(value))
ds = s.Dataset(ctx, rows)
cent = np.arange(bench.K_CENT, dtype=np.uint64)
res = ds.assign(0, cent)
f = res.fetch(best=False, dmin=False)
med = ds.update_medoids_from(0, res, cent)
res.free()
idx = s.DeviceIndex.pack(ds, f.offsets, f.members, med)
qp = torch.empty((nq, bench.DIM), dtype=torch.float32, pin_memory=True)
qp.numpy()[:] = bench.make_queries(nq)
q = qp.numpy()
out = (torch.empty((nq, 10), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64),
       torch.empty((nq, 10), dtype=torch.float32, pin_memory=True).numpy(),
       torch.empty(nq, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32))
for _ in range(3):
    idx.search(q, 10, nprobe, out=out)
ts = []
for _ in range(reps):
    t0 = time.perf_counter()
    idx.search(q, 10, nprobe, out=out)
    ts.append((time.perf_counter() - t0) * 1e3)
print(f"wall per call: median {np.median(ts):.3f} ms, min {min(ts):.3f} ms -> {nq / np.median(ts) / 1e3:.2f} M QPS")
