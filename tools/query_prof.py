"""Query-path driver for profiling: index over a 1M x 128 assignment, 10k-query batches.
usage: python tools/query_prof.py [nq] [nprobe] [reps] [mode]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import spfresh_b200 as s  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nprobe = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 1
rows = bench.make_rows(0)
ctx = s.Context(0)
ctx.set_param("scan_list_major", mode)
ds = s.Dataset(ctx, rows)
cent = np.arange(bench.K_CENT, dtype=np.uint64)
res = ds.assign(0, cent)
f = res.fetch(best=False, dmin=False)
med = ds.update_medoids_from(0, res, cent)
res.free()
idx = s.DeviceIndex.pack(ds, f.offsets, f.members, med)
q = bench.make_queries(nq)
ctx.set_profiling(True)
for i in range(reps):
    ids, dists, counts = idx.search(q, 10, nprobe)
    b = idx.last_scan_bytes()
    print(f"rep {i}: scan {ctx.kernel_ms('scan'):.3f} ms probe {ctx.kernel_ms('probe'):.3f} ms "
          f"algorithmic {b / 1e9:.1f} GB -> {b / ctx.kernel_ms('scan') / 1e6:.0f} GB/s, mean count {counts.mean():.2f}", flush=True)
