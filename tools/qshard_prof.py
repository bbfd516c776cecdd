"""Scan passes of ONE rank of an N-way list-sharded index (emulated on one GPU): the index holds the
lists [0, nlists / N) of the config-2 assignment, the batch is the whole 100 k-query sweep batch.
usage: python tools/qshard_prof.py [N] [nprobe] [nq]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import spfresh_b200 as s  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nprobe = int(sys.argv[2]) if len(sys.argv) > 2 else 8
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
rows = bench.make_rows(0)
ctx = s.Context(0)
ds = s.Dataset(ctx, rows)
cent = np.arange(bench.K_CENT, dtype=np.uint64)
res = ds.assign(0, cent)
f = res.fetch(best=False, dmin=False)
med = ds.update_medoids_from(0, res, cent)
res.free()
idx = s.DeviceIndex.pack(ds, f.offsets, f.members, med, list_range=(0, bench.K_CENT // N))
q = bench.make_queries(nq)
for prof in (False, True):
    ctx.set_profiling(prof)
    for i in range(3):
        t0 = time.perf_counter()
        ids, dists, counts = idx.search(q, 10, nprobe)
        dt = (time.perf_counter() - t0) * 1e3
        line = f"profiling={prof} rep {i}: call {dt:.3f} ms"
        if prof:
            line += f" scan {ctx.kernel_ms('scan'):.3f} probe {ctx.kernel_ms('probe'):.3f} | " + " ".join(
                f"{n} {ctx.kernel_ms('scan_tc_' + n):.3f}" for n in ("gather", "a", "tau", "b", "flag", "refine", "fallback", "units", "groups", "candidates"))
        print(line, flush=True)
