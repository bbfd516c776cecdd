"""Where a k-means update goes (1M x 128, k = 4096): python tools/update_prof.py [knob=value ...]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import spfresh_b200 as s  # noqa: E402

rows = bench.make_rows(0)
ctx = s.Context(0)
for arg in sys.argv[1:]:                      # knobs: python tools/update_prof.py medoid_direct=0
    name, val = arg.split("=")
    ctx.set_param(name, int(val))
ds = s.Dataset(ctx, rows)
cent = np.arange(bench.K_CENT, dtype=np.uint64)
res = ds.assign(0, cent)


def t(f, n=3):
    best = 1e9
    out = None
    for _ in range(n):
        t0 = time.perf_counter()
        out = f()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3, out


ms, _ = t(lambda: ds.update_medoids_from(0, res, cent))
print(f"update_medoids_from      {ms:8.2f} ms")
ms, (sums, counts) = t(lambda: ds.cluster_sums(res))
print(f"cluster_sums             {ms:8.2f} ms")
means = (sums / np.maximum(counts, 1).astype(np.float32)[:, None]).astype(np.float32)
ms, _ = t(lambda: ds.medoid_candidates(0, res, means))
print(f"medoid_candidates        {ms:8.2f} ms")
ms, _ = t(lambda: ds.fetch_rows(cent))
print(f"fetch_rows(k)            {ms:8.2f} ms")
ms, r2 = t(lambda: ds.assign_vectors(0, rows[:bench.K_CENT]), 2)
print(f"assign_vectors           {ms:8.2f} ms")
ctx.set_profiling(True)
ds.update_medoids_from(0, res, cent)
print("kernels: cluster_mean %.2f ms, medoid %.2f ms" % (ctx.kernel_ms("cluster_mean"), ctx.kernel_ms("medoid")))
sizes = np.diff(res.fetch(best=False, dmin=False).offsets.astype(np.int64))
print("cluster sizes: mean %.0f max %d" % (sizes.mean(), sizes.max()))
