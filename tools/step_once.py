"""A few bench steps (device-resident assign, 1M x 128, k=4096) for launch-list profiling."""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import spfresh_b200 as s  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "gauss"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = bench.make_rows(0)
if kind == "clustered":
    g = np.random.Generator(np.random.Philox(key=44))
    cen = 2.0 * g.standard_normal((1024, bench.DIM), dtype=np.float32)
    rows = (cen[g.integers(0, 1024, rows.shape[0])] + 0.5 * rows).astype(np.float32)
ctx = s.Context(0)
ds = s.Dataset(ctx, rows)
cent = np.arange(bench.K_CENT, dtype=np.uint64)
for i in range(steps):
    r = ds.assign(0, cent)
    print("step", i, "members", r.total, flush=True)
    r.free()
