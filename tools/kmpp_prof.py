"""Per-kernel times of one k-means++ round on the bench data (1M x 128)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import spfresh_b200 as s  # noqa: E402

rows = bench.make_rows(0)
ctx = s.Context(0)
ds = s.Dataset(ctx, rows)
sess = ds.kmeanspp(0, 12345)
ctx.set_profiling(True)
for i in range(4):
    r = sess.round(0.37 + 0.1 * i)
    print(i, r, {n: round(ctx.kernel_ms(n), 4) for n in ("kmpp_update", "kmpp_sum", "kmpp_pick")}, flush=True)
sess.free()
