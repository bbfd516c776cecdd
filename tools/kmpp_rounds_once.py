"""64 k-means++ rounds on the bench data through spf_kmpp_rounds (profiling driver for ncu)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import spfresh_b200 as s  # noqa: E402

rows = bench.make_rows(0)
ctx = s.Context(0)
ds = s.Dataset(ctx, rows)
sess = ds.kmeanspp(0, 12345)
u = np.random.Generator(np.random.Philox(key=9)).random(64)
picked, failed = sess.rounds(u)
print("rounds", len(picked), "failed", failed, "last", int(picked[-1]))
sess.free()
