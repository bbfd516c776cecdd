"""Per-kernel times of one k-means++ round on the bench data (1M x 128)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import spfresh_b200 as s  # noqa: E402

rows = bench.make_rows(0)
if len(sys.argv) > 1 and sys.argv[1] == "clustered":
    g = np.random.Generator(np.random.Philox(key=44))
    cen = 2.0 * g.standard_normal((1024, bench.DIM), dtype=np.float32)
    rows = (cen[g.integers(0, 1024, rows.shape[0])] + 0.5 * rows).astype(np.float32)
ctx = s.Context(0)
ds = s.Dataset(ctx, rows)
sess = ds.kmeanspp(0, 12345)
ctx.set_profiling(True)
import time  # noqa: E402
prev = {n: 0.0 for n in ("kmpp_update", "kmpp_sum", "kmpp_pick")}
for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 4):
    t0 = time.perf_counter()
    r = sess.round((0.37 + 0.1 * i) % 1.0)
    dt = (time.perf_counter() - t0) * 1e3
    cur = {n: ctx.kernel_ms(n) for n in prev}
    if i < 4 or i % 200 == 0:
        print(i, r, {n: round(cur[n] - prev[n], 4) for n in prev}, f"call {dt:.3f} ms", flush=True)
    prev = cur
sess.free()

# batched rounds (spf_kmpp_rounds): no host round trip per round
ctx.set_profiling(False)
sess = ds.kmeanspp(0, 12345)
u = np.random.Generator(np.random.Philox(key=9)).random(1024)
sess.rounds(u[:8])
for cnt in (64, 256, 512):
    t0 = time.perf_counter()
    rows_, failed = sess.rounds(u[:cnt])
    dt = (time.perf_counter() - t0) * 1e3
    print(f"batch of {cnt}: {dt:.2f} ms = {dt / cnt:.4f} ms per round, failed={failed}", flush=True)
sess.free()
import ctypes as C  # noqa: E402
for mode, name in ((1, "single-CTA scan"), (3, "cluster scan"), (2, "serial chain")):
    v = np.random.default_rng(1).random(1_000_000, dtype=np.float32) * 300
    ctx.set_profiling(True)
    ctx.seq_sum_f32(v, mode)
    ctx.seq_sum_f32(v, mode)
    print(name, "1M elements:", round(ctx.kernel_ms("seq_sum"), 4), "ms", flush=True)
