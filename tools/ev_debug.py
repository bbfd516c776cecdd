"""Timing experiment for exact_eval (debug switches change results: timing only)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import spfresh_b200 as s
n, k, d = 1_000_000, 4096, 128
g = np.random.Generator(np.random.Philox(key=42))
data = g.standard_normal((n, d), dtype=np.float32)
cent = np.random.Generator(np.random.Philox(key=7)).choice(n, k, replace=False)
ctx = s.Context.default()
ctx.set_profiling(True)
ds = s.Dataset(ctx, data)
for dbg in (0, 0, 1, 2, 4, 6, 7):
    ctx.set_param("debug", dbg)
    r = ds.assign(0, cent)
    print(f"debug {dbg}: exact_eval {ctx.kernel_ms('exact_eval'):.3f} ms resolve {ctx.kernel_ms('resolve'):.3f}", flush=True)
    r.free()
