"""Minimal driver for profiling: a few assign calls on one shape (default 1M x 128, k = 4096)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import spfresh_b200 as s  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
kind = sys.argv[3] if len(sys.argv) > 3 else "gauss"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
d = 128
g = np.random.Generator(np.random.Philox(key=42))
if kind == "gauss":
    data = g.standard_normal((n, d), dtype=np.float32)
else:
    cen = 2.0 * g.standard_normal((1024, d), dtype=np.float32)
    data = cen[g.integers(0, 1024, n)] + 0.5 * g.standard_normal((n, d), dtype=np.float32)
cent = np.random.Generator(np.random.Philox(key=7)).choice(n, k, replace=False)
ctx = s.Context.default()
ctx.set_profiling(True)
ds = s.Dataset(ctx, data)
for i in range(reps):
    r = ds.assign(0, cent)
    print(f"rep {i}: assign_tc {ctx.kernel_ms('assign_tc'):.3f} ms resolve {ctx.kernel_ms('resolve'):.3f} "
          f"[classify {ctx.kernel_ms('classify'):.3f} eval {ctx.kernel_ms('exact_eval'):.3f} finalize {ctx.kernel_ms('finalize'):.3f} "
          f"ovf {ctx.kernel_ms('overflow'):.3f}] "
          f"cc {ctx.kernel_ms('cc_matrix'):.3f} csr {ctx.kernel_ms('csr'):.3f} total members {r.total} "
          f"overflow rows {ctx.last_overflow_rows()}", flush=True)
    r.free()
