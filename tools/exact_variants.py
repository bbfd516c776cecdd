"""Direct-form kernel variants at the GIST row length (d = 960, k = 4096, clustered rows): the 4 x 4
kernel, the TMA-staged 8 x 8 kernel, and its packed-subtract (FADD2) form, for Manhattan and
Chebyshev.  Every variant must return the same assignment bit for bit."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import spfresh_b200 as s  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
dev = torch.device("cuda", 0)
x = bench.device_clustered(torch, dev, n, 960, 1024, 45, 46)
ctx = s.Context(0)
ds = s.Dataset(ctx, device_ptr=x.data_ptr(), n=n, d=960)
cent = np.random.Generator(np.random.Philox(key=7)).choice(n, 4096, replace=False).astype(np.uint64)
ctx.set_profiling(True)
for metric, name in ((s.METRIC_MANHATTAN, "manhattan"), (s.METRIC_CHEBYSHEV, "chebyshev")):
    ref = None
    for tag, tma, packed, minb in (("4x4", 0, 0, 2), ("tma 8x8", 7, 0, 2), ("tma 8x8 packed", 7, 7, 2), ("tma 8x8 packed 1 CTA/SM", 7, 7, 1), ("tma 8x8 packed 64 rows 3 CTA/SM", 7, 7, 3)):
        ctx.set_param("exact_tma", tma)
        ctx.set_param("exact_packed", packed)
        ctx.set_param("exact_one_cta", 7 if minb == 1 else 0)
        ctx.set_param("exact_three_cta", 7 if minb == 3 else 0)
        ctx.set_param("exact_seed", 0)
        r = ds.assign(metric, cent)
        ms = ctx.kernel_ms("assign_exact")
        f = r.fetch()
        r.free()
        same = None
        if ref is None:
            ref = f
        else:
            same = bool(np.array_equal(ref.best, f.best) and np.array_equal(ref.dmin.view(np.uint32), f.dmin.view(np.uint32))
                        and np.array_equal(ref.offsets, f.offsets) and np.array_equal(ref.members, f.members))
        lane = 2.0 * n * 4096 * 960
        print(f"{name:10s} {tag:33s} assign_exact {ms:8.2f} ms = {lane / (ms * 1e-3) / (148 * 128 * 1.965e9):.3f} of 2 lane-instr/element-op at 1965 MHz"
              f"  identical={same}", flush=True)
