"""One Manhattan and one Chebyshev assign at the GIST row length (d = 960, k = 4096) on a 200k-row
clustered shard: the launch profiled for the FP32-pipe evidence of assign_exact_kernel<1> / <2>."""
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
for metric in (s.METRIC_MANHATTAN, s.METRIC_CHEBYSHEV, s.METRIC_MANHATTAN, s.METRIC_CHEBYSHEV):
    r = ds.assign(metric, cent)
    ms = ctx.kernel_ms("assign_exact")
    lane = 2.0 * n * 4096 * 960
    print("metric", metric, "members", r.total, f"assign_exact {ms:.2f} ms = {lane / (ms * 1e-3) / (148 * 128 * 1.965e9):.3f} of the FP32 issue peak at 1965 MHz",
          file=sys.stderr, flush=True)
    r.free()
