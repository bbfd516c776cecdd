"""Times the device-resident assign step (1M x 128, k = 4096, N(0,1)) under the kernel variants the
knobs select — epilogue layout of the tcgen05 kernel (8 / 16 epilogue warps), counting-sort CSR vs
library sort, cached centroid matrix — and checks that all variants return identical results.
Prints one JSON line per variant (round-2 kernel experiments, profiles/)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import spfresh_b200 as s  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
rows = bench.make_rows(0)
cent = np.arange(bench.K_CENT, dtype=np.uint64)
kn = ["assign_tc", "classify", "exact_eval", "finalize", "overflow", "cc_matrix", "csr"]
ref = None
for split, csr_sort, cc_cache in [(2, 0, 0), (2, 1, 1), (4, 0, 0), (4, 1, 1), (4, 1, 0)]:
    ctx = s.Context(0)
    ctx.set_param("tc_epi_split", split)
    ctx.set_param("csr_sort", csr_sort)
    ctx.set_param("cc_cache", cc_cache)
    ds = s.Dataset(ctx, rows)
    ext = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(3):
        ds.assign(0, cent).free()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(steps):
        ds.assign(0, cent).free()
    e1.record(ext)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ctx.set_profiling(True)
    acc = {k: [] for k in kn}
    for _ in range(3):
        ds.assign(0, cent).free()
        for k in kn:
            acc[k].append(max(ctx.kernel_ms(k), 0.0))
    ctx.set_profiling(False)
    r = ds.assign(0, cent)
    f = r.fetch()
    r.free()
    same = None
    if ref is None:
        ref = f
    else:
        same = bool(np.array_equal(ref.best, f.best) and np.array_equal(ref.dmin.view(np.uint32), f.dmin.view(np.uint32))
                    and np.array_equal(ref.offsets, f.offsets) and np.array_equal(ref.members, f.members))
    print(json.dumps({"tc_epi_split": split, "csr_sort": csr_sort, "cc_cache": cc_cache, "ms_per_step": ms,
                      "kernels_ms": {k: float(np.mean(v)) for k, v in acc.items()},
                      "overflow_rows": ctx.last_overflow_rows(), "members": int(f.members.size),
                      "identical_to_first_variant": same}), flush=True)
    ds.free()
    ctx.close()
