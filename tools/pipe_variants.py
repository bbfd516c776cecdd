"""Kernel variants selected by one context parameter (default tc_pipe: assign_tc_kernel epilogue 0 / 1) on
the device-resident assign step (1M x 128, k = 4096, N(0,1)): per-kernel times and result identity.
usage: python tools/pipe_variants.py [steps] [values, comma separated] [parameter name]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
OUT = os.dup(1)
import bench  # noqa: E402  (redirects fd 1 to stderr)
import spfresh_b200 as s  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1]
pname = sys.argv[3] if len(sys.argv) > 3 else "tc_pipe"
rows = bench.make_rows(0)
cent = np.arange(bench.K_CENT, dtype=np.uint64)
kn = ["assign_tc", "classify", "exact_eval", "finalize", "overflow", "cc_matrix", "csr"]
ref = None
for pipe in variants:
    ctx = s.Context(0)
    ctx.set_param(pname, pipe)
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
    for _ in range(5):
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
    os.write(OUT, (json.dumps({pname: pipe, "ms_per_step": ms,
                               "kernels_ms": {k: round(float(np.median(v)), 4) for k, v in acc.items()},
                               "overflow_rows": ctx.last_overflow_rows(), "members": int(f.members.size),
                               "identical_to_first_variant": same}) + "\n").encode())
    ds.free()
    ctx.close()
