"""Where the end-to-end step goes: python tools/e2e_prof.py
pinned host buffers (the bench's e2e) against ordinary heap memory (what the reference's ndarray callers
hold), the latter through the library's staging ring and through the driver's own pageable path."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import spfresh_b200 as s  # noqa: E402
from spfresh_b200._capi import check, lib, ptr  # noqa: E402

rows_np = bench.make_rows(0)
pinned = torch.empty((bench.N_ROWS, bench.DIM), dtype=torch.float32, pin_memory=True)
pinned.numpy()[:] = rows_np
ctx = s.Context(0)
cent = np.arange(bench.K_CENT, dtype=np.uint64)
dev = torch.device("cuda", 0)
x = torch.empty((bench.N_ROWS, bench.DIM), device=dev)
for _ in range(3):
    t0 = time.perf_counter(); x.copy_(pinned, non_blocking=True); torch.cuda.synchronize(); t_h2d = time.perf_counter() - t0
print(f"raw H2D 512 MB pinned: {t_h2d * 1e3:.2f} ms = {512.0 / t_h2d / 1e3:.1f} GB/s")
for kind in ("pinned", "pageable + staging ring", "pageable, driver-staged"):
    pin = kind == "pinned"
    ctx.set_param("no_host_staging", 1 if kind.endswith("driver-staged") else 0)
    rows = pinned.numpy() if pin else rows_np
    mk = (lambda n, dt: torch.empty(n, dtype=dt, pin_memory=True).numpy()) if pin else (lambda n, dt: torch.empty(n, dtype=dt).numpy())
    ob = mk(bench.N_ROWS, torch.int32).view(np.uint32)
    od = mk(bench.N_ROWS, torch.float32)
    oo = np.empty(bench.K_CENT + 1, np.uint64)
    om = mk(12_000_000, torch.int64).view(np.uint64)
    om[:] = 0                                           # touch the pages once
    for i in range(4):
        t0 = time.perf_counter()
        ds, r = s.Dataset.assign_from_host(ctx, rows, 0, cent)
        t1 = time.perf_counter()
        check(lib().spf_assign_fetch(r.handle, ptr(ob), ptr(od), ptr(oo), ptr(om)))
        t2 = time.perf_counter()
        r.free(); ds.free()
        t3 = time.perf_counter()
        if i:
            print(f"{kind}: step {i}: assign_host {1e3 * (t1 - t0):.2f} ms, fetch {1e3 * (t2 - t1):.2f} ms, "
                  f"free {1e3 * (t3 - t2):.2f} ms, total {1e3 * (t3 - t0):.2f}", flush=True)
