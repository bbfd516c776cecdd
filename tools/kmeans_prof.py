"""Per-part timings of spf_kmeans iterations (one GPU): 1M x 128 N(0,1) and a DEEP-shape shard."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
OUT = os.dup(1)
import bench  # noqa: E402
import spfresh_b200 as s  # noqa: E402

names = ["assign_tc", "classify", "exact_eval", "finalize", "resolve", "cc_matrix", "csr", "overflow", "kmeans_seed", "kmeans_sums",
         "kmeans_exchange", "kmeans_means", "kmeans_medoid", "csr_scan", "csr_fill", "csr_sort"]
dev = torch.device("cuda", 0)
ctx = s.Context(0)
ext = torch.cuda.ExternalStream(ctx.stream)


def run(tag, ds, init_rows, init_vec, steps=5, sum_hub=0, sum_slices=0):
    ctx.set_param("sum_hub", sum_hub)
    ctx.set_param("sum_slices", sum_slices)
    tag = f"{tag} hub={sum_hub} slices={sum_slices}"
    sess = s.KMeansSession(ds, None, 0, 0, init_rows.size)
    sess.set_centroids(init_rows, init_vec)
    for it in range(steps):
        ctx.set_profiling(it >= 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        sess.step()
        e1.record(ext)
        e1.synchronize()
        rec = {"case": tag, "iteration": it, "ms": e0.elapsed_time(e1), "members": sess.assignment().total,
               "overflow_rows": ctx.last_overflow_rows()}
        if it >= 1:
            rec["parts_ms"] = {n: round(ctx.kernel_ms(n), 3) for n in names if ctx.kernel_ms(n) >= 0}
        os.write(OUT, (json.dumps(rec) + "\n").encode())
    ctx.set_profiling(False)
    each = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        sess.step()
        e1.record(ext)
        e1.synchronize()
        each.append(round(e0.elapsed_time(e1), 2))
    os.write(OUT, (json.dumps({"case": tag, "steady_ms_each": each}) + "\n").encode())
    sess.free()


rows = bench.make_rows(0)
ds = s.Dataset(ctx, rows)
ref = None
for hub, sl in ((0, 1),):
    run("1M x 128 gauss", ds, np.arange(4096, dtype=np.uint64), rows[:4096], steps=4, sum_hub=hub, sum_slices=sl)
ds.free()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 25_000_000
x = bench.device_clustered(torch, dev, n, 96, 4096, 1234, 99)
ds = s.Dataset(ctx, device_ptr=x.data_ptr(), n=n, d=96)
del x
torch.cuda.empty_cache()
init = np.sort(np.random.Generator(np.random.Philox(key=7)).choice(n, 4096, replace=False)).astype(np.uint64)
for cache in (1, 0):
    ctx.set_param("scratch_cache", cache)
    ctx.trim()
    hub, sl = 0, cache
    run(f"{n} x 96 clustered cache={cache}", ds, init, ds.fetch_rows(init), steps=3, sum_hub=hub, sum_slices=sl)
