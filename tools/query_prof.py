"""Query-path driver for profiling: index over a 1M x 128 assignment, 10k-query batches.
usage: python tools/query_prof.py [nq] [nprobe] [reps] [mode]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import spfresh_b200 as s  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nprobe = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 1
tc = int(sys.argv[5]) if len(sys.argv) > 5 else 1
kind = sys.argv[6] if len(sys.argv) > 6 else "gauss"
rows = bench.make_rows(0)
if kind == "clustered":
    g = np.random.Generator(np.random.Philox(key=44))
    cen = 2.0 * g.standard_normal((1024, bench.DIM), dtype=np.float32)
    rows = (cen[g.integers(0, 1024, rows.shape[0])] + 0.5 * rows).astype(np.float32)
ctx = s.Context(0)
ctx.set_param("scan_list_major", mode)
ctx.set_param("scan_tc", tc)
for name in ("scan_tc_bucket", "scan_tc_tau_probes", "scan_tc_cmax_mb", "scan_tc_split"):
    import os
    if os.environ.get(name.upper()):
        ctx.set_param(name, int(os.environ[name.upper()]))
ds = s.Dataset(ctx, rows)
cent = np.arange(bench.K_CENT, dtype=np.uint64)
res = ds.assign(0, cent)
f = res.fetch(best=False, dmin=False)
med = ds.update_medoids_from(0, res, cent)
res.free()
idx = s.DeviceIndex.pack(ds, f.offsets, f.members, med)
q = bench.make_queries(nq)
if kind == "clustered":
    g = np.random.Generator(np.random.Philox(key=46))
    q = (cen[g.integers(0, 1024, nq)] + 0.5 * q).astype(np.float32)
ctx.set_profiling(True)
for i in range(reps):
    ids, dists, counts = idx.search(q, 10, nprobe)
    b = idx.last_scan_bytes()
    print(f"rep {i}: scan {ctx.kernel_ms('scan'):.3f} ms probe {ctx.kernel_ms('probe'):.3f} ms "
          f"algorithmic {b / 1e9:.1f} GB -> {b / ctx.kernel_ms('scan') / 1e6:.0f} GB/s, mean count {counts.mean():.2f}", flush=True)
    if ctx.kernel_ms("scan_tc_b") > 0:
        print("   tensor scan: " + " ".join(f"{n} {ctx.kernel_ms('scan_tc_' + n):.3f}" for n in
                                            ("a", "tau", "b", "refine", "fallback", "candidates", "flagged", "units", "groups", "split", "hub_units", "hub_tiles", "tiles")), flush=True)
    if ctx.kernel_ms("probe_tc_b") > 0:
        print("   tensor probe: " + " ".join(f"{n} {ctx.kernel_ms('probe_tc_' + n):.3f}" for n in
                                             ("a", "tau", "b", "refine", "candidates", "flagged", "units")), flush=True)
if tc and len(sys.argv) > 7:          # cross-check against the exact scan
    ctx.set_param("scan_tc", 0)
    i2, d2, c2 = idx.search(q, 10, nprobe)
    print("same as exact scan:", bool(np.array_equal(ids, i2) and np.array_equal(dists.view(np.uint32), d2.view(np.uint32))
                                      and np.array_equal(counts, c2)), f"exact scan {ctx.kernel_ms('scan'):.3f} ms")
