"""GIST-shape query check (d = 960): tensor-core scan with streamed query K blocks vs the exact scan.
usage: python tools/gist_query_prof.py [n] [nlists] [nq] [nprobe]"""
import sys

import numpy as np

sys.path.insert(0, ".")
import spfresh_b200 as s  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
nlists = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
nprobe = int(sys.argv[4]) if len(sys.argv) > 4 else 10
d = 960
g = np.random.Generator(np.random.Philox(key=45))
cen = 2.0 * g.standard_normal((256, d), dtype=np.float32)
rows = (cen[g.integers(0, 256, n)] + 0.5 * g.standard_normal((n, d), dtype=np.float32)).astype(np.float32)
q = (cen[g.integers(0, 256, nq)] + 0.5 * g.standard_normal((nq, d), dtype=np.float32)).astype(np.float32)
ctx = s.Context(0)
ds = s.Dataset(ctx, rows)
cent = np.random.Generator(np.random.Philox(key=7)).choice(n, nlists, replace=False)
res = ds.assign(0, cent)
f = res.fetch(best=False, dmin=False)
res.free()
idx = s.DeviceIndex.pack(ds, f.offsets, f.members, cent)
ctx.set_profiling(True)
out = {}
import os  # noqa: E402
if os.environ.get("SCAN_TC_BUCKET"):
    ctx.set_param("scan_tc_bucket", int(os.environ["SCAN_TC_BUCKET"]))
for tc in (1, 0):
    ctx.set_param("scan_tc", tc)
    idx.search(q, 10, nprobe)
    out[tc] = idx.search(q, 10, nprobe)
    print(f"scan_tc={tc}: scan {ctx.kernel_ms('scan'):.3f} ms probe {ctx.kernel_ms('probe'):.3f} ms "
          f"(bound pass {ctx.kernel_ms('scan_tc_a'):.3f}, group refinement {ctx.kernel_ms('scan_tc_b'):.3f}, "
          f"refine {ctx.kernel_ms('scan_tc_refine'):.3f}, fallback {ctx.kernel_ms('scan_tc_fallback'):.3f}, "
          f"candidates/query {ctx.kernel_ms('scan_tc_candidates') / nq:.1f}, flagged {ctx.kernel_ms('scan_tc_flagged'):.0f})", flush=True)
same = all(np.array_equal(a.view(np.uint8), b.view(np.uint8)) for a, b in zip(out[1], out[0]))
print("identical results:", same, "index vectors", idx.nvectors)
