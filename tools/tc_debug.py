"""Diagnostic run of the tcgen05 assign path against the exact CUDA-core path (same library).
Prints timing and mismatch statistics instead of asserting, so one GPU call tells the story."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import spfresh_b200 as s  # noqa: E402


def run(n, d, k, kind="gauss", seed=1):
    g = np.random.Generator(np.random.Philox(key=seed))
    if kind == "gauss":
        data = g.standard_normal((n, d), dtype=np.float32)
    else:
        cen = 2.0 * g.standard_normal((max(k // 4, 2), d), dtype=np.float32)
        data = (cen[g.integers(0, cen.shape[0], n)] + 0.5 * g.standard_normal((n, d), dtype=np.float32)).astype(np.float32)
    cent = g.choice(n, k, replace=False)
    ctx = s.Context.default()
    ctx.set_profiling(True)
    ds = s.Dataset(ctx, data)
    t0 = time.time()
    r = ds.assign(0, cent)
    t1 = time.time()
    a = r.fetch()
    ms_tc = ctx.kernel_ms("assign_tc")
    ms_res, ms_csr, ms_cc = ctx.kernel_ms("resolve"), ctx.kernel_ms("csr"), ctx.kernel_ms("cc_matrix")
    t2 = time.time()
    rb = ds.assign(0, cent, flags=s.ASSIGN_FORCE_EXACT)
    t3 = time.time()
    b = rb.fetch()
    ms_ex = ctx.kernel_ms("assign_exact")
    ok_best = np.array_equal(a.best, b.best)
    ok_dmin = np.array_equal(a.dmin.view(np.uint32), b.dmin.view(np.uint32))
    ok_csr = np.array_equal(a.offsets, b.offsets) and np.array_equal(a.members, b.members)
    nb = int((a.best != b.best).sum())
    print(f"n={n} d={d} k={k} {kind}: tc {ms_tc:.3f} ms ({n / ms_tc / 1e3:.1f} Mpts/s, "
          f"{2.0 * n * k * d / ms_tc / 1e9:.1f} TFLOP/s) resolve {ms_res:.3f} cc {ms_cc:.3f} csr {ms_csr:.3f} | "
          f"exact {ms_ex:.3f} ms | call tc {1e3 * (t1 - t0):.1f} ms exact {1e3 * (t3 - t2):.1f} ms | "
          f"members {a.members.size} ({a.members.size / n:.2f}/pt) | best_ok={ok_best} ({nb} differ) "
          f"dmin_ok={ok_dmin} csr_ok={ok_csr}", flush=True)
    r.free(); rb.free(); ds.free()
    return ok_best and ok_dmin and ok_csr


if __name__ == "__main__":
    ok = True
    ok &= run(4096, 128, 256)
    ok &= run(4096, 128, 256, "clustered")
    ok &= run(20000, 96, 1000, "clustered")
    ok &= run(100000, 128, 4096, "clustered")
    ok &= run(100000, 128, 4096, "gauss")
    if "--big" in sys.argv:
        ok &= run(1000000, 128, 4096, "clustered")
        ok &= run(1000000, 128, 4096, "gauss")
    print("ALL OK" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)
