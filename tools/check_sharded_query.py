"""List-sharded query over the ranks of a group (spf_search_sharded) against the unsharded search of
the same index: every rank checks its own slice of the batch bit for bit.  Used by
tools/check_sharded_nccl.py and bench.py (N > 1)."""
import numpy as np
import torch
import torch.distributed as dist

import spfresh_b200 as spf


def balanced_list_ranges(offsets, world):
    """Contiguous list ranges with about the same number of vectors each (balance by bytes)."""
    sizes = np.diff(offsets.astype(np.int64))
    cum = np.concatenate([[0], np.cumsum(sizes)])
    cuts = [int(np.searchsorted(cum, cum[-1] * r / world)) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, sizes.size
    return [(cuts[r], max(cuts[r + 1], cuts[r])) for r in range(world)]


def check_query(ctx, comm, rank, world, dev, n=200_000, d=128, k_lists=1024, nq=8192, nprobes=(8, 32)):
    g = np.random.Generator(np.random.Philox(key=8100))
    cen = 2.0 * g.standard_normal((256, d), dtype=np.float32)
    data = (cen[g.integers(0, 256, n)] + 0.5 * g.standard_normal((n, d), dtype=np.float32)).astype(np.float32)
    cent = np.random.Generator(np.random.Philox(key=8101)).choice(n, k_lists, replace=False).astype(np.uint64)
    q = (cen[g.integers(0, 256, nq)] + 0.5 * g.standard_normal((nq, d), dtype=np.float32)).astype(np.float32)
    ds = spf.Dataset(ctx, data)
    res = ds.assign(spf.METRIC_EUCLIDEAN, cent)
    f = res.fetch(best=False, dmin=False)
    med = ds.update_medoids_from(spf.METRIC_EUCLIDEAN, res, cent)
    res.free()
    lb, le = balanced_list_ranges(f.offsets, world)[rank]
    mine = spf.DeviceIndex.pack(ds, f.offsets, f.members, med, list_range=(lb, le))
    full = spf.DeviceIndex.pack(ds, f.offsets, f.members, med)
    nql = nq // world
    qs = q[rank * nql:(rank + 1) * nql]
    ok = 1
    for nprobe in nprobes:
        ids, dists, counts = mine.search_sharded(comm, qs, 10, nprobe=nprobe)
        rid, rd, rc = full.search(qs, 10, nprobe=nprobe)
        same = (np.array_equal(counts, rc) and np.array_equal(ids, rid)
                and np.array_equal(dists.view(np.uint32), rd.view(np.uint32)))
        ok = ok and int(same)
    t = torch.tensor([ok], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    mine.free()
    full.free()
    ds.free()
    assert int(t.item()) == 1, "the list-sharded query differs from the unsharded one"
    return f"query_ok nq={nq} nprobes={list(nprobes)}"
