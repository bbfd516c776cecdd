"""N ranks (torch.distributed.run): the device-resident row-sharded k-means iteration over NCCL
(spf_kmeans) against the host-staged exchange of spfresh_b200.sharded over the SAME shards, which
rank 0 re-creates on its own GPU and drives with in-process ranks (threads).  Also the list-sharded
query against the unsharded one.  Prints SHARDED_NCCL_OK on success (rank 0)."""
import os
import sys
import threading

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spfresh_b200 as spf  # noqa: E402
from spfresh_b200.clustering import ScriptedRandomSource  # noqa: E402
from spfresh_b200.sharded import (DeviceShard, DeviceShardedKMeans, ShardedKMeans, ThreadComm, TorchComm,  # noqa: E402
                                  kmeans_plus_plus, kmeans_plus_plus_device)


def shard_rows(rank, n, d):
    g = np.random.Generator(np.random.Philox(key=7000 + rank))
    cen = 2.0 * np.random.Generator(np.random.Philox(key=6999)).standard_normal((64, d), dtype=np.float32)
    return (cen[g.integers(0, 64, n)] + 0.5 * g.standard_normal((n, d), dtype=np.float32)).astype(np.float32)


def check_kmeans(ctx, comm, rank, world, dev, n=60_000, d=128, k=512, iters=3):
    mine = shard_rows(rank, n, d)
    ds = spf.Dataset(ctx, mine)
    init_rows = (np.arange(k, dtype=np.uint64) * 97) % np.uint64(n)          # rows of rank 0's shard
    init_vec = shard_rows(0, n, d)[init_rows.astype(np.int64)]
    km = DeviceShardedKMeans(ds, comm, spf.METRIC_EUCLIDEAN, rank * n)
    km.init(init_rows, init_vec)
    hist = []
    for _ in range(iters):
        km.step()
        rows, vec, cnt = km.centroids()
        hist.append((rows.copy(), vec.copy(), cnt.copy()))
    # every rank must hold the same centroids
    t = torch.from_numpy(hist[-1][0].view(np.int64).copy()).to(dev)
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert bool((lo == hi).all()), "ranks disagree on the centroid rows"
    if rank == 0:
        shards = [shard_rows(r, n, d) for r in range(world)]
        ctxs = [spf.Context(torch.cuda.current_device()) for _ in range(world)]
        dss = [spf.Dataset(ctxs[r], shards[r]) for r in range(world)]
        grp = ThreadComm.Group(world)
        out = [None] * world

        def run(r):
            ref = ShardedKMeans(DeviceShard(dss[r], r * n, shards[r]), ThreadComm(grp, r), spf.METRIC_EUCLIDEAN)
            ref.rows, ref.vectors = init_rows.copy(), init_vec.copy()
            res = []
            for _ in range(iters):
                ref.step()
                res.append((np.array(ref.rows, copy=True), np.array(ref.vectors, copy=True)))
            out[r] = res
        th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        [x.start() for x in th]
        [x.join() for x in th]
        for it in range(iters):
            assert np.array_equal(hist[it][0], out[0][it][0]), f"iteration {it}: centroid rows differ"
            assert np.array_equal(hist[it][1].view(np.uint32), out[0][it][1].view(np.uint32)), f"iteration {it}: vectors differ"
        for x in dss:
            x.free()
        for x in ctxs:
            x.close()
    km.free()
    ds.free()
    return n * world


def check_kmpp(ctx, comm, rank, world, dev, n=50_000, d=64, k=40):
    """Device-resident sharded k-means++ rounds (spf_kmpp_rounds_sharded, three NCCL all-gathers per round)
    against the host-staged exchange of sharded.kmeans_plus_plus over the same shards and draws, for
    every metric; also a degenerate shard set (identical rows) that must take the uniform fallback."""
    import time
    mine = shard_rows(rank, n, d)
    ds = spf.Dataset(ctx, mine)
    shard = DeviceShard(ds, rank * n, mine)
    host = TorchComm(dev)
    u = np.random.Generator(np.random.Philox(key=31)).random(k).tolist()
    times = []
    for metric in (spf.METRIC_EUCLIDEAN, spf.METRIC_MANHATTAN, spf.METRIC_CHEBYSHEV):
        mk = lambda: ScriptedRandomSource(index=lambda m: (m * 5) // 7, u01=list(u))   # noqa: E731
        t0 = time.perf_counter()
        a = kmeans_plus_plus(shard, host, metric, k, mk())
        t1 = time.perf_counter()
        b = kmeans_plus_plus_device(shard, host, comm, metric, k, mk(), batch=16)
        t2 = time.perf_counter()
        assert np.array_equal(a, b), f"metric {metric}: device-resident picks differ from the host-staged ones"
        assert len(set(int(x) for x in b)) == k
        times.append(((t1 - t0) / (k - 1) * 1e3, (t2 - t1) / (k - 1) * 1e3))
    ds.free()
    same = np.ones((500, 8), np.float32)
    ds2 = spf.Dataset(ctx, same)
    sh2 = DeviceShard(ds2, rank * 500, same)
    a = kmeans_plus_plus(sh2, host, 0, 5, ScriptedRandomSource(index=[3, 700 % (500 * world), 11, 12, 13], u01=list(u)))
    b = kmeans_plus_plus_device(sh2, host, comm, 0, 5, ScriptedRandomSource(index=[3, 700 % (500 * world), 11, 12, 13], u01=list(u)))
    assert np.array_equal(a, b), (a, b)
    ds2.free()
    return "kmpp_ok ms_per_round host-staged/device " + " ".join(f"{x:.2f}/{y:.2f}" for x, y in times)


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = spf.Context(local)
    comm = spf.DeviceComm.from_torch(ctx)
    rows = check_kmeans(ctx, comm, rank, world, dev)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from check_sharded_query import check_query
    extra = check_query(ctx, comm, rank, world, dev)
    extra2 = check_kmpp(ctx, comm, rank, world, dev)
    dist.barrier()
    if rank == 0:
        print(f"SHARDED_NCCL_OK world={world} rows={rows} {extra} {extra2}", flush=True)
    comm.free()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
