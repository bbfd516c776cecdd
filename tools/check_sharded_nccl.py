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
from spfresh_b200.sharded import DeviceShard, DeviceShardedKMeans, ShardedKMeans, ThreadComm  # noqa: E402


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
    dist.barrier()
    if rank == 0:
        print(f"SHARDED_NCCL_OK world={world} rows={rows} {extra}", flush=True)
    comm.free()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
