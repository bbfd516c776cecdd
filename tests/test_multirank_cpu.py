"""CPU test of the N > 1 query path: list-sharded partial top-k exchanged with an all_gather over
gloo (world size 2) and merged on the stable key with spf_topk_merge — the exchange §8(e)
describes, minus the GPU scan that produces the partials."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

NQ, K = 40, 6


def make_partials(world):
    rng = np.random.default_rng(5)
    pool = []
    for q in range(NQ):
        n = int(rng.integers(0, 3 * K))
        d = rng.random(n).astype(np.float32)
        seq = rng.choice(10_000, n, replace=False)
        keys = [(int(x.view(np.uint32)) << 32) | int(s) for x, s in zip(d, seq)]
        owner = rng.integers(0, world, n)
        pool.append((keys, owner))
    parts = []
    for r in range(world):
        keys = np.full((NQ, K), np.iinfo(np.uint64).max, np.uint64)
        counts = np.zeros(NQ, np.uint32)
        for q, (kk, owner) in enumerate(pool):
            mine = sorted(k for k, o in zip(kk, owner) if o == r)[:K]
            counts[q] = len(mine)
            keys[q, :len(mine)] = mine
        parts.append((keys, counts))
    expect = [sorted(kk)[:K] for kk, _ in pool]
    return parts, expect


def worker(rank, world, port, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import spfresh_b200 as s
    parts, expect = make_partials(world)
    keys, counts = parts[rank]
    ids = (keys & np.uint64(0xffff)).astype(np.uint64)
    dists = (keys >> np.uint64(32)).astype(np.uint32).view(np.float32)
    gk = [torch.zeros((NQ, K), dtype=torch.int64) for _ in range(world)]
    gi = [torch.zeros((NQ, K), dtype=torch.int64) for _ in range(world)]
    gd = [torch.zeros((NQ, K), dtype=torch.float32) for _ in range(world)]
    gc = [torch.zeros(NQ, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gk, torch.from_numpy(keys.view(np.int64)))
    dist.all_gather(gi, torch.from_numpy(ids.view(np.int64)))
    dist.all_gather(gd, torch.from_numpy(dists.copy()))
    dist.all_gather(gc, torch.from_numpy(counts.view(np.int32)))
    o_ids, o_d, o_c = s.topk_merge(np.stack([t.numpy().view(np.uint64) for t in gk]),
                                   np.stack([t.numpy().view(np.uint64) for t in gi]),
                                   np.stack([t.numpy() for t in gd]),
                                   np.stack([t.numpy().view(np.uint32) for t in gc]))
    good = True
    for q in range(NQ):
        good &= int(o_c[q]) == len(expect[q])
        good &= o_ids[q, :o_c[q]].tolist() == [e & 0xffff for e in expect[q]]
    ok[rank] = 1 if good else 0
    dist.barrier()
    dist.destroy_process_group()


def test_list_sharded_merge_over_gloo():
    from spfresh_b200 import build as b
    b.build()
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ok = mp.get_context("spawn").Array("i", [0, 0])
    mp.spawn(worker, args=(2, port, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]


# ----------------------------------------------------------------------------------------------
# row-sharded build: partial sums / counts and medoid candidates exchanged over gloo
# ----------------------------------------------------------------------------------------------
def sharded_inputs():
    g = np.random.Generator(np.random.Philox(key=31))
    cen = 3.0 * g.standard_normal((12, 10), dtype=np.float32)
    data = (cen[g.integers(0, 12, 900)] + 0.4 * g.standard_normal((900, 10), dtype=np.float32)).astype(np.float32)
    init = np.random.default_rng(2).choice(900, 12, replace=False)
    return data, init


def run_sharded(comm, data, init, bounds, metric, iters=3):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from shard_ref import OracleShard
    from spfresh_b200.sharded import ShardedKMeans
    lo, hi = bounds[comm.rank], bounds[comm.rank + 1]
    km = ShardedKMeans(OracleShard(data[lo:hi], lo), comm, metric)
    km.init_rows(init)
    hist = []
    for _ in range(iters):
        hist.append(np.array(km.step(), copy=True))
    return hist, km.vectors


def sharded_worker(rank, world, port, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spfresh_b200.sharded import TorchComm
    data, init = sharded_inputs()
    hist, vecs = run_sharded(TorchComm(), data, init, [0, 400, 900], 0)
    out = np.concatenate([h.astype(np.float64) for h in hist] + [vecs.astype(np.float64).ravel()])
    t = torch.from_numpy(out)
    got = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(got, t)
    same = all(torch.equal(got[0], g) for g in got)              # every rank ends with the same centroids
    ok[rank] = 1 if same else 0
    if rank == 0:
        np.save(os.path.join(os.environ["SPF_TEST_TMP"], "gloo_rows.npy"), np.stack(hist))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_update_over_gloo(tmp_path):
    """World size 2 over gloo == the same two shards exchanged in-process (ThreadComm), and the
    sharded medoids agree with the single-process oracle on this well-separated data."""
    import sys
    import threading
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle
    from spfresh_b200.sharded import SingleComm, ThreadComm
    oracle.build()
    os.environ["SPF_TEST_TMP"] = str(tmp_path)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ok = mp.get_context("spawn").Array("i", [0, 0])
    mp.spawn(sharded_worker, args=(2, port, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]
    gloo_rows = np.load(tmp_path / "gloo_rows.npy")

    data, init = sharded_inputs()
    grp = ThreadComm.Group(2)
    res = [None, None]

    def run(r):
        res[r] = run_sharded(ThreadComm(grp, r), data, init, [0, 400, 900], 0)
    th = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert np.array_equal(np.stack(res[0][0]), gloo_rows) and np.array_equal(np.stack(res[1][0]), gloo_rows)
    # one shard holding everything == the reference's own update_centroids, bit for bit
    single, _ = run_sharded(SingleComm(), data, init, [0, 900], 0, iters=1)
    a = oracle.assign(data, 0, init)
    assert np.array_equal(single[0], oracle.update_medoids(data, 0, a.offsets, a.members, init.astype(np.uint64)))
    # two shards: same medoids (the mean differs only in f32 summation order; no near-tie here)
    assert np.array_equal(gloo_rows[0], single[0])


def run_sharded_kmpp(comm, data, bounds, metric, k, first, u01):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from shard_ref import OracleShard
    from spfresh_b200.clustering import ScriptedRandomSource
    from spfresh_b200.sharded import kmeans_plus_plus
    lo, hi = bounds[comm.rank], bounds[comm.rank + 1]
    return kmeans_plus_plus(OracleShard(data[lo:hi], lo), comm, metric, k,
                            ScriptedRandomSource(index=[first], u01=u01))


KMPP_U = [0.11, 0.93, 0.5, 0.27, 0.68, 0.04, 0.81, 0.39, 0.99]


def kmpp_worker(rank, world, port, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spfresh_b200.sharded import TorchComm
    data, _ = sharded_inputs()
    rows = run_sharded_kmpp(TorchComm(), data, [0, 400, 900], 0, 10, 321, KMPP_U)
    if rank == 0:
        np.save(os.path.join(os.environ["SPF_TEST_TMP"], "kmpp_rows.npy"), rows)
    ok[rank] = 1
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_kmeanspp_over_gloo(tmp_path):
    """Sharded k-means++: world size 2 over gloo == two in-process shards; one shard == the oracle's
    k-means++ for the same draws; two shards pick the same rows on this data (only the f32 sum
    order differs)."""
    import sys
    import threading
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle
    from spfresh_b200.sharded import SingleComm, ThreadComm
    oracle.build()
    os.environ["SPF_TEST_TMP"] = str(tmp_path)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ok = mp.get_context("spawn").Array("i", [0, 0])
    mp.spawn(kmpp_worker, args=(2, port, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]
    gloo_rows = np.load(tmp_path / "kmpp_rows.npy")
    data, _ = sharded_inputs()
    grp = ThreadComm.Group(2)
    res = [None, None]

    def run(r):
        res[r] = run_sharded_kmpp(ThreadComm(grp, r), data, [0, 400, 900], 0, 10, 321, KMPP_U)
    th = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert np.array_equal(res[0], gloo_rows) and np.array_equal(res[1], gloo_rows)
    single = run_sharded_kmpp(SingleComm(), data, [0, 900], 0, 10, 321, KMPP_U)
    ref, _ = oracle.kmeanspp(data, 0, 10, 321, KMPP_U)
    assert np.array_equal(single, ref)
    assert np.array_equal(gloo_rows, ref)


# ----------------------------------------------------------------------------------------------
# row-sharded fit: assign + update + bisect work-list (hierarchical.rs:65-135)
# ----------------------------------------------------------------------------------------------
def run_sharded_fit(comm, data, init, bounds, metric, desired, make_shard=None):
    """initialize (given rows) -> assign_points -> update_centroids -> subdivide_clusters on row
    shards; returns the clusters with GLOBAL member lists (slices concatenated in rank order)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from shard_ref import OracleShard
    from spfresh_b200.clustering import ScriptedRandomSource
    from spfresh_b200.sharded import ShardedKMeans, clusters_from_assignment, subdivide_clusters
    lo, hi = bounds[comm.rank], bounds[comm.rank + 1]
    shard = make_shard(comm.rank) if make_shard else OracleShard(data[lo:hi], lo)
    km = ShardedKMeans(shard, comm, metric)
    km.init_rows(init)
    km.step()                                            # assign (old centroids) + update (new rows)
    clusters = clusters_from_assignment(shard, comm, km.last, km.rows)
    rng = ScriptedRandomSource(index=lambda m: (m * 5) // 7)
    clusters = subdivide_clusters(shard, comm, metric, clusters, desired, rng, km.starts)
    out = []
    for c in clusters:
        width = int(c.counts.max()) if c.counts.size else 0
        pad = np.full(max(width, 1), -1, np.int64)
        pad[:c.local_points.size] = c.local_points.astype(np.int64) + int(km.starts[comm.rank])
        parts = comm.allgather(pad)
        pts = np.concatenate([p[p >= 0] for p in parts])
        out.append((c.centroid_idx, pts.tolist(), c.depth))
    return out


def fit_worker(rank, world, port, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spfresh_b200.sharded import TorchComm
    import pickle
    data, init = sharded_inputs()
    out = run_sharded_fit(TorchComm(), data, init[:4], [0, 400, 900], 0, 60)
    with open(os.path.join(os.environ["SPF_TEST_TMP"], f"fit_{rank}.pkl"), "wb") as f:
        pickle.dump(out, f)
    ok[rank] = 1
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_bisect_over_gloo(tmp_path):
    """Row-sharded fit (assign, update, bisect work-list): world size 2 over gloo == two in-process
    shards == one shard == the single-process oracle fit for the same random decisions."""
    import pickle
    import sys
    import threading
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle
    from spfresh_b200.sharded import SingleComm, ThreadComm
    oracle.build()
    os.environ["SPF_TEST_TMP"] = str(tmp_path)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ok = mp.get_context("spawn").Array("i", [0, 0])
    mp.spawn(fit_worker, args=(2, port, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]
    gloo = [pickle.load(open(tmp_path / f"fit_{r}.pkl", "rb")) for r in range(2)]
    assert gloo[0] == gloo[1]                            # every rank ends with the same global clusters

    data, init = sharded_inputs()
    grp = ThreadComm.Group(2)
    res = [None, None]

    def run(r):
        res[r] = run_sharded_fit(ThreadComm(grp, r), data, init[:4], [0, 400, 900], 0, 60)
    th = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert res[0] == gloo[0] and res[1] == gloo[0]
    single = run_sharded_fit(SingleComm(), data, init[:4], [0, 900], 0, 60)
    ref = oracle.fit(data, 0, init[:4], 60, pick=lambda m: (m * 5) // 7)
    ref_t = [(int(c.centroid_idx), np.asarray(c.points).tolist(), int(c.depth)) for c in ref]
    assert single == ref_t                               # one shard: the reference's fit, bit for bit
    assert len(ref_t) > 4                                # the bisect did fire
    assert gloo[0] == ref_t                              # two shards: same clusters on this data
