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
