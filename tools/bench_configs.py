#!/usr/bin/env python
"""Secondary benchmarks for the BASELINE.json configs that are not the bench.py headline.

    python tools/bench_configs.py gist  [--rows N]          config 3: 1M x 960, Manhattan / Chebyshev (CUDA-core path)
    python tools/bench_configs.py sweep [--nq N] [--kind gauss|clustered]
                                                           config 5: nprobe 8..256 query sweep (QPS vs recall@10)
    python tools/bench_configs.py deep  [--rows-total N]    config 4: 100M x 96 row-sharded assign + k-means iteration
        (multi-GPU: python -m torch.distributed.run --nproc-per-node G tools/bench_configs.py deep ...)

Every run prints one JSON object per measured point on stdout.  Synthetic data only; the shapes
follow SURVEY.md §8(d).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FP32_PEAK = 148 * 128 * 1.965e9      # lane instructions / s at the nominal max clock


def clustered_rows(n, d, ncent, key):
    g = np.random.Generator(np.random.Philox(key=key))
    cen = 2.0 * g.standard_normal((ncent, d), dtype=np.float32)
    out = np.empty((n, d), np.float32)
    step = 1 << 18
    for i in range(0, n, step):
        m = min(step, n - i)
        out[i:i + m] = cen[g.integers(0, ncent, m)] + 0.5 * g.standard_normal((m, d), dtype=np.float32)
    return out


def gist(args):
    import spfresh_b200 as s
    n, d, k = args.rows, 960, 4096
    rows = clustered_rows(n, d, 1024, 45) if args.kind == "clustered" else \
        np.random.Generator(np.random.Philox(key=45)).standard_normal((n, d), dtype=np.float32)
    cent = np.random.Generator(np.random.Philox(key=7)).choice(n, k, replace=False)
    ctx = s.Context(0)
    ctx.set_profiling(True)
    if args.cand_cap:
        ctx.set_param("cand_cap", args.cand_cap)
    ds = s.Dataset(ctx, rows)
    metrics = [(s.METRIC_MANHATTAN, "Manhattan", 2), (s.METRIC_CHEBYSHEV, "Chebyshev", 2), (s.METRIC_EUCLIDEAN, "Euclidean", 3)]
    if args.only:
        metrics = [m for m in metrics if m[1] == args.only]
    for metric, name, instr in metrics:
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            r = ds.assign(metric, cent)
            dt = time.perf_counter() - t0
            km, tc = ctx.kernel_ms("assign_exact"), ctx.kernel_ms("assign_tc")
            other = {x: ctx.kernel_ms(x) for x in ("resolve", "cc_matrix", "csr", "overflow")}
            tot, ovf = r.total, ctx.last_overflow_rows()
            r.free()
            if best is None or dt < best[1]:
                best = (km, dt, other, tot, tc, ovf)
        km, dt, other, tot, tc, ovf = best
        rec = {"config": "gist", "rows": n, "dim": d, "k": k, "metric": name, "data": args.kind,
               "call_ms": dt * 1e3, "points_per_s": n / dt, "other_kernels_ms": other, "members": tot,
               "overflow_rows": ovf}
        if tc > 0:        # squared-Euclidean: tcgen05 TF32 candidate GEMM with streamed point K blocks (d > 128)
            flop = 2.0 * n * k * d
            rec["assign_tc_ms"] = tc
            rec["roofline"] = {"bound": "tensor", "achieved": flop / (tc * 1e-3) / 1e12, "unit": "TFLOP/s",
                               "flop_per_launch_total": flop,
                               "note": "against ~700 TFLOP/s cuBLAS TF32 on this pool (bench.py measures it per run)"}
        else:
            lane = float(n) * k * d * instr
            rec["assign_exact_ms"] = km
            rec["roofline"] = {"bound": "fp32", "achieved": lane / (km * 1e-3) / 1e12, "peak": FP32_PEAK / 1e12,
                               "unit": "T lane-instr/s", "frac": lane / (km * 1e-3) / FP32_PEAK,
                               "lane_instr_per_launch": lane,
                               "peak_source": "148 SM x 128 lanes x 1.965 GHz (nominal max clock)"}
        print(json.dumps(rec), flush=True)


def sweep(args):
    import torch

    import spfresh_b200 as s
    n, d, k, topk = 1_000_000, 128, 4096, 10
    if args.kind == "clustered":
        rows = clustered_rows(n, d, 1024, 44)
        q = clustered_rows(args.nq, d, 1024, 44)           # same centres (Philox key), different draws below
        q = q[::-1].copy()
    else:
        rows = np.random.Generator(np.random.Philox(key=42)).standard_normal((n, d), dtype=np.float32)
        q = np.random.Generator(np.random.Philox(key=46)).standard_normal((args.nq, d), dtype=np.float32)
    ctx = s.Context(0)
    ds = s.Dataset(ctx, rows)
    cent = np.random.Generator(np.random.Philox(key=7)).choice(n, k, replace=False)
    res = ds.assign(0, cent)
    med = ds.update_medoids_from(0, res, cent)             # one k-means step, like fit()
    res.free()
    res = ds.assign(0, med)
    f = res.fetch(best=False, dmin=False)
    res.free()
    idx = s.DeviceIndex.pack(ds, f.offsets, f.members, med)
    # exact ground truth for the first 1000 queries (fp32 brute force on the device)
    dev = torch.device("cuda", 0)
    x = torch.from_numpy(rows).to(dev)
    xq = torch.from_numpy(q[:1000]).to(dev)
    d2 = (xq * xq).sum(1, keepdim=True) - 2.0 * xq @ x.T + (x * x).sum(1)[None, :]
    gt = torch.topk(d2, topk, dim=1, largest=False).indices.cpu().numpy()
    del x, xq, d2
    torch.cuda.empty_cache()
    ctx.set_profiling(True)
    for prune in (1.2, float("inf")):
        for nprobe in (8, 10, 16, 32, 64, 128, 256):
            idx.search(q, topk, nprobe, prune_factor=prune)     # warm-up at full size (scratch allocation)
            t0 = time.perf_counter()
            ids, dists, counts = idx.search(q, topk, nprobe, prune_factor=prune)
            dt = time.perf_counter() - t0
            scan_ms, probe_ms = ctx.kernel_ms("scan"), ctx.kernel_ms("probe")
            b = idx.last_scan_bytes()
            hit = sum(len(set(gt[i].tolist()) & set(ids[i, :counts[i]].tolist())) for i in range(1000))
            lane = b / 4.0 * 3.0
            rec = {"config": "sweep", "data": args.kind, "nq": args.nq, "k": topk, "nprobe": nprobe,
                   "prune_factor": prune if prune < 1e30 else "inf", "qps": args.nq / dt,
                   "recall_at_10": hit / 10000.0, "mean_results": float(counts.mean()),
                   "scan_ms": scan_ms, "probe_ms": probe_ms, "call_ms": dt * 1e3,
                   "scan_algorithmic_gbs": b / (scan_ms * 1e-3) / 1e9, "index_vectors": idx.nvectors}
            if ctx.kernel_ms("scan_tc_a") > 0:          # tensor-core candidate scan (scan_tc.cu)
                rec["scan_path"] = "tcgen05 candidate scan + exact refinement"
                rec["scan_passes_ms"] = {n: ctx.kernel_ms("scan_tc_" + n) for n in ("a", "tau", "b", "refine", "fallback")}
                rec["candidates_per_query"] = ctx.kernel_ms("scan_tc_candidates") / args.nq
                rec["queries_on_exact_fallback"] = int(ctx.kernel_ms("scan_tc_flagged"))
            else:
                rec["scan_path"] = "exact CUDA-core scan"
                rec["scan_fp32_frac"] = lane / (scan_ms * 1e-3) / FP32_PEAK
            rec["probe_path"] = ("tcgen05 candidate scan over the centroid list" if ctx.kernel_ms("probe_tc_a") > 0 else
                                 "tcgen05 dense s matrix + certified selection" if ctx.kernel_ms("probe_tc_select") > 0 else
                                 "exact CUDA-core")
            print(json.dumps(rec), flush=True)


def qshard(args):
    """Config 5 at N GPUs: posting lists sharded by list (balanced by vectors), queries and centroids
    replicated, every rank scans its own lists, partial top-k all-gathered and merged on the stable
    key (spf_topk_merge).  Run under torchrun."""
    import torch
    import torch.distributed as dist

    import spfresh_b200 as s
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n, d, k, topk = 1_000_000, 128, 4096, 10
    rows = clustered_rows(n, d, 1024, 44) if args.kind == "clustered" else \
        np.random.Generator(np.random.Philox(key=42)).standard_normal((n, d), dtype=np.float32)
    q = (clustered_rows(args.nq, d, 1024, 44)[::-1].copy() if args.kind == "clustered" else
         np.random.Generator(np.random.Philox(key=46)).standard_normal((args.nq, d), dtype=np.float32))
    ctx = s.Context(local)
    ds = s.Dataset(ctx, rows)
    cent = np.random.Generator(np.random.Philox(key=7)).choice(n, k, replace=False)
    res = ds.assign(0, cent)
    med = ds.update_medoids_from(0, res, cent)
    res.free()
    res = ds.assign(0, med)
    f = res.fetch(best=False, dmin=False)
    res.free()
    # contiguous list ranges with (nearly) equal numbers of vectors
    cum = f.offsets.astype(np.int64)
    cuts = [int(np.searchsorted(cum, cum[-1] * r // world)) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, k
    idx = s.DeviceIndex.pack(ds, f.offsets, f.members, med, list_range=(cuts[rank], cuts[rank + 1]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx.set_profiling(True)
    if world > 1:                                    # NCCL communicator set-up outside the timed region
        w = torch.zeros(1, device=dev)
        dist.all_gather([torch.empty_like(w) for _ in range(world)], w)
    for it, nprobe in enumerate((8, 8, 32, 128)):        # the first pass warms the merge path up and is not reported
        idx.search(q, topk, nprobe, want_keys=True)
        barrier()
        t0 = time.perf_counter()
        ids, dists, counts, keys = idx.search(q, topk, nprobe, want_keys=True)
        t_scan = time.perf_counter() - t0
        scan_ms = ctx.kernel_ms("scan")
        if world > 1:
            # exchange on the device: all-gather of the (key, id) pairs, k smallest keys per query
            # (keys are globally consistent; distances and counts follow from them)
            sent = np.iinfo(np.int64).max
            kt = torch.from_numpy(keys.view(np.int64)).to(dev)
            kt = torch.where(kt == -1, torch.full_like(kt, sent), kt)          # empty slots sort last
            it = torch.from_numpy(ids.view(np.int64)).to(dev)
            gk = [torch.empty_like(kt) for _ in range(world)]
            gi = [torch.empty_like(it) for _ in range(world)]
            dist.all_gather(gk, kt)
            dist.all_gather(gi, it)
            allk, alli = torch.cat(gk, dim=1), torch.cat(gi, dim=1)
            topv, topi = torch.topk(allk, topk, dim=1, largest=False, sorted=True)
            m_ids = torch.gather(alli, 1, topi)
            m_counts = (topv != sent).sum(dim=1)
            if rank == 0:
                m_ids_h, m_keys_h, m_counts_h = m_ids.cpu().numpy(), topv.cpu().numpy(), m_counts.cpu().numpy()
                if nprobe == 8:                  # cross-check the device merge against spf_topk_merge once
                    gkh = np.stack([g.cpu().numpy() for g in gk]).view(np.uint64)
                    gkh[gkh == np.uint64(sent)] = np.uint64(np.iinfo(np.uint64).max)
                    gih = np.stack([g.cpu().numpy() for g in gi]).view(np.uint64)
                    gdh = (gkh >> np.uint64(32)).astype(np.uint32).view(np.float32)
                    gch = (gkh != np.uint64(np.iinfo(np.uint64).max)).sum(axis=2).astype(np.uint32)
                    r_ids, _, r_counts = s.topk_merge(gkh, gih, gdh, gch)
                    assert np.array_equal(r_counts, m_counts_h.astype(np.uint32))
                    for qi in range(0, args.nq, 997):
                        assert np.array_equal(r_ids[qi, :r_counts[qi]], m_ids_h[qi, :r_counts[qi]].view(np.uint64))
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt, t_scan, scan_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, t_scan, scan_ms = float(t[0]), float(t[1]), float(t[2])
        if rank == 0 and it > 0:
            print(json.dumps({"config": "qshard", "data": args.kind, "n_gpus": world, "nq": args.nq, "k": topk,
                              "nprobe": nprobe, "qps_merged": args.nq / dt, "qps_scan_only": args.nq / t_scan,
                              "scan_kernel_ms_max": scan_ms, "call_ms": t_scan * 1e3, "total_ms": dt * 1e3,
                              "lists_rank0": cuts[1] - cuts[0], "vectors_rank0": idx.nvectors,
                              "merge": "NCCL all-gather of (key, id) + per-query k smallest keys on the device (checked against spf_topk_merge)"}),
                  flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def deep(args):
    import torch
    import torch.distributed as dist

    import spfresh_b200 as s
    from spfresh_b200.sharded import DeviceShard, ShardedKMeans, SingleComm, TorchComm
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    d, k = 96, 4096
    n_total = args.rows_total
    lo, hi = rank * n_total // world, (rank + 1) * n_total // world
    n = hi - lo
    # distribution B generated on the device: 4096 centres ~ 2 N(0, I), x = centre + 0.5 N(0, I)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    centres = 2.0 * torch.randn((4096, d), generator=gen, device=dev)
    gen.manual_seed(99 + rank)
    x = torch.empty((n, d), dtype=torch.float32, device=dev)
    step = 1 << 22
    for i in range(0, n, step):
        m = min(step, n - i)
        lab = torch.randint(0, 4096, (m,), generator=gen, device=dev)
        x[i:i + m] = centres[lab] + 0.5 * torch.randn((m, d), generator=gen, device=dev)
    torch.cuda.synchronize()
    ctx = s.Context(local)
    ds = s.Dataset(ctx, device_ptr=x.data_ptr(), n=n, d=d)
    del x
    torch.cuda.empty_cache()
    comm = TorchComm(dev) if world > 1 else SingleComm()
    km = ShardedKMeans(DeviceShard(ds, lo), comm, s.METRIC_EUCLIDEAN)
    init = np.random.Generator(np.random.Philox(key=7)).choice(n_total, k, replace=False)
    km.init_rows(init)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx.set_profiling(True)
    # flat assign (device resident result, no fetch): the hot path of every k-means iteration
    times, kms = [], []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        r = ds.assign_vectors(0, km.vectors)
        barrier()
        times.append(time.perf_counter() - t0)
        kms.append({x_: ctx.kernel_ms(x_) for x_ in ("assign_tc", "resolve", "cc_matrix", "csr")})
        tot, ovf = r.total, ctx.last_overflow_rows()
        r.free()
    ctx.set_profiling(False)
    it_times = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        km.step()
        barrier()
        it_times.append(time.perf_counter() - t0)
    t_assign, t_iter = min(times[1:]), min(it_times[1:])
    if world > 1:
        t = torch.tensor([t_assign, t_iter], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_assign, t_iter = float(t[0]), float(t[1])
    if rank == 0:
        flops = 2.0 * n_total * k * d
        print(json.dumps({"config": "deep", "rows_total": n_total, "rows_per_gpu": n, "dim": d, "k": k, "n_gpus": world,
                          "scaling": "strong", "assign_ms": t_assign * 1e3, "assign_points_per_s": n_total / t_assign,
                          "assign_tflops_algorithmic": flops / t_assign / 1e12,
                          "kmeans_iteration_ms": t_iter * 1e3, "iteration_points_per_s": n_total / t_iter,
                          "rank0_kernels_ms": kms[-1], "rank0_members": tot, "rank0_overflow_rows": ovf}), flush=True)
    if km.last is not None:
        km.last.free()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["gist", "sweep", "deep", "qshard"])
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--rows-total", type=int, default=100_000_000)
    ap.add_argument("--nq", type=int, default=100_000)
    ap.add_argument("--kind", default="gauss", choices=["gauss", "clustered"])
    ap.add_argument("--cand-cap", type=int, default=0, help="gist: candidate group records per point (default 128)")
    ap.add_argument("--only", default=None, help="gist: run a single metric (Manhattan / Chebyshev / Euclidean)")
    args = ap.parse_args()
    {"gist": gist, "sweep": sweep, "deep": deep, "qshard": qshard}[args.what](args)


if __name__ == "__main__":
    main()
