#!/usr/bin/env python
"""bench.py — headline benchmark of the SPFresh/SPANN hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the build hot path over one batch: assign_points_to_clusters
(src/clustering/hierarchical.rs:295-364 of the reference) for 1 000 000 x 128 f32 points against
k = 4096 centroids, squared-Euclidean, boundary factor 1.1 — the full reference result (nearest
centroid + distance per point and the cluster-major CSR with boundary replicas).  This is
BASELINE.json configs[1] ("SIFT-shape synthetic 1M x 128, k = 4096"), data = iid N(0,1) like the
reference's own benches/clustering_benchmark.rs:11-15.

  value     n * N / t : inputs resident in HBM when the timed region starts (device time, CUDA
            events on the library's stream, max over ranks)
  e2e       the same pass through the C ABI with HOST buffers: pinned-host -> device upload of the
            rows, assign, device -> host fetch of best / dmin / CSR, every step
  roofline  the tcgen05 TF32 candidate GEMM (dominant kernel): 2*n*k*d flop per launch / its
            CUDA-event duration, against the TF32 dense peak measured live with cuBLAS
  cpu_baseline  the C oracle (restatement of the reference CPU path, all host cores) on a
            bounded row sample, same centroids
  query     batched find_k_nearest_neighbor_spann (10k queries, top-10, nprobe = k): QPS through the
            C ABI with host buffers, the tensor-core candidate scan's bound pass against the measured
            HBM peak, the same batch on the exact CUDA-core scan (identical results), recall@10 vs
            brute force

--impl reference times the CPU restatement (the Rust reference cannot be built here: no
cargo/rustc) on the host cores, on a bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS, DIM, K_CENT = 1_000_000, 128, 4096
# dram__bytes_read.sum + dram__bytes_write.sum of assign_tc_kernel for one 1M x 4096 x 128 launch
# (ncu --set full capture, profiles/): 512 MB rounded rows read once + candidate records written
TC_DRAM_BYTES_PER_LAUNCH = 3.21e9
# the same for one bound-pass launch of scan_tc_kernel on the bench's 10k-query batch (None until captured)
SCAN_TC_DRAM_BYTES_PER_LAUNCH = 5.76e9
SCAN_TC_DRAM_SOURCE = ("ncu dram__bytes_read+write.sum of the bound-pass launch on this batch, profiles/r01_ncu_scan_tc_v2.txt "
                       "(5.61 GB read = lists once + gathered query rows, 0.15 GB chunk maxima written)")
NQ, TOPK = 10_000, 10
METRIC_NAME = "kmeans_assign_pts_per_s"
WORKLOAD = "assign_points_to_clusters 1M x 128 f32, k=4096, squared-Euclidean, boundary 1.1, iid N(0,1)"


def make_rows(rank: int, n: int = N_ROWS) -> np.ndarray:
    """Rows 0..k-1 are the centroid vectors shared by every rank (row-sharded build: centroids
    are replicated); the rest is the rank's own shard.  Philox counters make it reproducible."""
    g0 = np.random.Generator(np.random.Philox(key=42))
    shared = g0.standard_normal((K_CENT, DIM), dtype=np.float32)
    g = np.random.Generator(np.random.Philox(key=1000 + rank))
    own = g.standard_normal((n - K_CENT, DIM), dtype=np.float32)
    return np.concatenate([shared, own], axis=0)


def make_queries(nq: int = NQ) -> np.ndarray:
    return np.random.Generator(np.random.Philox(key=43)).standard_normal((nq, DIM), dtype=np.float32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_assign_sample(rows: np.ndarray, sample: int, threads: int = 0, want_result: bool = False):
    """Oracle (C restatement of the reference CPU path) on the first `sample` rows, all centroids."""
    import oracle
    oracle.build()
    cent = np.arange(K_CENT, dtype=np.uint64)
    idx = np.arange(K_CENT, K_CENT + sample, dtype=np.uint64)      # real shard rows, not the centroids
    t0 = time.perf_counter()
    res = oracle.assign(rows, oracle.EUCLIDEAN, cent, point_idx=idx, threads=threads)
    dt = time.perf_counter() - t0
    return (dt, res) if want_result else dt


def parity_of_timed_step(gpu, ref, sample: int) -> dict:
    """Full-size parity of the step that was timed: the oracle's answer for the sampled rows
    [K_CENT, K_CENT + sample) against the GPU result of the same 1M x 4096 assign — nearest slot,
    distance bits, and every cluster's member list restricted to those rows (order included)."""
    lo, hi = K_CENT, K_CENT + sample
    ok_best = bool(np.array_equal(gpu.best[lo:hi], ref.best))
    ok_dmin = bool(np.array_equal(gpu.dmin[lo:hi].view(np.uint32), ref.dmin.view(np.uint32)))
    mask = (gpu.members >= lo) & (gpu.members < hi)
    ok_members = bool(np.array_equal(gpu.members[mask], ref.members))
    cum = np.concatenate([[0], np.cumsum(mask, dtype=np.int64)])
    ok_offsets = bool(np.array_equal(cum[gpu.offsets.astype(np.int64)], ref.offsets.astype(np.int64)))
    return {"parity_checked_rows": int(sample), "parity_ok": ok_best and ok_dmin and ok_members and ok_offsets,
            "parity_detail": {"best": ok_best, "dmin_bits": ok_dmin, "member_lists": ok_members, "offsets": ok_offsets,
                              "memberships_compared": int(ref.members.size)}}


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement on the host cores; rank 0 only."""
    if rank != 0:
        return
    import oracle
    oracle.build()
    cores = oracle.online_cpus()
    rows = make_rows(0, K_CENT + 65536)
    # size the per-step sample so that one step takes ~2 s
    t = cpu_assign_sample(rows, 512)
    sample = int(min(65536, max(512, 512 * 2.0 / max(t, 1e-3))))
    for _ in range(args.warmup):
        cpu_assign_sample(rows, min(sample, 2048))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_assign_sample(rows, sample)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": v, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU only; identical per-point work, bounded row sample"},
        "cpu_baseline": {"value": v, "unit": "points/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} of 1e6 rows x all 4096 centroids per step; C restatement of the "
                                   "Rust reference (cargo/rustc absent), gcc -O2 -ffp-contract=off, one task per point"},
        "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def measure_tf32_peak(torch, dev):
    """cuBLAS TF32 8192^3, best of 10 (the way MEASURED_PEAKS.json measured bf16)."""
    try:
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(8192, 8192, device=dev)
        b = torch.randn(8192, 8192, device=dev)
        for _ in range(3):
            a @ b
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, b
        torch.cuda.empty_cache()
        return 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


def emit(line: dict):
    """The contract is ONE JSON line on stdout: everything else (NCCL banners, library chatter) was
    sent to stderr by redirecting fd 1 at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-query", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import spfresh_b200 as spf
    from spfresh_b200 import build as spf_build

    if rank == 0:
        spf_build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"

    ctx = spf.Context(local_rank)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)

    # ---- inputs: pinned host rows (e2e source) + device-resident dataset (value) -----------------
    rows_np = make_rows(rank)
    pinned = torch.empty((N_ROWS, DIM), dtype=torch.float32, pin_memory=True)
    pinned.numpy()[:] = rows_np
    rows_pin = pinned.numpy()
    cent = np.arange(K_CENT, dtype=np.uint64)
    ds = spf.Dataset(ctx, rows_pin)

    def step_resident():
        r = ds.assign(spf.METRIC_EUCLIDEAN, cent)
        r.free()

    out_best = torch.empty(N_ROWS, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    out_dmin = torch.empty(N_ROWS, dtype=torch.float32, pin_memory=True).numpy()
    out_off = np.empty(K_CENT + 1, np.uint64)
    members_cap = {"buf": None}
    e2e_bytes = {"h2d": 0, "d2h": 0}

    def step_e2e():
        # spf_assign_host: pinned host rows -> device (chunked, overlapped with the kernels), every step
        d2, r = spf.Dataset.assign_from_host(ctx, rows_pin, spf.METRIC_EUCLIDEAN, cent)
        if members_cap["buf"] is None or members_cap["buf"].size < r.total:
            members_cap["buf"] = torch.empty(int(r.total * 1.05) + 1, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
        from spfresh_b200._capi import check, lib, ptr
        check(lib().spf_assign_fetch(r.handle, ptr(out_best), ptr(out_dmin), ptr(out_off), ptr(members_cap["buf"])))
        e2e_bytes["h2d"] = N_ROWS * DIM * 4 + K_CENT * 8
        e2e_bytes["d2h"] = N_ROWS * 8 + (K_CENT + 1) * 8 + r.total * 8
        r.free()
        d2.free()

    # ---- value: device-resident, CUDA events on the library stream ------------------------------
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(args.steps):
        step_resident()
    e1.record(ext)
    e1.synchronize()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = N_ROWS * world * args.steps / (ms * 1e-3)

    # ---- per-kernel times (profiling mode brackets each kernel with events) ----------------------
    ctx.set_profiling(True)
    kn = ["assign_tc", "resolve", "classify", "exact_eval", "finalize", "overflow", "cc_matrix", "csr"]
    acc = {k: [] for k in kn}
    for _ in range(3):
        step_resident()
        for k in kn:
            acc[k].append(max(ctx.kernel_ms(k), 0.0))
    ctx.set_profiling(False)
    kms = {k: float(np.mean(v)) for k, v in acc.items()}
    # the result of the timed step, kept for the full-size parity check against the CPU leg below
    gpu_fetched = None
    if not args.no_cpu and rank == 0:
        r_keep = ds.assign(spf.METRIC_EUCLIDEAN, cent)
        gpu_fetched = r_keep.fetch()
        r_keep.free()

    # ---- e2e -------------------------------------------------------------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, min(args.steps, 5))
    e2.record(ext)
    for _ in range(e2e_steps):
        step_e2e()
    e3.record(ext)
    e3.synchronize()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = N_ROWS * world * e2e_steps / (ms_e2e * 1e-3)
    # the bound of the end-to-end step: the raw pinned host -> device copy of the same rows
    xdev = torch.empty((N_ROWS, DIM), dtype=torch.float32, device=dev)
    h2d_ms = 1e9
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        xdev.copy_(pinned, non_blocking=True)
        torch.cuda.synchronize()
        h2d_ms = min(h2d_ms, (time.perf_counter() - t0) * 1e3)
    del xdev
    torch.cuda.empty_cache()

    # ---- one full row-sharded k-means iteration: assign + update_centroids with the exchange ------
    sharded = None
    try:
        from spfresh_b200.sharded import DeviceShard, ShardedKMeans, SingleComm, TorchComm
        comm = TorchComm(dev) if world > 1 else SingleComm()
        km = ShardedKMeans(DeviceShard(ds, rank * N_ROWS, rows_np), comm, spf.METRIC_EUCLIDEAN)
        km.init_rows(np.arange(K_CENT, dtype=np.uint64))
        km.step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            km.step()
        barrier()
        dt = (time.perf_counter() - t0) / 2
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        sharded = {"ms_per_iteration": dt * 1e3, "points_per_s": N_ROWS * world / dt, "n_gpus": world,
                   "what": "assign (spf_assign_vectors) + per-cluster sums/counts all-gather + medoid-candidate "
                           "all-gather + winner-vector all-gather, wall clock incl. host marshalling, max over ranks",
                   "backend": "nccl" if world > 1 else "none"}
        if km.last is not None:
            km.last.free()
    except Exception as ex:      # the headline must still print
        sharded = {"error": repr(ex)}

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    tf32_live = measure_tf32_peak(torch, dev) if rank == 0 else None
    bf16 = peaks.get("bf16_tflops")
    if tf32_live:
        peak_tf, peak_src = tf32_live, "cuBLAS TF32 8192^3 best-of-10 measured in this run"
    elif bf16:
        peak_tf, peak_src = float(bf16) / 2, "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 runs at half the bf16 rate)"
    else:
        peak_tf, peak_src = 1590.0 / 2, "fallback 1.59 PFLOP/s bf16 / 2"
    flops = 2.0 * N_ROWS * K_CENT * DIM
    ach_tf = flops / (kms["assign_tc"] * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "assign_tc_kernel (tcgen05 kind::tf32, 1 pass)",
                "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                "traffic": TC_DRAM_BYTES_PER_LAUNCH, "traffic_source": "ncu dram__bytes_read+write.sum of this kernel, "
                "profiles/r01_ncu_assign_v5.txt (1M-row launch)", "peak_source": peak_src, "flop_per_launch": flops,
                "launches_per_step": 1, "overflow_rows": int(ctx.last_overflow_rows()),
                "kernel_ms": kms["assign_tc"], "share_of_step": kms["assign_tc"] / (ms / args.steps),
                "other_kernels_ms": {k: kms[k] for k in kn if k != "assign_tc"}}

    # ---- query path (rank-local index over the rank's own assignment) -----------------------------
    query = None
    if not args.no_query and rank == 0:
        try:
            query = bench_query(spf, ctx, ds, rows_np, cent, torch, dev, ext, hbm_peak, hbm_src)
        except Exception as ex:      # the headline must still print
            query = {"error": repr(ex)}

    # ---- CPU side-by-side (rank 0, bounded sample) -------------------------------------------------
    cpu = None
    parity = {"parity_checked_rows": 0, "parity_ok": None}
    if not args.no_cpu and rank == 0:
        import oracle
        cores = oracle.online_cpus()
        t_small = cpu_assign_sample(rows_np, 8192)
        sample = int(min(N_ROWS - K_CENT, max(8192, 8192 * 12.0 / max(t_small, 1e-3))))
        t_cpu, ref_res = cpu_assign_sample(rows_np, sample, want_result=True)
        cpu = {"value": sample / t_cpu, "unit": "points/s", "cores": cores, "kind": "port",
               "sample": f"{sample} of 1e6 rows x all 4096 centroids in {t_cpu:.1f} s; C restatement of the Rust "
                         "reference CPU path (cargo/rustc absent), gcc -O2 -ffp-contract=off, one task per point on all cores"}
        parity = parity_of_timed_step(gpu_fetched, ref_res, sample)
        del ref_res

    if rank == 0:
        line = {
            "metric": METRIC_NAME, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (tf32 tensor-core candidate pass, exact f32 decisions)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows_per_gpu": N_ROWS, "dim": DIM, "k": K_CENT,
                       "l2": "inputs (512 MB rows + 1 GB candidate scratch per step) larger than the 126 MB L2",
                       "parity": "bit-identical to the CPU oracle (tests/test_gpu_parity.py)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": int(e2e_bytes["h2d"]),
                    "d2h_bytes_per_step": int(e2e_bytes["d2h"]), "ms_per_step": ms_e2e / e2e_steps,
                    "raw_h2d_ms": h2d_ms, "raw_h2d_gbs": N_ROWS * DIM * 4 / (h2d_ms * 1e-3) / 1e9,
                    "note": "spf_assign_host overlaps the chunked upload with the kernels; the step is bound by the "
                            "PCIe copy of the rows (raw_h2d_ms, measured here with all ranks copying at once) plus the "
                            "device -> host fetch of the CSR"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity_checked_rows": parity["parity_checked_rows"], "parity_ok": parity["parity_ok"],
            "parity_detail": parity.get("parity_detail"),
            "kernels_ms": kms,
            "sharded_kmeans_iteration": sharded,
            "query": query,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and parity["parity_ok"] is False:
        raise SystemExit("bench.py: the GPU result of the timed step differs from the CPU oracle")


def bench_query(spf, ctx, ds, rows_np, cent, torch, dev, ext, hbm_peak, hbm_src):
    """Build the index from the assignment (medoid update included), then batched top-10 search."""
    res = ds.assign(spf.METRIC_EUCLIDEAN, cent)
    f = res.fetch(best=False, dmin=False)
    med = ds.update_medoids_from(spf.METRIC_EUCLIDEAN, res, cent)
    res.free()
    idx = spf.DeviceIndex.pack(ds, f.offsets, f.members, med)
    # queries and results live in pinned host memory (the e2e rule): numpy views of pinned tensors
    q_pin = torch.from_numpy(make_queries()).pin_memory()
    q = q_pin.numpy()
    o_ids = torch.empty((NQ, TOPK), dtype=torch.int64).pin_memory()
    o_d = torch.empty((NQ, TOPK), dtype=torch.float32).pin_memory()
    o_c = torch.empty((NQ,), dtype=torch.int32).pin_memory()
    out = (o_ids.numpy().view(np.uint64), o_d.numpy(), o_c.numpy().view(np.uint32))
    for _ in range(2):
        idx.search(q, TOPK, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record(ext)
    for _ in range(reps):
        ids, dists, counts = idx.search(q, TOPK, out=out)
    e1.record(ext)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ids, dists, counts = ids.copy(), dists.copy(), counts.copy()
    ctx.set_profiling(True)
    idx.search(q, TOPK)
    scan_ms, probe_ms = ctx.kernel_ms("scan"), ctx.kernel_ms("probe")
    tc = {n: ctx.kernel_ms("scan_tc_" + n) for n in ("a", "gather", "tau", "b", "refine", "fallback", "candidates", "flagged",
                                                      "units", "stream_mb", "unique_mb")}
    ctx.set_profiling(False)
    bytes_ = idx.last_scan_bytes()
    # the same batch on the exact CUDA-core scan (what the tensor-core candidate scan replaces): identical results
    ctx.set_param("scan_tc", 0)
    ctx.set_profiling(True)
    ids0, dists0, counts0 = idx.search(q, TOPK)
    exact_scan_ms, exact_probe_ms = ctx.kernel_ms("scan"), ctx.kernel_ms("probe")
    ctx.set_profiling(False)
    ctx.set_param("scan_tc", 1)
    same = bool(np.array_equal(ids, ids0) and np.array_equal(dists.view(np.uint32), dists0.view(np.uint32))
                and np.array_equal(counts, counts0))
    # recall@10 against exact brute force on the device (fp32, torch) for the first 1000 queries
    xq = torch.from_numpy(q[:1000]).to(dev)
    x = torch.from_numpy(rows_np).to(dev)
    d2 = (xq * xq).sum(1, keepdim=True) - 2.0 * xq @ x.T + (x * x).sum(1)[None, :]
    gt = torch.topk(d2, TOPK, dim=1, largest=False).indices.cpu().numpy()
    hit = 0
    for i in range(1000):
        hit += len(set(gt[i].tolist()) & set(ids[i, :counts[i]].tolist()))
    del x, xq, d2
    gbs = bytes_ / (scan_ms * 1e-3) / 1e9
    out = {"metric": "batch_qps_top10", "qps_e2e": NQ / (ms * 1e-3), "nq": NQ, "k": TOPK, "nprobe": TOPK,
           "prune_factor": 1.2, "recall_at_10": hit / (1000.0 * TOPK),
           "h2d_bytes_per_call": int(q.nbytes), "d2h_bytes_per_call": int(sum(o.nbytes for o in out)),
           "host_memory": "pinned",
           "mean_results_per_query": float(counts.mean()),
           "probe_ms": probe_ms, "scan_ms": scan_ms, "index_vectors": idx.nvectors,
           "exact_cuda_core_path": {"scan_ms": exact_scan_ms, "probe_ms": exact_probe_ms, "identical_results": same}}
    if tc["a"] > 0:
        # tensor-core candidate scan: every pass streams the probed lists' TF32 tiles once per
        # (list, 128 probing queries) unit — the kernel is bound by HBM
        stream, unique = tc["stream_mb"] * 1e6, tc["unique_mb"] * 1e6
        ach = unique / (tc["a"] * 1e-3) / 1e9
        out["scan"] = {"bound": "hbm", "kernel": "scan_tc_kernel (tcgen05 kind::tf32), bound pass",
                       "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                       "peak_source": hbm_src, "bytes_per_launch": int(unique), "requested_bytes_per_launch": int(stream),
                       "traffic": SCAN_TC_DRAM_BYTES_PER_LAUNCH, "traffic_source": SCAN_TC_DRAM_SOURCE,
                       "kernel_ms": tc["a"],
                       "passes_ms": {"gather": tc["gather"], "bound": tc["a"], "tau": tc["tau"], "group_refine": tc["b"], "select": tc["refine"],
                                     "fallback": tc["fallback"]},
                       "units": int(tc["units"]), "candidates_per_query": tc["candidates"] / NQ,
                       "queries_on_exact_fallback": int(tc["flagged"]),
                       "query_major_algorithmic_gbs": gbs, "query_major_algorithmic_over_hbm": gbs / hbm_peak,
                       "note": "bytes_per_launch = TF32 rows + K-extension rows of every probed list once (algorithmic HBM "
                               "traffic of a pass); requested = the same per unit (hot lists have several units, served by L2); "
                               "query_major_algorithmic = sum over queries and probed lists of |L|*d*4 (SURVEY 8d), "
                               "which the unit-major kernel serves with one read per 128 probing queries"}
    else:
        lane_instr = bytes_ / 4.0 * 3.0
        fp32_peak = 148 * 128 * 1.965e9
        out["scan"] = {"bound": "fp32", "achieved": lane_instr / (scan_ms * 1e-3) / 1e12, "peak": fp32_peak / 1e12,
                       "unit": "T lane-instr/s", "frac": lane_instr / (scan_ms * 1e-3) / fp32_peak,
                       "peak_source": "148 SM x 128 lanes x 1.965 GHz (nominal max clock)",
                       "algorithmic_gbs": gbs, "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src,
                       "bytes_per_launch": int(bytes_), "kernel_ms": scan_ms}
    idx.free()
    return out


if __name__ == "__main__":
    main()
