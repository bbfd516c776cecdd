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
# (ncu --set full capture, profiles/r02_ncu_assign_tc_v1.txt): 0.52 GB rounded rows read once + 1.35 GB of
# candidate records written as whole sectors (no read-fill)
TC_DRAM_BYTES_PER_LAUNCH = 1.88e9
# the same for one bound-pass launch of scan_tc_kernel on the bench's 10k-query batch (None until captured)
SCAN_TC_DRAM_BYTES_PER_LAUNCH = 5.76e9
SCAN_TC_DRAM_SOURCE = ("ncu dram__bytes_read+write.sum of the bound pass's two launches on this batch, "
                       "profiles/r02_ncu_scan_tc_split_v1.txt (multi-unit lists on 74 SMs: 1.19 GB read + 0.10 GB written; "
                       "single-unit lists on 74 SMs: 4.40 GB read + 0.08 GB written; under ncu the launches run one after the other)")
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
    """SM clock and throttle reasons DURING the timed region.  NVML is polled from a thread every
    millisecond (the timed region of the default run is a few tens of milliseconds, so `nvidia-smi
    -lms 100` saw 0-2 samples of it: VERDICT r1 weak 11); the nvidia-smi process is the fallback when
    the NVML binding cannot be loaded."""

    _REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"),
                (0x4, "sw_power_cap"))

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread = index, [], None, None
        self.sm, self.mask, self.max_mhz, self.h, self.run = [], 0, None, None, False
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
            self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:
            self.h = None

    def _poll(self):
        nv = self.nv
        while self.run:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        if self.h is not None:
            self.run = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
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
        if self.h is not None:
            self.run = False
            self.thread.join(timeout=1)
            reasons = sorted(nm for bit, nm in self._REASONS if self.mask & bit)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(self.sm), "source": "nvml, 1 ms poll"}
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
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


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
    """cuBLAS TF32 8192^3 the way MEASURED_PEAKS.json measured bf16: best of 10 (burst) and back to back
    for 4 s (sustained, under the power cap).  Returns (burst, sustained) in TFLOP/s."""
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
        flop = 2.0 * 8192 ** 3
        n_loop = max(8, int(4.0 / (best * 1e-3)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_loop):
            a @ b
        e1.record()
        e1.synchronize()
        sustained = flop * n_loop / (e0.elapsed_time(e1) * 1e-3) / 1e12
        del a, b
        torch.cuda.empty_cache()
        return flop / (best * 1e-3) / 1e12, sustained
    except Exception:
        return None, None


def bind_to_gpu_numa_node(torch, local_rank: int) -> dict:
    """One process per GPU: run this rank's host threads (and, by first touch, its pinned buffers) on the
    NUMA node the GPU hangs off, so that N ranks do not all stage through one socket's memory (round 1:
    the per-GPU upload rate fell from 55 to 23 GB/s at N = 8)."""
    info = {"numa_node": None, "bound": False}
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
            info["cpus"] = len(allowed)
    except Exception as ex:
        info["error"] = repr(ex)
    return info


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
    ap.add_argument("--no-configs", action="store_true", help="skip the compact blocks of BASELINE configs 3 / 4 / 5")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    t_main = time.perf_counter()

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
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else {"numa_node": None, "bound": False}
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
    # the same step with ORDINARY heap memory on both sides (what the reference's ndarray callers hold):
    # the library stages it through its pinned ring with worker threads; rank 0 only, wall clock
    pageable = None
    if rank == 0:
        pg_best, pg_dmin = np.empty(N_ROWS, np.uint32), np.empty(N_ROWS, np.float32)
        pg_mem = np.zeros(members_cap["buf"].size, np.uint64)
        from spfresh_b200._capi import check, lib, ptr

        def step_pageable():
            d2, r = spf.Dataset.assign_from_host(ctx, rows_np, spf.METRIC_EUCLIDEAN, cent)
            check(lib().spf_assign_fetch(r.handle, ptr(pg_best), ptr(pg_dmin), ptr(out_off), ptr(pg_mem)))
            r.free()
            d2.free()
        times = {}
        for name, flag in (("staging_ring", 0), ("driver_staged", 1)):
            ctx.set_param("no_host_staging", flag)
            step_pageable()
            t0 = time.perf_counter()
            for _ in range(3):
                step_pageable()
            times[name] = (time.perf_counter() - t0) / 3 * 1e3
        ctx.set_param("no_host_staging", 0)
        pageable = {"value": N_ROWS / (times["staging_ring"] * 1e-3), "unit": "points/s",
                    "ms_per_step": times["staging_ring"], "ms_per_step_driver_staged": times["driver_staged"],
                    "same_results_as_pinned": bool(np.array_equal(pg_best, out_best) and
                                                   np.array_equal(pg_dmin.view(np.uint32), out_dmin.view(np.uint32))),
                    "note": "rows, best, dmin and member lists in ordinary (pageable) numpy arrays; worker threads copy "
                            "4 MB blocks through 12 pinned slots in both directions (assign_api.cu: staged_upload / "
                            "staged_download); driver_staged = the same call with cudaMemcpyAsync on the heap pointers"}

    # ---- one full row-sharded k-means iteration, device resident (spf_kmeans): assign + update_centroids,
    # the two exchanges of the update as NCCL all-gathers on the library's stream -----------------------
    comm = spf.DeviceComm.from_torch(ctx)
    sharded = bench_kmeans_iteration(spf, ctx, ds, comm, rank, world, rows_np, torch, dist, dev, ext)
    sharded_check = None
    if world > 1:
        # N > 1 correctness inside the run: the NCCL iteration against the host-staged exchange over the
        # same shards (rank 0 re-creates them), and the list-sharded query against the unsharded one
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from check_sharded_nccl import check_kmeans
        from check_sharded_query import check_query
        rows_checked = check_kmeans(ctx, comm, rank, world, dev, n=40_000, d=DIM, k=512, iters=2)
        qmsg = check_query(ctx, comm, rank, world, dev, n=120_000, d=DIM, k_lists=512, nq=4096, nprobes=(8, 32))
        sharded_check = {"ok": True, "kmeans": f"{rows_checked} rows over {world} ranks, 2 iterations: NCCL device-resident "
                                               "iteration == host-staged exchange over the same shards (rows + vector bits)",
                         "query": qmsg + ": list-sharded merged top-k == unsharded (ids, distance bits, counts)"}

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    tf32_live, tf32_sustained = measure_tf32_peak(torch, dev) if rank == 0 else (None, None)
    bf16 = peaks.get("bf16_tflops")
    if tf32_live:
        peak_tf, peak_src = tf32_live, "cuBLAS TF32 8192^3 best-of-10 (burst) measured in this run; the kernel is timed alone"
    elif bf16:
        peak_tf, peak_src = float(bf16) / 2, "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 runs at half the bf16 rate)"
    else:
        peak_tf, peak_src = 1590.0 / 2, "fallback 1.59 PFLOP/s bf16 / 2"
    flops = 2.0 * N_ROWS * K_CENT * DIM
    ach_tf = flops / (kms["assign_tc"] * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "assign_tc_kernel (tcgen05 kind::tf32, 1 pass)",
                "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                "traffic": TC_DRAM_BYTES_PER_LAUNCH, "traffic_source": "ncu dram__bytes_read+write.sum of this kernel, "
                "profiles/r02_ncu_assign_tc_v1.txt (1M-row launch: 0.52 GB read + 1.35 GB written)", "peak_source": peak_src,
                "tf32_peaks_this_run": {"burst_tflops": tf32_live, "sustained_4s_tflops": tf32_sustained,
                                        "bf16_burst_over_2": (float(bf16) / 2) if bf16 else None,
                                        "frac_of_sustained": (ach_tf / tf32_sustained) if tf32_sustained else None,
                                        "frac_of_bf16_burst_over_2": (ach_tf / (float(bf16) / 2)) if bf16 else None},
                "flop_per_launch": flops,
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

    # ---- the other BASELINE configs, compact (3: GIST L1 / Linf, 4: 100M x 96 strong scaling, 5: query sweep) ----
    configs = {}
    if not args.no_configs:
        ds.free()                                  # the headline dataset is no longer needed

        def release():
            ctx.trim()                             # the library's pool and torch's allocator share the device
            torch.cuda.empty_cache()
        release()
        if world == 1:
            for name, fn in (("sweep", lambda: bench_sweep(spf, ctx, comm, rank, world, torch, dist, dev, ext, hbm_peak, not args.no_cpu)),
                             ("gist", lambda: bench_gist(spf, ctx, torch, dev, not args.no_cpu)),
                             ("deep_strong", lambda: bench_deep(spf, ctx, comm, rank, world, torch, dist, dev, ext))):
                try:
                    r_ = fn()
                    if name == "gist":
                        configs.update(r_)
                    else:
                        configs[name] = r_
                except Exception as ex:            # the headline must still print
                    configs[name] = {"error": repr(ex)}
                release()
        else:
            # collective blocks: an exception ends the job (no rank is left waiting in a collective)
            configs["sweep"] = bench_sweep(spf, ctx, comm, rank, world, torch, dist, dev, ext, hbm_peak, not args.no_cpu)
            release()
            if rank == 0:
                configs.update(bench_gist(spf, ctx, torch, dev, not args.no_cpu))
            release()
            dist.barrier()
            configs["deep_strong"] = bench_deep(spf, ctx, comm, rank, world, torch, dist, dev, ext)
            release()

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
                            "device -> host fetch of the CSR",
                    "pageable_host_buffers": pageable},
            "gpu_launches": int(launches),
            "host_numa": numa,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity_checked_rows": parity["parity_checked_rows"], "parity_ok": parity["parity_ok"],
            "parity_detail": parity.get("parity_detail"),
            "kernels_ms": kms,
            "sharded_kmeans_iteration": sharded,
            "sharded_check": sharded_check,
            "query": query,
            "configs": configs,
            "wall_s": round(time.perf_counter() - t_main, 1),
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and parity["parity_ok"] is False:
        raise SystemExit("bench.py: the GPU result of the timed step differs from the CPU oracle")


def max_over_ranks(torch, dist, dev, world, *vals):
    if world == 1:
        return vals if len(vals) > 1 else vals[0]
    t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = [float(x) for x in t.tolist()]
    return out if len(out) > 1 else out[0]


def all_ranks(torch, dist, dev, world, val):
    """The value of every rank, in rank order."""
    if world == 1:
        return [float(val)]
    t = torch.tensor([float(val)], dtype=torch.float64, device=dev)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return [float(x.item()) for x in parts]


def bench_kmeans_iteration(spf, ctx, ds, comm, rank, world, rows_np, torch, dist, dev, ext, iters=4):
    """HierarchicalClustering assign_points + update_centroids (hierarchical.rs:368-390, 138-181), rows
    sharded (1M per GPU, weak), state resident on the device, exchanges over NCCL."""
    from spfresh_b200.sharded import DeviceShardedKMeans
    km = DeviceShardedKMeans(ds, comm, spf.METRIC_EUCLIDEAN, rank * N_ROWS)
    init_rows = np.arange(K_CENT, dtype=np.uint64)
    km.init(init_rows, make_rows(0, K_CENT) if rank else rows_np[:K_CENT])    # rows 0..k-1 are shared by all ranks
    km.step()                                      # unseeded first iteration (also warms NCCL up)
    km.step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(iters):
        km.step()
    e1.record(ext)
    e1.synchronize()
    ms = max_over_ranks(torch, dist, dev, world, e0.elapsed_time(e1) / iters)
    ctx.set_profiling(True)
    km.step()
    names = ["assign_tc", "resolve", "cc_matrix", "csr", "kmeans_sums", "kmeans_exchange", "kmeans_means", "kmeans_medoid"]
    parts = {n: max(ctx.kernel_ms(n), 0.0) for n in names}
    ctx.set_profiling(False)
    exch = max_over_ranks(torch, dist, dev, world, parts["kmeans_exchange"])
    _, _, sizes = km.centroids()
    out = {"ms_per_iteration": ms, "points_per_s": N_ROWS * world / (ms * 1e-3), "n_gpus": world, "rows_per_gpu": N_ROWS,
           "scaling": "weak", "timing": "CUDA events on the library stream, max over ranks",
           "collective_ms": exch, "collective": "2 NCCL all-gathers per iteration on the library stream (C1: k x (d+1) partial "
           "sums + counts, 2.1 MB per rank; C2: best member per cluster + its vector, 2.2 MB per rank); rank-ordered f32 sums",
           "backend": "nccl (library-owned communicator)" if world > 1 else "none (one rank)",
           "kernels_ms": parts, "seeded": "iterations after the first seed the candidate pass with d(x, c_new[best_old])",
           "global_members": int(sizes.sum())}
    km.free()
    # k-means++ rounds over the same shards, resident on the devices (spf_kmpp_rounds_sharded: three small
    # all-gathers per round, one host synchronisation per batch); CUDA events on the library stream
    from spfresh_b200.device import KmppShardSession
    sess = KmppShardSession(ds, spf.METRIC_EUCLIDEAN)
    sess.set_vector(rows_np[7] if rank == 0 else make_rows(0, K_CENT)[7])      # row 7 is shared by all ranks
    u = np.random.Generator(np.random.Philox(key=77)).random(64 + 8)
    sess.rounds_sharded(comm, rank * N_ROWS, u[:8])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(ext)
    picked, failed = sess.rounds_sharded(comm, rank * N_ROWS, u[8:])
    k1.record(ext)
    k1.synchronize()
    kms = k0.elapsed_time(k1) / 64
    if world > 1:
        t = torch.tensor([kms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kms = float(t.item())
    sess.free()
    out["kmeanspp_round"] = {"ms_per_round": kms, "rounds": 64, "rows_per_gpu": N_ROWS, "failed": bool(failed),
                             "distinct_rows": int(len(set(int(x) for x in picked))),
                             "note": "hierarchical.rs:259-291 over row shards: running-minimum update at HBM rate, bit-exact "
                                     "sequential f32 sum per shard (thread-block-cluster scan), sums and f64 weight totals "
                                     "all-gathered and added in rank order, owner picks, vector all-gathered"}
    return out


def device_clustered(torch, dev, n, d, ncent, seed_c, seed_x, scale=0.5):
    """Distribution B of SURVEY 8(d) generated on the device: x = centre[label] + 0.5 N(0, I)."""
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed_c)
    centres = 2.0 * torch.randn((ncent, d), generator=gen, device=dev)
    gen.manual_seed(seed_x)
    x = torch.empty((n, d), dtype=torch.float32, device=dev)
    step = 1 << 21
    for i in range(0, n, step):
        m = min(step, n - i)
        lab = torch.randint(0, ncent, (m,), generator=gen, device=dev)
        x[i:i + m] = centres[lab] + scale * torch.randn((m, d), generator=gen, device=dev)
    torch.cuda.synchronize()
    return x


def bench_gist(spf, ctx, torch, dev, with_cpu):
    """BASELINE config 3: 1M x 960 f32, k = 4096, Manhattan and Chebyshev on the CUDA-core direct-form
    kernel (distance.rs:25-43 behind hierarchical.rs:302-326).  FP32-issue bound: 2 lane instructions per
    element-op.  A row sample is checked against the CPU oracle at the full dimension."""
    import oracle
    n, d, k = 1_000_000, 960, K_CENT
    x = device_clustered(torch, dev, n, d, 1024, 45, 46)
    ds = spf.Dataset(ctx, device_ptr=x.data_ptr(), n=n, d=d)
    del x
    torch.cuda.empty_cache()
    cent = np.random.Generator(np.random.Philox(key=7)).choice(n, k, replace=False).astype(np.uint64)
    sample = np.sort(np.random.Generator(np.random.Philox(key=8)).choice(n, 1024, replace=False)).astype(np.uint64)
    host_small = ds.fetch_rows(np.concatenate([cent, sample])) if with_cpu else None
    out = {}
    ctx.set_profiling(True)
    for metric, name in ((spf.METRIC_MANHATTAN, "manhattan"), (spf.METRIC_CHEBYSHEV, "chebyshev")):
        sampler = ClockSampler(ctx_device(ctx))
        sampler.start()
        best_ms, kern_ms, other = 1e30, 0.0, {}
        for _ in range(2):
            t0 = time.perf_counter()
            r = ds.assign(metric, cent)
            dt = (time.perf_counter() - t0) * 1e3
            if dt < best_ms:
                best_ms, kern_ms = dt, ctx.kernel_ms("assign_exact")
                other = {x_: ctx.kernel_ms(x_) for x_ in ("resolve", "cc_matrix", "csr", "overflow")}
            members, ovf = r.total, ctx.last_overflow_rows()
            gpu = r.fetch(best=True, dmin=True, csr=False) if with_cpu else None
            r.free()
        clocks = sampler.stop()
        mhz = clocks.get("sm_mhz") or 1965.0
        lane = 2.0 * n * k * d
        rec = {"rows": n, "dim": d, "k": k, "data": "clustered (1024 centres), generated on the device",
               "assign_exact_kernel_ms": kern_ms, "call_ms": best_ms, "points_per_s": n / (best_ms * 1e-3),
               "other_kernels_ms": other, "members": int(members), "overflow_rows": int(ovf),
               "roofline": {"bound": "fp32 issue", "lane_instr_per_launch": lane, "unit": "T lane-instr/s",
                            "achieved": lane / (kern_ms * 1e-3) / 1e12,
                            "peak_at_max_clock": 148 * 128 * 1.965e9 / 1e12, "frac_of_max_clock_peak": lane / (kern_ms * 1e-3) / (148 * 128 * 1.965e9),
                            "sm_mhz_under_load": mhz, "frac_at_measured_clock": lane / (kern_ms * 1e-3) / (148 * 128 * mhz * 1e6),
                            "note": "2 lane instructions per element-op (FADD + FADD|x| / FMNMX|x|); ncu: profiles/r02_ncu_assign_exact_*.txt"}}
        if with_cpu:
            # the same rows on the CPU oracle (centroids = rows 0..k-1 of the small host copy)
            t0 = time.perf_counter()
            ref = oracle.assign(host_small, metric, np.arange(k, dtype=np.uint64),
                                point_idx=np.arange(k, k + sample.size, dtype=np.uint64))
            t_cpu = time.perf_counter() - t0
            ok = bool(np.array_equal(gpu.best[sample.astype(np.int64)], ref.best)
                      and np.array_equal(gpu.dmin[sample.astype(np.int64)].view(np.uint32), ref.dmin.view(np.uint32)))
            rec["cpu_baseline"] = {"value": sample.size / t_cpu, "unit": "points/s", "cores": oracle.online_cpus(), "kind": "port",
                                   "sample": f"{sample.size} rows x 4096 centroids x 960 dims in {t_cpu:.2f} s"}
            rec["parity_checked_rows"] = int(sample.size)
            rec["parity_ok"] = ok
        out["gist_" + name] = rec
    ctx.set_profiling(False)
    ds.free()
    return out


def ctx_device(ctx):
    from spfresh_b200._capi import lib
    return int(lib().spf_ctx_device(ctx.handle))


def bench_deep(spf, ctx, comm, rank, world, torch, dist, dev, ext, rows_total=100_000_000):
    """BASELINE config 4: 100M x 96 f32 (distribution B, 4096 centres), k = 4096, rows sharded over the
    ranks (strong scaling: rows_total is fixed), device-resident k-means iterations over NCCL."""
    from spfresh_b200.sharded import DeviceShardedKMeans
    d, k = 96, K_CENT
    free_b, _ = torch.cuda.mem_get_info()
    need = lambda rows: rows * (d * 4 * 3 + 64 * 4 + 48) + (12 << 30)      # rows (+ torch copy + TF32 copy), member lists, scratch
    note = None
    while need(rows_total // world) > free_b and rows_total > 8_000_000:
        rows_total //= 2
        note = "rows_total reduced to fit the free device memory"
    lo, hi = rank * rows_total // world, (rank + 1) * rows_total // world
    n = hi - lo
    x = device_clustered(torch, dev, n, d, 4096, 1234, 99 + rank)
    ds = spf.Dataset(ctx, device_ptr=x.data_ptr(), n=n, d=d)
    # the k initial centroids: rows of rank 0's shard (every rank regenerates them from the seeds)
    init_rows = np.sort(np.random.Generator(np.random.Philox(key=7)).choice(rows_total // world, k, replace=False)).astype(np.uint64)
    if rank == 0:
        vec = ds.fetch_rows(init_rows)
    del x
    torch.cuda.empty_cache()
    vt = torch.zeros((k, d), dtype=torch.float32, device=dev)
    if rank == 0:
        vt.copy_(torch.from_numpy(vec))
    if world > 1:
        dist.broadcast(vt, src=0)
    vec = vt.cpu().numpy()
    km = DeviceShardedKMeans(ds, comm, spf.METRIC_EUCLIDEAN, lo)
    km.init(init_rows, vec)
    km.step()                                                   # first iteration: unseeded (largest member lists)
    km.step()
    ctx.set_profiling(True)
    km.step()
    names = ["assign_tc", "classify", "exact_eval", "finalize", "resolve", "cc_matrix", "csr", "csr_scan", "csr_fill", "csr_sort",
             "overflow", "kmeans_seed", "kmeans_sums", "kmeans_exchange", "kmeans_means", "kmeans_medoid"]
    parts = {nm: max(ctx.kernel_ms(nm), 0.0) for nm in names}
    ctx.set_profiling(False)
    times = []
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        km.step()
        e1.record(ext)
        e1.synchronize()
        times.append(max_over_ranks(torch, dist, dev, world, e0.elapsed_time(e1)))
    ms = float(np.median(times))
    tc_ms, exch = max_over_ranks(torch, dist, dev, world, parts["assign_tc"], parts["kmeans_exchange"])
    _, _, sizes = km.centroids()
    flop = 2.0 * rows_total * k * d
    out = {"rows_total": rows_total, "rows_per_gpu": n, "dim": d, "k": k, "n_gpus": world, "scaling": "strong",
           "data": "clustered (4096 centres), generated on the device", "kmeans_iteration_ms": ms,
           "iteration_ms_all": times, "timing": "median of 3 iterations, each CUDA-event timed, max over ranks",
           "iteration_points_per_s": rows_total / (ms * 1e-3), "assign_tc_ms": tc_ms,
           "assign_tc_tflops_all_gpus": flop / (tc_ms * 1e-3) / 1e12, "collective_ms": exch,
           "rank0_kernels_ms": parts, "global_members": int(sizes.sum()), "note": note}
    km.free()
    ds.free()
    torch.cuda.empty_cache()
    return out


def bench_sweep(spf, ctx, comm, rank, world, torch, dist, dev, ext, hbm_peak, with_cpu):
    """BASELINE config 5: 100k queries, top-10, nprobe 8 / 32 / 256 over the config-2 index (1M x 128 N(0,1),
    4096 lists, HBM resident).  N > 1: the posting lists are sharded over the ranks (balanced by vectors),
    every rank brings 100k / N queries in pinned host memory and receives their merged top-10
    (spf_search_sharded).  QPS is end to end: host queries in, host results out, max over ranks."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from check_sharded_query import balanced_list_ranges
    rows0 = make_rows(0)
    ds = spf.Dataset(ctx, rows0)
    cent = np.arange(K_CENT, dtype=np.uint64)
    res = ds.assign(spf.METRIC_EUCLIDEAN, cent)
    f = res.fetch(best=False, dmin=False)
    med = ds.update_medoids_from(spf.METRIC_EUCLIDEAN, res, cent)
    res.free()
    lb, le = balanced_list_ranges(f.offsets, world)[rank]
    idx = spf.DeviceIndex.pack(ds, f.offsets, f.members, med, list_range=(lb, le))
    # N > 1 and the whole index fits one GPU (4.8 GB here): the other deployment, every rank holds all
    # lists and answers its own slice of the queries — no exchange at all ("replicas only")
    idx_full = spf.DeviceIndex.pack(ds, f.offsets, f.members, med) if world > 1 else None
    nq_total = 100_000 // world * world
    nql = nq_total // world
    q_all = np.random.Generator(np.random.Philox(key=46)).standard_normal((nq_total, DIM), dtype=np.float32)
    q_pin = torch.from_numpy(q_all[rank * nql:(rank + 1) * nql].copy()).pin_memory()
    q = q_pin.numpy()
    o_ids = torch.empty((nql, TOPK), dtype=torch.int64).pin_memory()
    o_d = torch.empty((nql, TOPK), dtype=torch.float32).pin_memory()
    o_c = torch.empty((nql,), dtype=torch.int32).pin_memory()
    out_bufs = (o_ids.numpy().view(np.uint64), o_d.numpy(), o_c.numpy().view(np.uint32))
    # ground truth for recall (first 500 queries of rank 0's slice, fp32 brute force on the device)
    gt = None
    if rank == 0:
        xq = torch.from_numpy(q[:500]).to(dev)
        x = torch.from_numpy(rows0).to(dev)
        d2 = (xq * xq).sum(1, keepdim=True) - 2.0 * xq @ x.T + (x * x).sum(1)[None, :]
        gt = torch.topk(d2, TOPK, dim=1, largest=False).indices.cpu().numpy()
        del x, xq, d2
        torch.cuda.empty_cache()
    points = {}
    for nprobe in (8, 32, 256):
        for _ in range(2):
            idx.search_sharded(comm, q, TOPK, nprobe=nprobe, out=out_bufs)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        reps = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(reps):
            ids, dists, counts = idx.search_sharded(comm, q, TOPK, nprobe=nprobe, out=out_bufs)
        e1.record(ext)
        e1.synchronize()
        ms = max_over_ranks(torch, dist, dev, world, e0.elapsed_time(e1) / reps)
        ctx.set_profiling(True)
        idx.search_sharded(comm, q, TOPK, nprobe=nprobe, out=out_bufs)
        km = {nm: max(ctx.kernel_ms(nm), 0.0) for nm in ("probe", "scan", "exchange", "merge", "scan_tc_a")}
        passes = {nm: round(max(ctx.kernel_ms("scan_tc_" + nm), 0.0), 3) for nm in ("gather", "a", "tau", "b", "flag", "refine", "fallback")}
        stream_mb = ctx.kernel_ms("scan_tc_unique_mb")
        ctx.set_profiling(False)
        scan_ms, exch_ms = max_over_ranks(torch, dist, dev, world, km["scan"], km["exchange"])
        rec = {"nprobe": nprobe, "nq": nq_total, "k": TOPK, "qps_e2e": nq_total / (ms * 1e-3), "ms_per_batch": ms,
               "scan_ms_max": scan_ms, "probe_ms": km["probe"], "exchange_ms_max": exch_ms, "merge_ms": km["merge"],
               "h2d_bytes_per_rank": int(q.nbytes), "d2h_bytes_per_rank": int(sum(o.nbytes for o in out_bufs))}
        rec["scan_passes_ms_this_rank"] = passes
        if world > 1:
            rec["scan_ms_all_ranks"] = [round(v, 3) for v in all_ranks(torch, dist, dev, world, km["scan"])]
            sh = (ids.copy(), dists.copy(), counts.copy())
            for _ in range(2):
                idx_full.search(q, TOPK, nprobe=nprobe, out=out_bufs)
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            for _ in range(reps):
                ids, dists, counts = idx_full.search(q, TOPK, nprobe=nprobe, out=out_bufs)
            e1.record(ext)
            e1.synchronize()
            ms_rep = max_over_ranks(torch, dist, dev, world, e0.elapsed_time(e1) / reps)
            same = bool(np.array_equal(sh[2], counts) and np.array_equal(sh[0], ids)
                        and np.array_equal(sh[1].view(np.uint32), dists.view(np.uint32)))
            rec["replicated_lists"] = {"qps_e2e": nq_total / (ms_rep * 1e-3), "ms_per_batch": ms_rep,
                                       "identical_to_list_sharded": same,
                                       "what": "every rank holds all lists and searches its own query slice (spf_search_batch), no exchange"}
        if km["scan_tc_a"] > 0 and stream_mb > 0:
            ach = stream_mb * 1e6 / (km["scan_tc_a"] * 1e-3) / 1e9
            rec["bound_pass"] = {"kernel_ms": km["scan_tc_a"], "list_bytes_streamed_once": int(stream_mb * 1e6),
                                 "gbs": ach, "frac_of_hbm_peak": ach / hbm_peak,
                                 "note": "HBM-bound while a list is probed by <= 128 queries (one unit); tensor-bound beyond"}
        if rank == 0:
            hit = sum(len(set(gt[i].tolist()) & set(ids[i, :counts[i]].tolist())) for i in range(500))
            rec["recall_at_10"] = hit / (500.0 * TOPK)
            if with_cpu and nprobe == 32:
                import oracle
                t0 = time.perf_counter()
                rid, rd, rc = oracle.search_batch(rows0, f.offsets, f.members, med, q[:512], TOPK, nprobe=nprobe)
                t_cpu = time.perf_counter() - t0
                ok = bool(np.array_equal(counts[:512], rc) and all(
                    np.array_equal(ids[i, :rc[i]], rid[i, :rc[i]]) and
                    np.array_equal(dists[i, :rc[i]].view(np.uint32), rd[i, :rc[i]].view(np.uint32)) for i in range(512)))
                rec["cpu_baseline"] = {"value": 512 / t_cpu, "unit": "queries/s", "cores": oracle.online_cpus(), "kind": "port",
                                       "sample": f"512 of the 100k queries in {t_cpu:.2f} s (in-memory lists, no file I/O)"}
                rec["parity_checked_queries"] = 512
                rec["parity_ok"] = ok
        points[f"nprobe_{nprobe}"] = rec
    out = {"index": "config-2 assignment (1M x 128 N(0,1), 4096 lists, boundary replicas kept)", "n_gpus": world,
           "index_vectors_this_rank": idx.nvectors, "lists_this_rank": [int(lb), int(le)],
           "sharding": "posting lists by contiguous list range balanced by vectors; queries sharded for upload / probe / merge",
           "recall_note": "recall_at_10 counts distinct true neighbours among the returned ids; the reference keeps boundary "
                          "replicas of a point as separate results (spann_index.rs:168-193, no de-duplication), so with 9.4 "
                          "replicas per point on this data more probes put more duplicates into the top-10 and recall falls "
                          "with nprobe; the values equal the oracle's",
           "points": points}
    idx.free()
    if idx_full is not None:
        idx_full.free()
    ds.free()
    return out


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
                                                      "units", "stream_mb", "unique_mb", "flag", "groups", "split", "hub_units",
                                                      "hub_tiles", "tiles")}
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
                       "passes_ms": {"gather": tc["gather"], "bound": tc["a"], "tau": tc["tau"], "group_refine": tc["b"], "group_refine_flagging": tc["flag"], "select": tc["refine"],
                                     "fallback": tc["fallback"]},
                       "units": int(tc["units"]),
                       "split_launch": {"sms_multi_unit_lists": int(max(tc["split"], 0)), "multi_unit_units": int(max(tc["hub_units"], 0)),
                                        "multi_unit_tiles": int(max(tc["hub_tiles"], 0)), "tiles": int(max(tc["tiles"], 0)),
                                        "what": "the bound pass runs as two concurrent persistent launches on disjoint SMs: units of lists "
                                                "probed by more than 128 pairs (tensor-bound, L2 re-reads) and single-unit lists (HBM-bound); "
                                                "kernel_ms spans both"},
                       "subgroups_refined_per_query": tc["groups"] / NQ, "candidates_per_query": tc["candidates"] / NQ,
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
