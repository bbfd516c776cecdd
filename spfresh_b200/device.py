"""Object wrappers over the C ABI handles (context, dataset, assignment result, index)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi as capi
from ._capi import as_f32, as_u64, check, lib, ptr


class Context:
    """spf_ctx: one B200 + its stream.  `Context.default()` is a per-process singleton."""
    _default = None

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        check(lib().spf_ctx_create(device, C.byref(h)))
        self._h = h

    @classmethod
    def default(cls) -> "Context":
        if cls._default is None:
            cls._default = Context(0)
        return cls._default

    def close(self):
        if self._h:
            lib().spf_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def stream(self) -> int:
        return int(lib().spf_ctx_stream(self._h) or 0)

    def synchronize(self):
        check(lib().spf_ctx_synchronize(self._h))

    def trim(self):
        """Return the unused part of the library's device memory pool to the driver."""
        check(lib().spf_ctx_trim(self._h))

    def set_profiling(self, on: bool):
        check(lib().spf_ctx_set_profiling(self._h, 1 if on else 0))

    def kernel_ms(self, name: str) -> float:
        return float(lib().spf_ctx_kernel_ms(self._h, name.encode()))

    def launch_count(self) -> int:
        return int(lib().spf_ctx_launch_count(self._h))

    def last_overflow_rows(self) -> int:
        return int(lib().spf_ctx_last_overflow_rows(self._h))

    def seq_sum_f32(self, values, mode: int = 1) -> np.float32:
        """spf_seq_sum_f32 (test hook): the sequential f32 fold, mode 1 = scan kernel, 2 = serial chain."""
        v = as_f32(values).reshape(-1)
        out = C.c_float()
        check(lib().spf_seq_sum_f32(self._h, ptr(v), v.size, int(mode), C.byref(out)))
        return np.float32(out.value)

    def set_param(self, name: str, value: int):
        check(lib().spf_ctx_set_param(self._h, name.encode(), int(value)))

    def distance_pairs(self, metric: int, a, b) -> np.ndarray:
        """Batched DistanceMetric::compute (distance.rs:7-43)."""
        a, b = as_f32(a), as_f32(b)
        if a.shape != b.shape:
            raise ValueError("shape mismatch")   # the reference panics (ShapeMismatch.unwrap())
        a2, b2 = a.reshape(-1, a.shape[-1]), b.reshape(-1, b.shape[-1])
        out = np.empty(a2.shape[0], np.float32)
        check(lib().spf_distance_pairs(self._h, metric, ptr(a2), ptr(b2), a2.shape[1], a2.shape[0], ptr(out)))
        return out


class _CtxHolder:
    """Stands in for a Dataset as the owner reference of a result when the device copy was dropped."""

    def __init__(self, ctx):
        self.ctx = ctx


class Dataset:
    """spf_dataset: n x d f32 rows resident in HBM."""

    def __init__(self, ctx: Context, rows=None, *, device_ptr: int | None = None, n: int = 0, d: int = 0):
        self.ctx = ctx
        h = C.c_void_p()
        if device_ptr is not None:
            check(lib().spf_dataset_from_device(ctx.handle, C.c_void_p(device_ptr), n, d, C.byref(h)))
            self.n, self.d = n, d
        else:
            rows = np.asarray(rows)
            if rows.ndim != 2:
                raise ValueError("rows must be 2-D")
            if rows.dtype != np.float32 or rows.strides[1] != 4 or rows.strides[0] % 4 != 0 or rows.strides[0] < 4 * rows.shape[1]:
                rows = np.ascontiguousarray(rows, dtype=np.float32)
            self.n, self.d = rows.shape
            check(lib().spf_dataset_upload(ctx.handle, ptr(rows), self.n, self.d, rows.strides[0] // 4, C.byref(h)))
        self._h = h

    @property
    def handle(self):
        return self._h

    @classmethod
    def assign_from_host(cls, ctx: Context, rows, metric: int, centroid_rows, boundary_factor: float = 1.1,
                         flags: int = capi.ASSIGN_DEFAULT, keep_dataset: bool = True):
        """spf_assign_host: upload (chunked, overlapped with compute) + assign of all rows in one call.
        Returns (dataset or None, AssignResult)."""
        rows = np.asarray(rows)
        if rows.ndim != 2:
            raise ValueError("rows must be 2-D")
        if rows.dtype != np.float32 or rows.strides[1] != 4 or rows.strides[0] % 4 != 0 or rows.strides[0] < 4 * rows.shape[1]:
            rows = np.ascontiguousarray(rows, dtype=np.float32)
        cr = as_u64(centroid_rows)
        hd, hr = C.c_void_p(), C.c_void_p()
        check(lib().spf_assign_host(ctx.handle, ptr(rows), rows.shape[0], rows.shape[1], rows.strides[0] // 4, metric,
                                    ptr(cr), cr.size, boundary_factor, flags,
                                    C.byref(hd) if keep_dataset else None, C.byref(hr)))
        ds = None
        if keep_dataset:
            ds = cls.__new__(cls)
            ds.ctx, ds._h = ctx, hd
            ds.n, ds.d = rows.shape
        holder = ds if ds is not None else _CtxHolder(ctx)
        return ds, AssignResult(holder, hr)

    def fetch_rows(self, rows) -> np.ndarray:
        rows = as_u64(rows)
        out = np.empty((rows.size, self.d), np.float32)
        check(lib().spf_dataset_fetch_rows(self._h, ptr(rows), rows.size, ptr(out)))
        return out

    def free(self):
        if self._h:
            lib().spf_dataset_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # ---- hot-path calls -------------------------------------------------------------------
    def assign(self, metric: int, centroid_rows, point_idx=None, boundary_factor: float = 1.1,
               flags: int = capi.ASSIGN_DEFAULT) -> "AssignResult":
        cr = as_u64(centroid_rows)
        if point_idx is None:
            pi, m = None, self.n
        else:
            pi = as_u64(point_idx)
            m = pi.size
        h = C.c_void_p()
        check(lib().spf_assign(self._h, metric, ptr(pi), m, ptr(cr), cr.size, boundary_factor, flags, C.byref(h)))
        return AssignResult(self, h)

    def assign_vectors(self, metric: int, centroids, point_idx=None, boundary_factor: float = 1.1,
                       flags: int = capi.ASSIGN_DEFAULT) -> "AssignResult":
        """spf_assign_vectors: centroids as explicit k x d vectors (row-sharded build)."""
        cv = as_f32(centroids).reshape(-1, self.d)
        if point_idx is None:
            pi, m = None, self.n
        else:
            pi = as_u64(point_idx)
            m = pi.size
        h = C.c_void_p()
        check(lib().spf_assign_vectors(self._h, metric, ptr(pi), m, ptr(cv), cv.shape[0], boundary_factor, flags,
                                       C.byref(h)))
        return AssignResult(self, h)

    def assign_balanced(self, metric: int, centroids, penalty=None, point_idx=None) -> "AssignResult":
        """spf_assign_balanced (extension, no reference counterpart): argmin_j fl(d(x, c_j) + penalty[j]),
        one cluster per point."""
        cv = as_f32(centroids).reshape(-1, self.d)
        pen = None if penalty is None else as_f32(penalty).reshape(cv.shape[0])
        if point_idx is None:
            pi, m = None, self.n
        else:
            pi = as_u64(point_idx)
            m = pi.size
        h = C.c_void_p()
        check(lib().spf_assign_balanced(self._h, metric, ptr(pi), m, ptr(cv), ptr(pen), cv.shape[0], 0, C.byref(h)))
        return AssignResult(self, h)

    def cluster_sums(self, result: "AssignResult"):
        sums = np.zeros((result.k, self.d), np.float32)
        counts = np.zeros(result.k, np.uint64)
        check(lib().spf_cluster_sums(self._h, result.handle, ptr(sums), ptr(counts)))
        return sums, counts

    def medoid_candidates(self, metric: int, result: "AssignResult", means):
        means = as_f32(means).reshape(result.k, self.d)
        dist = np.zeros(result.k, np.float32)
        row = np.zeros(result.k, np.uint64)
        check(lib().spf_medoid_candidates(self._h, metric, result.handle, ptr(means), ptr(dist), ptr(row)))
        return dist, row

    def update_medoids(self, metric: int, offsets, members, old_rows, want_means: bool = False):
        offsets, members, old_rows = as_u64(offsets), as_u64(members), as_u64(old_rows)
        k = old_rows.size
        new_rows = np.zeros(k, np.uint64)
        means = np.zeros((k, self.d), np.float32) if want_means else None
        check(lib().spf_update_medoids(self._h, metric, ptr(offsets), ptr(members), k, ptr(old_rows),
                                       ptr(new_rows), ptr(means)))
        return (new_rows, means) if want_means else new_rows

    def update_medoids_from(self, metric: int, result: "AssignResult", old_rows, want_means: bool = False):
        old_rows = as_u64(old_rows)
        k = old_rows.size
        new_rows = np.zeros(k, np.uint64)
        means = np.zeros((k, self.d), np.float32) if want_means else None
        check(lib().spf_update_medoids_from(self._h, metric, result.handle, ptr(old_rows), ptr(new_rows), ptr(means)))
        return (new_rows, means) if want_means else new_rows

    def farthest(self, metric: int, c1_row: int, members) -> int:
        members = as_u64(members)
        out = C.c_uint64()
        check(lib().spf_farthest(self._h, metric, int(c1_row), ptr(members), members.size, C.byref(out)))
        return int(out.value)

    def farthest_from(self, metric: int, c1_vector, members, skip_row=None):
        """spf_farthest_from (row-sharded bisect): (largest distance, earliest local member), or
        (0.0, None) when no member of this shard has a distance > 0."""
        members = as_u64(members)
        v = as_f32(c1_vector).reshape(self.d)
        dist, row = C.c_float(), C.c_uint64()
        skip = np.iinfo(np.uint64).max if skip_row is None else int(skip_row)
        check(lib().spf_farthest_from(self._h, metric, ptr(v), skip, ptr(members), members.size, C.byref(dist),
                                      C.byref(row)))
        r = int(row.value)
        return float(dist.value), (None if r == int(np.iinfo(np.uint64).max) else r)

    def kmeanspp(self, metric: int, first_row: int) -> "KmppSession":
        return KmppSession(self, metric, first_row)


class AssignResult:
    """spf_assign_result (device resident until fetched)."""

    def __init__(self, ds: Dataset, h):
        self.ds, self._h = ds, h
        self.m = int(lib().spf_assign_points(h))
        self.k = int(lib().spf_assign_clusters(h))
        self.total = int(lib().spf_assign_total(h))

    @property
    def handle(self):
        return self._h

    def fetch(self, best=True, dmin=True, csr=True):
        b = np.empty(self.m, np.uint32) if best else None
        dm = np.empty(self.m, np.float32) if dmin else None
        off = np.empty(self.k + 1, np.uint64) if csr else None
        mem = np.empty(self.total, np.uint64) if csr else None
        check(lib().spf_assign_fetch(self._h, ptr(b), ptr(dm), ptr(off), ptr(mem)))
        return Fetched(off, mem, b, dm)

    def free(self):
        if self._h:
            lib().spf_assign_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


@dataclass
class Fetched:
    offsets: np.ndarray | None
    members: np.ndarray | None
    best: np.ndarray | None
    dmin: np.ndarray | None

    def lists(self):
        return [self.members[int(self.offsets[j]):int(self.offsets[j + 1])] for j in range(len(self.offsets) - 1)]


class KmppSession:
    """spf_kmpp: device state of initialize_clusters_kmeans_plus_plus (hierarchical.rs:249-293)."""

    def __init__(self, ds: Dataset, metric: int, first_row: int):
        self.ds = ds
        h = C.c_void_p()
        check(lib().spf_kmpp_begin(ds.handle, metric, int(first_row), C.byref(h)))
        self._h = h

    def round(self, u01: float):
        """Returns the chosen row, or None when the weighted pick is impossible (caller then
        draws uniformly and calls push)."""
        out = C.c_uint64()
        rc = check(lib().spf_kmpp_round(self._h, float(u01), C.byref(out)))
        return None if rc == 1 else int(out.value)

    def rounds(self, u01):
        """spf_kmpp_rounds: len(u01) rounds on the device with one host synchronisation.  Returns
        (rows picked, failed): failed means round len(rows) could not pick (caller draws uniformly,
        calls push and carries on with the unused draws)."""
        u = np.ascontiguousarray(u01, np.float64)
        chosen = np.zeros(u.size, np.uint64)
        done = C.c_uint32()
        rc = check(lib().spf_kmpp_rounds(self._h, u.ctypes.data_as(C.POINTER(C.c_double)), u.size,
                                         chosen.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(done)))
        return chosen[:int(done.value)].copy(), rc == 1

    def push(self, row: int):
        check(lib().spf_kmpp_push(self._h, int(row)))

    def last_sums(self):
        s, t = C.c_float(), C.c_double()
        check(lib().spf_kmpp_last_sums(self._h, C.byref(s), C.byref(t)))
        return float(s.value), float(t.value)

    def free(self):
        if self._h:
            lib().spf_kmpp_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class KmppShardSession:
    """Row-sharded k-means++ state of one rank (spf_kmpp_begin_sharded / fold_vector / weight_total /
    pick_local)."""

    def __init__(self, ds: Dataset, metric: int):
        self.ds = ds
        h = C.c_void_p()
        check(lib().spf_kmpp_begin_sharded(ds.handle, metric, C.byref(h)))
        self._h = h

    def fold_vector(self, centroid) -> float:
        v = as_f32(centroid).reshape(self.ds.d)
        out = C.c_float()
        check(lib().spf_kmpp_fold_vector(self._h, ptr(v), C.byref(out)))
        return float(out.value)

    def weight_total(self, global_sum: float):
        """Returns (local f64 total, ok); ok is False when a local weight is invalid."""
        out = C.c_double()
        rc = check(lib().spf_kmpp_weight_total(self._h, float(np.float32(global_sum)), C.byref(out)))
        return float(out.value), rc == 0

    def set_vector(self, centroid):
        """spf_kmpp_set_vector: the newest centroid of the device-resident rounds."""
        v = as_f32(centroid).reshape(self.ds.d)
        check(lib().spf_kmpp_set_vector(self._h, ptr(v)))

    def rounds_sharded(self, comm, row_base: int, u01):
        """spf_kmpp_rounds_sharded (collective): len(u01) rounds without a host round trip.  Returns
        (global rows picked, failed); failed: round len(rows) could not pick and did not use its draw."""
        u = np.ascontiguousarray(u01, np.float64)
        chosen = np.zeros(u.size, np.uint64)
        done = C.c_uint32()
        rc = check(lib().spf_kmpp_rounds_sharded(self._h, comm.handle if comm is not None else None, int(row_base),
                                                 u.ctypes.data_as(C.POINTER(C.c_double)), u.size,
                                                 chosen.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(done)))
        return chosen[:int(done.value)].copy(), rc == 1

    def pick_local(self, target: float):
        out = C.c_uint64()
        rc = check(lib().spf_kmpp_pick_local(self._h, float(target), C.byref(out)))
        return None if rc == 1 else int(out.value)

    def free(self):
        if self._h:
            lib().spf_kmpp_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceIndex:
    """spf_index: posting lists + centroids resident in HBM."""

    def __init__(self, ctx: Context, h, d: int):
        self.ctx, self._h, self.d = ctx, h, d

    @classmethod
    def pack(cls, ds: Dataset, offsets, members, centroid_rows, list_range=None) -> "DeviceIndex":
        offsets, members, cr = as_u64(offsets), as_u64(members), as_u64(centroid_rows)
        lb, le = list_range if list_range is not None else (0, cr.size)
        h = C.c_void_p()
        check(lib().spf_index_pack(ds.handle, ptr(offsets), ptr(members), ptr(cr), cr.size, lb, le, C.byref(h)))
        return cls(ds.ctx, h, ds.d)

    @classmethod
    def load_dir(cls, ctx: Context, directory: str, centroids) -> "DeviceIndex":
        cen = as_f32(centroids)
        h = C.c_void_p()
        check(lib().spf_index_load_dir(ctx.handle, directory.encode(), ptr(cen), cen.shape[0], cen.shape[1], C.byref(h)))
        return cls(ctx, h, cen.shape[1])

    def save_dir(self, directory: str):
        check(lib().spf_index_save_dir(self._h, directory.encode()))

    @property
    def handle(self):
        return self._h

    @property
    def nlists(self) -> int:
        return int(lib().spf_index_lists(self._h))

    @property
    def nvectors(self) -> int:
        return int(lib().spf_index_vectors(self._h))

    def last_scan_bytes(self) -> int:
        return int(lib().spf_index_last_scan_bytes(self._h))

    def search(self, queries, k: int, nprobe: int = 0, prune_factor: float = 1.2, want_vectors=False,
               want_keys=False, out=None):
        """Batched find_k_nearest_neighbor_spann.  `out` = (ids, dists, counts) lets the caller supply
        the result arrays (e.g. views of pinned host memory); they are filled in place."""
        q = as_f32(queries).reshape(-1, self.d)
        nq = q.shape[0]
        if out is not None:
            ids, dists, counts = out
            if (ids.shape != (nq, k) or ids.dtype != np.uint64 or dists.shape != (nq, k) or dists.dtype != np.float32
                    or counts.shape != (nq,) or counts.dtype != np.uint32
                    or not (ids.flags.c_contiguous and dists.flags.c_contiguous and counts.flags.c_contiguous)):
                raise ValueError("out must be C-contiguous (nq,k) uint64, (nq,k) float32, (nq,) uint32")
        else:
            ids = np.empty((nq, k), np.uint64)
            dists = np.empty((nq, k), np.float32)
            counts = np.empty(nq, np.uint32)
        vec = np.empty((nq, k, self.d), np.float32) if want_vectors else None
        keys = np.empty((nq, k), np.uint64) if want_keys else None
        check(lib().spf_search_batch(self._h, ptr(q), nq, k, nprobe, prune_factor, ptr(ids), ptr(dists),
                                     ptr(counts), ptr(vec), ptr(keys)))
        out = [ids, dists, counts]
        if want_vectors:
            out.append(vec)
        if want_keys:
            out.append(keys)
        return tuple(out)

    def search_sharded(self, comm, queries, k: int, nprobe: int = 0, prune_factor: float = 1.2, out=None):
        """spf_search_sharded: this rank's slice of the batch against the list-sharded index of the
        group; returns the global top-k of these queries (collective)."""
        q = as_f32(queries).reshape(-1, self.d)
        nq = q.shape[0]
        if out is not None:
            ids, dists, counts = out
        else:
            ids = np.empty((nq, k), np.uint64)
            dists = np.empty((nq, k), np.float32)
            counts = np.empty(nq, np.uint32)
        check(lib().spf_search_sharded(self._h, comm.handle if comm is not None else None, ptr(q), nq, k, nprobe,
                                       prune_factor, ptr(ids), ptr(dists), ptr(counts)))
        return ids, dists, counts

    def free(self):
        if self._h:
            lib().spf_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def topk_merge(keys, ids, dists, counts):
    """Host merge of per-rank partial results: arrays are (parts, nq, k) / (parts, nq)."""
    keys, ids = as_u64(keys), as_u64(ids)
    dists = as_f32(dists)
    counts = np.ascontiguousarray(counts, np.uint32)
    parts, nq, k = keys.shape
    o_ids = np.empty((nq, k), np.uint64)
    o_d = np.empty((nq, k), np.float32)
    o_c = np.empty(nq, np.uint32)
    check(lib().spf_topk_merge(parts, nq, k, ptr(keys), ptr(ids), ptr(dists), ptr(counts), ptr(o_ids), ptr(o_d), ptr(o_c)))
    return o_ids, o_d, o_c


class DeviceComm:
    """spf_comm: this rank's handle of a multi-GPU group (NCCL over NVLink, one process per GPU).
    `id_bytes` is the 128-byte id rank 0 obtained from `unique_id()` and passed to every rank."""

    def __init__(self, ctx: Context, world: int = 1, rank: int = 0, id_bytes: bytes | None = None):
        self.ctx, self.world, self.rank = ctx, int(world), int(rank)
        h = C.c_void_p()
        buf = None
        if id_bytes is not None:
            buf = np.frombuffer(bytes(id_bytes), np.uint8).copy()
            if buf.size != 128:
                raise ValueError("the communicator id is 128 bytes")
        check(lib().spf_comm_create(ctx.handle, self.world, self.rank, ptr(buf), C.byref(h)))
        self._h = h

    @staticmethod
    def unique_id() -> bytes:
        buf = np.zeros(128, np.uint8)
        check(lib().spf_comm_unique_id(ptr(buf)))
        return buf.tobytes()

    @classmethod
    def from_torch(cls, ctx: Context) -> "DeviceComm":
        """Group over the ranks of the initialised torch.distributed process group: rank 0 creates
        the id, torch broadcasts the 128 bytes (host plumbing only; the data path is the library's
        own NCCL communicator on the library's stream)."""
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return cls(ctx, 1, 0, None)
        box = [cls.unique_id() if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(ctx, dist.get_world_size(), dist.get_rank(), box[0])

    @property
    def handle(self):
        return self._h

    def free(self):
        if self._h:
            lib().spf_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _BorrowedAssign(AssignResult):
    """The assignment owned by a KMeansSession (must not be freed by the wrapper)."""

    def free(self):
        self._h = None


class KMeansSession:
    """spf_kmeans: device-resident row-sharded k-means iterations (assign_points + update_centroids,
    hierarchical.rs:368-390, 138-181).  `comm` None = one GPU."""

    def __init__(self, ds: Dataset, comm: DeviceComm | None, metric: int, row0: int, k: int,
                 boundary_factor: float = 1.1, seeded: bool = True, balance_lambda: float | None = None,
                 lloyd_means: bool = False):
        self.ds, self.comm, self.k = ds, comm, int(k)
        flags = (0 if seeded else 1) | (2 if balance_lambda is not None else 0) | (4 if lloyd_means else 0)
        h = C.c_void_p()
        check(lib().spf_kmeans_create(ds.handle, comm.handle if comm is not None else None, metric, int(row0), self.k,
                                      boundary_factor, flags, C.byref(h)))
        self._h = h
        if balance_lambda is not None:
            check(lib().spf_kmeans_set_balance(self._h, float(balance_lambda)))

    def set_centroids(self, global_rows, vectors):
        rows = as_u64(global_rows)
        vec = as_f32(vectors).reshape(self.k, self.ds.d)
        if rows.size != self.k:
            raise ValueError("k rows expected")
        check(lib().spf_kmeans_set_centroids(self._h, ptr(rows), ptr(vec)))

    def step(self):
        check(lib().spf_kmeans_step(self._h))

    def fetch(self, rows=True, vectors=True, means=False, counts=True):
        r = np.empty(self.k, np.uint64) if rows else None
        v = np.empty((self.k, self.ds.d), np.float32) if vectors else None
        m = np.empty((self.k, self.ds.d), np.float32) if means else None
        c = np.empty(self.k, np.uint64) if counts else None
        check(lib().spf_kmeans_fetch(self._h, ptr(r), ptr(v), ptr(m), ptr(c)))
        return r, v, m, c

    def assignment(self) -> AssignResult:
        h = lib().spf_kmeans_assignment(self._h)
        if not h:
            raise RuntimeError("no assignment yet")
        return _BorrowedAssign(self.ds, C.c_void_p(h))

    def free(self):
        if self._h:
            lib().spf_kmeans_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
