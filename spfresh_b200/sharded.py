"""Row-sharded k-means build step: one process per GPU, every rank owns a
contiguous block of dataset rows, the k centroid vectors are replicated.

    assign            local (no exchange): spf_assign_vectors on the rank's shard
    update_centroids  hierarchical.rs:138-181 split at its two reductions
                        C1  per-cluster partial sums + counts      -> all-gather, summed in rank order
                        C2  per-cluster best local member (d, row) -> all-gather, minimum, lowest rank
                            wins ties (= the leftmost member, shards being contiguous row ranges)
                        the winners' vectors                       -> all-gather, selected per cluster

All messages are k x d floats or smaller (2 MB at k = 4096, d = 128), so the exchange is latency
bound; it goes through `torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  Partial sums are combined in rank order on every rank, which keeps the result identical
on all ranks and reproducible; it differs from the single-process mean only by the f32 rounding
of a different summation order (a documented near-tie class).

    subdivide_clusters / create_subclusters   hierarchical.rs:74-135: the serial work-list runs
                        identically on every rank from the global cluster sizes; a bisect draws c1
                        from the global member list, folds the farthest point per shard (combined
                        with strict > in rank order) and assigns the local slice to (c1, c2)

The shard object only has to provide `n`, `d`, `assign_vectors`, `cluster_sums`,
`medoid_candidates`, `farthest_from`, `member_lists` and `rows` — `DeviceShard` wraps a `Dataset` on the GPU; the CPU tests plug an
oracle-backed shard into the same exchange code.
"""
from __future__ import annotations

import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np


# ---------------------------------------------------------------------------------------------
# communicators
# ---------------------------------------------------------------------------------------------
class Comm:
    rank: int = 0
    world: int = 1

    def allgather(self, a: np.ndarray) -> List[np.ndarray]:
        """Every rank contributes an array of the same shape / dtype; returns them in rank order."""
        raise NotImplementedError

    def allgather_many(self, arrays: Sequence[np.ndarray]) -> List[List[np.ndarray]]:
        """Several arrays in ONE message (the exchange is latency bound): result[r][i] is rank r's
        i-th array."""
        arrays = [np.ascontiguousarray(a) for a in arrays]
        flat = [a.view(np.uint8).reshape(-1) for a in arrays]
        outs = self.allgather(np.concatenate(flat))
        res = []
        for o in outs:
            parts, pos = [], 0
            for a, f in zip(arrays, flat):
                parts.append(o[pos:pos + f.size].view(a.dtype).reshape(a.shape))
                pos += f.size
            res.append(parts)
        return res


class SingleComm(Comm):
    def allgather(self, a):
        return [np.array(a, copy=True)]


class TorchComm(Comm):
    """torch.distributed (NCCL: tensors staged on `device`; gloo: CPU tensors)."""

    def __init__(self, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device

    def allgather(self, a):
        import torch
        a = np.ascontiguousarray(a)
        view = a.view(np.uint8).reshape(-1)              # dtype-agnostic (uint64 has no torch dtype)
        t = torch.from_numpy(view.copy())
        if self.device is not None:
            t = t.to(self.device)
        outs = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(outs, t)
        return [o.cpu().numpy().view(a.dtype).reshape(a.shape) for o in outs]


class ThreadComm(Comm):
    """In-process ranks (threads) — used to exercise the exchange on one GPU / in CPU tests."""

    class Group:
        def __init__(self, world: int):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots: List[Optional[np.ndarray]] = [None] * world

    def __init__(self, group: "ThreadComm.Group", rank: int):
        self.g, self.rank, self.world = group, rank, group.world

    def allgather(self, a):
        self.g.slots[self.rank] = np.array(a, copy=True)
        self.g.barrier.wait()
        out = [np.array(x, copy=True) for x in self.g.slots]
        self.g.barrier.wait()
        return out


# ---------------------------------------------------------------------------------------------
# shards
# ---------------------------------------------------------------------------------------------
class DeviceShard:
    """A rank's rows resident on its B200 (spf_dataset) + the global id of its first row."""

    def __init__(self, dataset, row0: int, host_rows: Optional[np.ndarray] = None):
        self.ds, self.row0 = dataset, int(row0)
        self.n, self.d = dataset.n, dataset.d
        self._host = host_rows

    def assign_vectors(self, metric, centroids, boundary_factor=1.1, point_idx=None):
        return self.ds.assign_vectors(metric, centroids, point_idx=point_idx, boundary_factor=boundary_factor)

    def farthest_from(self, metric, c1_vector, members, skip_row=None):
        return self.ds.farthest_from(metric, c1_vector, members, skip_row)

    def member_lists(self, res):
        """Per cluster the shard-local rows assigned to it (input order)."""
        return res.fetch(best=False, dmin=False).lists()

    def cluster_sums(self, res):
        return self.ds.cluster_sums(res)

    def medoid_candidates(self, metric, res, means):
        return self.ds.medoid_candidates(metric, res, means)

    def kmpp_session(self, metric):
        from .device import KmppShardSession
        return KmppShardSession(self.ds, metric)

    def rows(self, local_rows: Sequence[int]) -> np.ndarray:
        if self._host is None:                       # shard created on the device: read the rows back
            return self.ds.fetch_rows(np.asarray(local_rows, np.uint64))
        return np.asarray(self._host[np.asarray(local_rows, np.int64)], np.float32)


# ---------------------------------------------------------------------------------------------
# the exchange
# ---------------------------------------------------------------------------------------------
def gather_rows(shard, comm: Comm, global_rows: np.ndarray, shard_starts: np.ndarray) -> np.ndarray:
    """Vectors of arbitrary global rows: the owner of each row contributes it."""
    global_rows = np.asarray(global_rows, np.int64)
    owner = np.searchsorted(shard_starts, global_rows, side="right") - 1
    mine = owner == comm.rank
    part = np.zeros((global_rows.size, shard.d), np.float32)
    if mine.any():
        part[mine] = shard.rows(global_rows[mine] - int(shard_starts[comm.rank]))
    parts = comm.allgather(part)
    out = np.zeros_like(part)
    for r in range(comm.world):
        sel = owner == r
        out[sel] = parts[r][sel]
    return out


def shard_layout(shard, comm: Comm) -> np.ndarray:
    """Global id of every rank's first row (contiguous row sharding)."""
    sizes = np.array([int(x[0]) for x in comm.allgather(np.array([shard.n], np.int64))], np.int64)
    return np.concatenate([[0], np.cumsum(sizes)[:-1]])


def update_centroids(shard, comm: Comm, metric: int, res, old_vectors: np.ndarray, old_rows: np.ndarray,
                     shard_starts: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """One sharded update_centroids.  Returns (new global rows, new vectors, global means)."""
    k, d = old_vectors.shape
    sums, counts = shard.cluster_sums(res)                                   # C1, local part
    parts = comm.allgather_many([np.asarray(sums, np.float32), np.asarray(counts, np.uint64)])
    tot = np.zeros((k, d), np.float32)
    cnt = np.zeros(k, np.uint64)
    for r in range(comm.world):                                              # fixed (rank) order
        tot = (tot + parts[r][0]).astype(np.float32)
        cnt = cnt + parts[r][1]
    means = np.zeros((k, d), np.float32)
    nz = cnt > 0
    means[nz] = (tot[nz] / cnt[nz].astype(np.float32)[:, None]).astype(np.float32)   # utils.rs:14

    dist, row = shard.medoid_candidates(metric, res, means)                  # C2, local part
    grow = np.where(row == np.uint64(np.iinfo(np.uint64).max), row, row + np.uint64(shard_starts[comm.rank]))
    cands = comm.allgather_many([np.asarray(dist, np.float32), np.asarray(grow, np.uint64)])
    best_d = np.full(k, np.inf, np.float32)
    best_row = np.zeros(k, np.uint64)                                        # identity (0, +inf) :163
    for r in range(comm.world):                                              # strict <: lowest rank wins ties
        better = cands[r][0] < best_d
        best_d[better] = cands[r][0][better]
        best_row[better] = cands[r][1][better]
    new_rows = np.where(nz, best_row, np.asarray(old_rows, np.uint64))       # empty cluster keeps its centroid :146-149
    new_vecs = gather_rows(shard, comm, new_rows, shard_starts)
    return new_rows, new_vecs, means


def kmeans_plus_plus_device(shard, comm: Comm, device_comm, metric: int, k: int, rng,
                            shard_starts: Optional[np.ndarray] = None, batch: int = 256):
    """The same initialisation with the rounds resident on the devices (spf_kmpp_rounds_sharded over
    the library's NCCL group `device_comm`; `comm` only gathers the vector of a uniformly drawn row):
    three small all-gathers per round on the library stream, one host synchronisation per batch.
    Same picks as kmeans_plus_plus() for the same draws."""
    starts = shard_layout(shard, comm) if shard_starts is None else shard_starts
    sizes = np.array([int(x[0]) for x in comm.allgather(np.array([shard.n], np.int64))], np.int64)
    n_total = int(sizes.sum())
    sess = shard.kmpp_session(metric)
    try:
        chosen = [int(rng.choose_index(n_total))]
        sess.set_vector(gather_rows(shard, comm, np.array([chosen[-1]], np.uint64), starts)[0])
        draws: List[float] = []
        while len(chosen) < k:
            want = min(k - len(chosen), batch)
            while len(draws) < want:
                draws.append(rng.uniform01())
            rows, failed = sess.rounds_sharded(device_comm, int(starts[comm.rank]), draws[:want])
            chosen.extend(int(r) for r in rows)
            del draws[:len(rows)]                                  # a failing round does not use its draw
            if failed:                                             # Err arm: uniform draw
                row = int(rng.choose_index(n_total))
                chosen.append(row)
                if len(chosen) < k:
                    sess.set_vector(gather_rows(shard, comm, np.array([row], np.uint64), starts)[0])
        return np.array(chosen, np.uint64)
    finally:
        sess.free()


def kmeans_plus_plus(shard, comm: Comm, metric: int, k: int, rng, shard_starts: Optional[np.ndarray] = None):
    """Row-sharded initialize_clusters_kmeans_plus_plus (hierarchical.rs:249-293).  `rng` is a
    RandomSource (clustering.py) that every rank seeds identically, so the draws need no exchange:
    choose_index(n_total) for the first centroid (:253-255) and the uniform fallback (:287-290),
    uniform01() for the weighted pick (:285-286).  Per round: every rank folds the newest centroid
    into its running minimum (local), the f32 sums are added in rank order (:278), every rank forms
    its f64 weight total with the global denominator (:279-282), and the rank whose weight range
    holds u * total picks inside its shard.  Returns the k global centroid rows."""
    starts = shard_layout(shard, comm) if shard_starts is None else shard_starts
    sizes = np.array([int(x[0]) for x in comm.allgather(np.array([shard.n], np.int64))], np.int64)
    n_total = int(sizes.sum())
    sess = shard.kmpp_session(metric)
    try:
        chosen = [int(rng.choose_index(n_total))]
        for _ in range(1, k):
            vec = gather_rows(shard, comm, np.array([chosen[-1]], np.uint64), starts)[0]
            s_local = np.float32(sess.fold_vector(vec))
            s = np.float32(0.0)
            for part in comm.allgather(np.array([s_local], np.float32)):          # rank order
                s = np.float32(s + part[0])
            t_local, ok = sess.weight_total(float(s))
            info = comm.allgather(np.array([t_local, 1.0 if ok else 0.0], np.float64))
            totals = np.array([x[0] for x in info], np.float64)
            total = np.float64(0.0)
            for t in totals:
                total = np.float64(total + t)
            all_ok = all(x[1] == 1.0 for x in info)
            row = None
            if all_ok and total > 0.0 and np.isfinite(total):
                u = np.float64(rng.uniform01()) * total
                prefix = np.float64(0.0)
                owner = comm.world - 1
                for r in range(comm.world - 1):
                    if u < prefix + totals[r]:
                        owner = r
                        break
                    prefix = np.float64(prefix + totals[r])
                else:
                    prefix = np.float64(sum(totals[:comm.world - 1], np.float64(0.0)))
                mine = -1
                if comm.rank == owner:
                    loc = sess.pick_local(float(u - prefix))
                    mine = -1 if loc is None else int(loc) + int(starts[owner])
                got = [int(x[0]) for x in comm.allgather(np.array([mine], np.int64))]
                row = got[owner] if got[owner] >= 0 else None
            if row is None:                                                       # Err arm: uniform draw
                row = int(rng.choose_index(n_total))
            chosen.append(row)
        return np.array(chosen, np.uint64)
    finally:
        sess.free()


class ShardedCluster:
    """A cluster of the row-sharded build: the global row of its centroid, this rank's slice of the
    member list (shard-local rows, input order), the member count of every rank, the depth."""

    def __init__(self, centroid_row: int, local_points: np.ndarray, counts: np.ndarray, depth: int):
        self.centroid_idx = int(centroid_row)
        self.local_points = np.asarray(local_points, np.uint64)
        self.counts = np.asarray(counts, np.int64)
        self.depth = int(depth)

    @property
    def size(self) -> int:
        return int(self.counts.sum())


def create_subclusters(shard, comm: Comm, metric: int, cluster: ShardedCluster, rng, shard_starts: np.ndarray,
                       boundary_factor: float = 1.1):
    """Row-sharded create_subclusters (hierarchical.rs:107-135): every rank works on its slice of the
    member list.  Shards are contiguous row ranges and member lists are in input order, so the
    global member list is the concatenation of the slices in rank order.
      c1   points.choose(rng) (:111): an index into the global list (identical draw on every rank);
           the owning rank turns it into a global row, its vector is gathered
      c2   the farthest-point fold (:112-126): local (max distance, earliest member) per rank,
           combined with strict > in rank order; no distance > 0 anywhere -> global row 0
      then assign_points_to_clusters with the two centroids on the local slice (:129)"""
    j = int(rng.choose_index(cluster.size))                                   # :111
    prefix = np.concatenate([[0], np.cumsum(cluster.counts)])
    owner = int(np.searchsorted(prefix, j, side="right") - 1)
    mine = -1
    if comm.rank == owner:
        mine = int(cluster.local_points[j - int(prefix[owner])]) + int(shard_starts[owner])
    c1 = [int(x[0]) for x in comm.allgather(np.array([mine], np.int64))][owner]
    v1 = gather_rows(shard, comm, np.array([c1], np.uint64), shard_starts)[0]
    skip = c1 - int(shard_starts[comm.rank]) if comm.rank == owner else None
    dist, row = shard.farthest_from(metric, v1, cluster.local_points, skip)
    grow = -1 if row is None else int(row) + int(shard_starts[comm.rank])
    cands = comm.allgather_many([np.array([dist], np.float32), np.array([grow], np.int64)])
    best_d, c2 = np.float32(0.0), 0                                            # identity (0, 0.0) :115
    for r in range(comm.world):                                                # strict >: lowest rank wins ties
        if int(cands[r][1][0]) >= 0 and cands[r][0][0] > best_d:
            best_d, c2 = cands[r][0][0], int(cands[r][1][0])
    vecs = gather_rows(shard, comm, np.array([c1, c2], np.uint64), shard_starts)
    res = shard.assign_vectors(metric, vecs, boundary_factor=boundary_factor, point_idx=cluster.local_points)
    try:
        lists = shard.member_lists(res)
    finally:
        res.free()
    sizes = comm.allgather(np.array([len(lists[0]), len(lists[1])], np.int64))
    cnt = np.stack(sizes)                                                      # world x 2
    return (ShardedCluster(c1, lists[0], cnt[:, 0], cluster.depth + 1),
            ShardedCluster(c2, lists[1], cnt[:, 1], cluster.depth + 1))


def subdivide_clusters(shard, comm: Comm, metric: int, clusters: List[ShardedCluster], desired: int, rng,
                       shard_starts: np.ndarray, boundary_factor: float = 1.1, max_splits: int = 1_000_000):
    """Row-sharded subdivide_clusters (hierarchical.rs:74-105): the reference's serial work-list,
    driven identically on every rank by the global cluster sizes."""
    i, splits = 0, 0
    while i < len(clusters):
        if clusters[i].size > desired:
            splits += 1
            if splits > max_splits:
                raise RuntimeError("subdivide_clusters: split limit reached (the reference would loop forever here)")
            s1, s2 = create_subclusters(shard, comm, metric, clusters[i], rng, shard_starts, boundary_factor)
            clusters[i] = s1                                                   # :95
            clusters.append(s2)                                                # :98
        else:
            i += 1
    return clusters


def clusters_from_assignment(shard, comm: Comm, res, centroid_rows) -> List[ShardedCluster]:
    """The sharded form of assign_points' result (hierarchical.rs:368-390): per cluster the local
    member slice and every rank's member count."""
    lists = shard.member_lists(res)
    counts = np.stack(comm.allgather(np.array([len(x) for x in lists], np.int64)))   # world x k
    return [ShardedCluster(int(centroid_rows[c]), lists[c], counts[:, c], 0) for c in range(len(lists))]


class DeviceShardedKMeans:
    """Flat k-means iterations over row shards, device resident (spf_kmeans): assign + the two
    all-gathers of update_centroids run on the library's stream through the library's own NCCL
    communicator; no host staging between iterations.  Same results as `ShardedKMeans` below
    (which stages every exchange through numpy and also runs on the CPU oracle shards)."""

    def __init__(self, dataset, device_comm, metric: int, row0: int, boundary_factor: float = 1.1, seeded: bool = True):
        self.ds, self.comm, self.metric, self.row0 = dataset, device_comm, metric, int(row0)
        self.factor, self.seeded = boundary_factor, seeded
        self.session = None

    def init(self, global_rows, vectors):
        from .device import KMeansSession
        rows = np.asarray(global_rows, np.uint64)
        if self.session is not None:
            self.session.free()
        self.session = KMeansSession(self.ds, self.comm, self.metric, self.row0, rows.size, self.factor, self.seeded)
        self.session.set_centroids(rows, vectors)

    def step(self):
        self.session.step()

    def centroids(self):
        """(global rows, vectors, global cluster sizes) after the last step."""
        rows, vec, _, cnt = self.session.fetch()
        return rows, vec, cnt

    @property
    def last(self):
        return self.session.assignment()

    def free(self):
        if self.session is not None:
            self.session.free()
            self.session = None


class ShardedKMeans:
    """Flat k-means iterations over row shards (the part of HierarchicalClustering.fit that is
    data-parallel: assign_points + update_centroids), every exchange staged through the host —
    the portable form (gloo / CPU oracle shards) and the checker of `DeviceShardedKMeans`."""

    def __init__(self, shard, comm: Comm, metric: int, boundary_factor: float = 1.1):
        self.shard, self.comm, self.metric, self.factor = shard, comm, metric, boundary_factor
        self.starts = shard_layout(shard, comm)
        self.rows: Optional[np.ndarray] = None       # global centroid rows
        self.vectors: Optional[np.ndarray] = None    # k x d centroid vectors
        self.last = None                             # the rank's last assignment (device resident)

    def init_rows(self, global_rows):
        self.rows = np.asarray(global_rows, np.uint64)
        self.vectors = gather_rows(self.shard, self.comm, self.rows, self.starts)

    def step(self):
        """assign + update; returns the cluster sizes over all shards."""
        if self.last is not None and hasattr(self.last, "free"):
            self.last.free()
        self.last = self.shard.assign_vectors(self.metric, self.vectors, boundary_factor=self.factor)
        self.rows, self.vectors, _ = update_centroids(self.shard, self.comm, self.metric, self.last,
                                                      self.vectors, self.rows, self.starts)
        return self.rows
