"""Host-side mirror of src/clustering/hierarchical.rs: same names, argument meaning and control
flow; every distance / mean / argmin runs on the B200 through the C ABI.

The reference draws from rand::SmallRng (hierarchical.rs:184-189).  Here the random decisions come
from a `RandomSource` so tests (and the parity oracle) can script them; `NumpyRandomSource` is the
default.  The draws are requested in exactly the reference's order.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from enum import Enum
from typing import List, Optional

import numpy as np

from . import _capi as capi
from .device import Context, Dataset

BOUNDARY_THRESHOLD = 1.1   # hierarchical.rs:55
KMPP_BATCH = 256           # k-means++ rounds per device batch (spf_kmpp_rounds): one host sync per batch


class DistanceMetric:
    """distance.rs:7-10.  `kind` is the routing hint SURVEY.md §8(b) adds to the trait."""
    kind: int = -1
    name = ""

    def compute(self, point1, point2, ctx: Optional[Context] = None):
        p1, p2 = np.asarray(point1), np.asarray(point2)
        if p1.shape != p2.shape or p1.size == 0:
            # ndarray-stats returns Err (ShapeMismatch / EmptyInput) and the reference unwraps it
            raise ValueError("called `Result::unwrap()` on an `Err` value: shape mismatch or empty input")
        return (ctx or Context.default()).distance_pairs(self.kind, p1[None, :], p2[None, :])[0]


class SquaredEuclideanDistance(DistanceMetric):   # distance.rs:14-21
    kind, name = capi.METRIC_EUCLIDEAN, "Euclidean"


class ManhattanDistance(DistanceMetric):          # distance.rs:25-32
    kind, name = capi.METRIC_MANHATTAN, "Manhattan"


class ChebyshevDistance(DistanceMetric):          # distance.rs:36-43
    kind, name = capi.METRIC_CHEBYSHEV, "Chebyshev"


class InitializationMethod(Enum):                 # hierarchical.rs:13-16
    Random = "Random"
    KMeansPlusPlus = "KMeansPlusPlus"


class RandomSource:
    """The random decisions the reference takes from its RNG, in its order."""

    def choose_multiple(self, n: int, k: int) -> List[int]:   # (0..n).choose_multiple(rng, k) :204
        raise NotImplementedError

    def choose_index(self, n: int) -> int:                    # (0..n).choose / slice.choose :111,253
        raise NotImplementedError

    def uniform01(self) -> float:                             # draw behind choose_weighted :285
        raise NotImplementedError


class NumpyRandomSource(RandomSource):
    def __init__(self, seed: Optional[int] = None):
        self.rng = np.random.default_rng(seed)

    def choose_multiple(self, n, k):
        return self.rng.choice(n, size=min(k, n), replace=False).tolist()

    def choose_index(self, n):
        return int(self.rng.integers(0, n))

    def uniform01(self):
        return float(self.rng.random())


class ScriptedRandomSource(RandomSource):
    """Replays explicit decisions (tests / parity runs)."""

    def __init__(self, multiple=None, index=None, u01=None):
        self.multiple = list(multiple) if multiple is not None else None
        self.index = index        # callable(n) -> int, or list consumed in order
        self.u01 = list(u01) if u01 is not None else []
        self._iu = 0
        self._ii = 0

    def choose_multiple(self, n, k):
        assert self.multiple is not None and len(self.multiple) == min(k, n)
        return list(self.multiple)

    def choose_index(self, n):
        if callable(self.index):
            return int(self.index(n))
        v = self.index[self._ii]
        self._ii += 1
        return int(v)

    def uniform01(self):
        v = self.u01[self._iu]
        self._iu += 1
        return float(v)


@dataclass
class ClusteringParams:                           # hierarchical.rs:18-24
    distance_metric: DistanceMetric
    initialization_method: InitializationMethod
    desired_cluster_size: Optional[int]
    initial_k: int
    rng_seed: Optional[int] = None
    random_source: Optional[RandomSource] = None   # overrides rng_seed when given


@dataclass
class Cluster:                                    # hierarchical.rs:26-41
    centroid_idx: Optional[int]
    points: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint64))
    depth: int = 0


class HierarchicalClustering:
    """hierarchical.rs:43-391 with the batched seams routed to the GPU."""

    def __init__(self, params: ClusteringParams, data, ctx: Optional[Context] = None,
                 dataset: Optional[Dataset] = None, max_splits: int = 1_000_000):
        self.params = params
        self.data = np.asarray(data) if data is not None else None
        self.ctx = ctx or (dataset.ctx if dataset is not None else Context.default())
        self.dataset = dataset if dataset is not None else Dataset(self.ctx, self.data)
        self.clusters: List[Cluster] = []
        self.max_splits = max_splits   # the reference loops forever on duplicate-heavy clusters
        self._last_assign = None       # device-resident result of the last full assign_points()

    # -- helpers -------------------------------------------------------------------------------
    @property
    def _metric(self) -> int:
        return self.params.distance_metric.kind

    def nrows(self) -> int:
        return self.dataset.n

    def get_rng(self) -> RandomSource:            # hierarchical.rs:184-189
        if self.params.random_source is not None:
            return self.params.random_source
        # a fresh generator per call, like SmallRng::seed_from_u64(seed) in every caller
        return NumpyRandomSource(self.params.rng_seed)

    # -- fit -------------------------------------------------------------------------------------
    def fit(self):                                # hierarchical.rs:65-71
        self.initialize_clusters(self.params.initial_k)
        self.assign_points()
        self.update_centroids()
        self.subdivide_clusters()

    def initialize_clusters(self, k: int):        # hierarchical.rs:192-197
        if self.params.initialization_method == InitializationMethod.Random:
            self.initialize_clusters_randomly(k)
        else:
            self.initialize_clusters_kmeans_plus_plus(k)

    def initialize_clusters_randomly(self, k: int):   # hierarchical.rs:200-210 (host side, RNG only)
        idx = self.get_rng().choose_multiple(self.nrows(), k)
        self.clusters = [Cluster(int(i), np.zeros(0, np.uint64), 0) for i in idx]

    def initialize_clusters_kmeans_plus_plus(self, k: int):   # hierarchical.rs:249-293
        rng = self.get_rng()
        n = self.nrows()
        first = rng.choose_index(n)                                   # :253-255
        self.clusters.append(Cluster(int(first), np.zeros(0, np.uint64), 0))
        sess = self.dataset.kmeanspp(self._metric, first)
        draws: List[float] = []                                       # drawn ahead, not yet used
        try:
            left = k - 1
            while left > 0:                                           # :259
                want = min(left, KMPP_BATCH)
                while len(draws) < want:
                    draws.append(rng.uniform01())
                rows, failed = sess.rounds(draws[:want])              # :260-286 on the device, one sync per batch
                for r in rows:
                    self.clusters.append(Cluster(int(r), np.zeros(0, np.uint64), 0))
                used = len(rows) + (1 if failed else 0)               # the failing round consumed its draw too
                del draws[:used]
                left -= len(rows)
                if failed:                                            # :287-290 uniform fallback
                    chosen = rng.choose_index(n)
                    sess.push(chosen)
                    self.clusters.append(Cluster(int(chosen), np.zeros(0, np.uint64), 0))
                    left -= 1
        finally:
            sess.free()

    # -- assign ----------------------------------------------------------------------------------
    def assign_points_to_clusters(self, point_indices, centroids) -> List[np.ndarray]:
        """hierarchical.rs:295-364.  centroids: list of (row_idx, depth)."""
        rows = [c[0] for c in centroids]
        res = self.dataset.assign(self._metric, rows, point_idx=point_indices,
                                  boundary_factor=BOUNDARY_THRESHOLD)
        try:
            return res.fetch(best=False, dmin=False).lists()
        finally:
            res.free()

    def assign_points(self):                      # hierarchical.rs:368-390
        rows = [c.centroid_idx for c in self.clusters]
        if self._last_assign is not None:
            self._last_assign.free()
        res = self.dataset.assign(self._metric, rows, boundary_factor=BOUNDARY_THRESHOLD)
        self._last_assign = res                   # kept on the device for update_centroids
        for c, pts in zip(self.clusters, res.fetch(best=False, dmin=False).lists()):
            c.points = pts

    def update_centroids(self):                   # hierarchical.rs:138-181
        old = np.array([c.centroid_idx for c in self.clusters], np.uint64)
        if self._last_assign is not None and self._last_assign.k == len(self.clusters):
            new = self.dataset.update_medoids_from(self._metric, self._last_assign, old)
            self._last_assign.free()
            self._last_assign = None
        else:
            offsets = np.zeros(len(self.clusters) + 1, np.uint64)
            for i, c in enumerate(self.clusters):
                offsets[i + 1] = offsets[i] + np.uint64(len(c.points))
            members = (np.concatenate([np.asarray(c.points, np.uint64) for c in self.clusters])
                       if self.clusters else np.zeros(0, np.uint64))
            new = self.dataset.update_medoids(self._metric, offsets, members, old)
        for c, r in zip(self.clusters, new):
            c.centroid_idx = int(r)

    # -- bisect ----------------------------------------------------------------------------------
    def subdivide_clusters(self):                 # hierarchical.rs:74-105
        desired = self.params.desired_cluster_size
        i, splits = 0, 0
        while i < len(self.clusters):
            if len(self.clusters[i].points) > desired:
                splits += 1
                if splits > self.max_splits:
                    raise RuntimeError("subdivide_clusters: split limit reached "
                                       "(the reference would loop forever here)")
                pts, depth = self.clusters[i].points, self.clusters[i].depth
                s1, s2 = self.create_subclusters(pts, depth + 1)
                self.clusters[i] = s1             # :95
                self.clusters.append(s2)          # :98
            else:
                i += 1

    def create_subclusters(self, points, new_depth: int):   # hierarchical.rs:107-135
        rng = self.get_rng()
        c1 = int(points[rng.choose_index(len(points))])                       # :111
        c2 = self.dataset.farthest(self._metric, c1, points)                  # :112-126
        lists = self.assign_points_to_clusters(points, [(c1, new_depth), (c2, new_depth)])   # :129
        return Cluster(c1, lists[0], new_depth), Cluster(c2, lists[1], new_depth)

    # -- labels ----------------------------------------------------------------------------------
    def labels(self) -> np.ndarray:               # hierarchical.rs:215-246
        """One label per point: among the clusters a point belongs to, the one whose centroid is
        strictly nearest, scanning clusters in order starting from label 0."""
        n = self.nrows()
        labels = np.zeros(n, np.int64)
        rows = np.array([c.centroid_idx for c in self.clusters], np.uint64)
        ctx, m = self.ctx, self._metric
        for c_idx, c in enumerate(self.clusters):
            pts = np.asarray(c.points, np.int64)
            if pts.size == 0:
                continue
            x = self.data[pts]
            this_d = ctx.distance_pairs(m, x, np.broadcast_to(self.data[int(rows[c_idx])], x.shape))
            old_d = ctx.distance_pairs(m, x, self.data[rows[labels[pts]].astype(np.int64)])
            upd = this_d < old_d
            labels[pts[upd]] = c_idx
        return labels
