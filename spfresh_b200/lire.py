"""Host-side mirror of the two LIRE operations that sit on the hot path's kernels
(src/spann/lire/operations.rs:8-120 `Split`, :222-300 `Reassign`).

The reference computes them with per-pair `DistanceMetric::compute` calls; here the distances run on
the B200 through the C ABI (farthest-point fold, k = 2 assignment without replication, batched
pair distances) and only the bookkeeping stays on the host.  Tie rules follow the reference:
`max_by` keeps the LAST maximum (the farthest-point seed), `dist1 <= dist2` sends ties to the first
partition, `min_by` keeps the FIRST minimum.  (NaN distances, which `partial_cmp(..).unwrap_or(Equal)`
treats as equal to everything, are not reproduced.)
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Set, Tuple

import numpy as np

from . import _capi as capi
from .clustering import DistanceMetric
from .device import Context, Dataset


class LireError(RuntimeError):
    pass


class Split:
    """operations.rs:8-120."""

    def __init__(self, posting_id: int, vectors: Sequence[Tuple[int, Sequence[float]]], distance_metric: DistanceMetric,
                 new_posting_ids: Tuple[int, int], ctx: Optional[Context] = None):
        self.posting_id = int(posting_id)
        self.vectors = [(int(i), np.asarray(v, np.float32)) for i, v in vectors]
        self.distance_metric = distance_metric
        self.new_posting_ids = (int(new_posting_ids[0]), int(new_posting_ids[1]))
        self.ctx = ctx                       # resolved on first device use (validate() needs no GPU)
        self._ds: Optional[Dataset] = None

    def _dataset(self) -> Dataset:
        if self._ds is None:
            if self.ctx is None:
                self.ctx = Context.default()
            self._ds = Dataset(self.ctx, np.stack([v for _, v in self.vectors]))
        return self._ds

    def select_initial_centroids(self) -> Tuple[np.ndarray, np.ndarray]:   # operations.rs:33-58
        if len(self.vectors) < 2:
            raise LireError("Not enough vectors to split")
        first = self.vectors[0][1]
        m = len(self.vectors)
        # the last maximum of d(first, v) over vectors[1..]: the fold keeps the earliest maximum, so it
        # is run over the members in reverse order
        members = np.arange(m - 1, 0, -1, dtype=np.uint64)
        dist, row = self._dataset().farthest_from(self.distance_metric.kind, first, members)
        if row is None:                     # every distance is 0: max_by returns the last element
            row = m - 1
        return first, self.vectors[int(row)][1]

    def assign_vectors(self, centroid1, centroid2):                        # operations.rs:61-82
        """(partition1, partition2): `dist1 <= dist2` → partition1."""
        res = self._dataset().assign_vectors(self.distance_metric.kind, np.stack([centroid1, centroid2]),
                                             flags=capi.ASSIGN_NO_CSR)
        try:
            best = res.fetch(csr=False).best
        finally:
            res.free()
        p1 = [(i, v) for (i, v), b in zip(self.vectors, best) if b == 0]
        p2 = [(i, v) for (i, v), b in zip(self.vectors, best) if b != 0]
        return p1, p2

    def execute(self) -> Set[int]:                                         # operations.rs:86-101
        c1, c2 = self.select_initial_centroids()
        self.partitions = self.assign_vectors(c1, c2)
        return self.get_affected_partitions()

    def validate(self) -> bool:                                            # operations.rs:103-112
        return (len(self.vectors) >= 2 and self.new_posting_ids[0] != self.posting_id
                and self.new_posting_ids[1] != self.posting_id and self.new_posting_ids[0] != self.new_posting_ids[1])

    def get_affected_partitions(self) -> Set[int]:                         # operations.rs:114-120
        return {self.posting_id, self.new_posting_ids[0], self.new_posting_ids[1]}

    def free(self):
        if self._ds is not None:
            self._ds.free()
            self._ds = None


class Reassign:
    """operations.rs:222-300."""

    def __init__(self, vector_id: int, vector: Sequence[float], from_posting: int,
                 candidate_postings: Sequence[Tuple[int, Sequence[float]]], distance_metric: DistanceMetric, version: int,
                 ctx: Optional[Context] = None):
        self.vector_id, self.from_posting, self.version = int(vector_id), int(from_posting), int(version)
        self.vector = np.asarray(vector, np.float32)
        self.candidate_postings = [(int(p), np.asarray(c, np.float32)) for p, c in candidate_postings]
        self.distance_metric = distance_metric
        self.ctx = ctx

    def find_best_posting(self) -> int:                                    # operations.rs:253-276
        if not self.candidate_postings:
            raise LireError("No candidate postings available")
        if self.ctx is None:
            self.ctx = Context.default()
        cen = np.stack([c for _, c in self.candidate_postings])
        d = self.ctx.distance_pairs(self.distance_metric.kind, np.broadcast_to(self.vector, cen.shape), cen)
        return self.candidate_postings[int(np.argmin(d))][0]               # min_by: the first minimum

    def get_affected_partitions(self) -> Set[int]:
        return {self.from_posting, self.find_best_posting()}


def reassign_batch(ctx: Context, metric: DistanceMetric, vectors, candidate_centroids) -> np.ndarray:
    """find_best_posting for many vectors against one shared candidate set: index of the nearest
    candidate per vector (first minimum), as one k = |candidates| assignment without replication."""
    ds = Dataset(ctx, np.asarray(vectors, np.float32))
    try:
        res = ds.assign_vectors(metric.kind, np.asarray(candidate_centroids, np.float32), flags=capi.ASSIGN_NO_CSR)
        try:
            return res.fetch(csr=False).best.astype(np.int64)
        finally:
            res.free()
    finally:
        ds.free()
