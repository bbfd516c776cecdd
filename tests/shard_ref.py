"""Oracle-backed shard for the sharded-build tests: the same interface as
spfresh_b200.sharded.DeviceShard, every value computed by the CPU oracle (test infrastructure)."""
import numpy as np

import oracle


class OracleAssign:
    def __init__(self, res, k):
        self.offsets, self.members, self.best, self.dmin, self.k = res.offsets, res.members, res.best, res.dmin, k

    def free(self):
        pass


class OracleShard:
    def __init__(self, rows, row0=0):
        self.x = np.ascontiguousarray(rows, np.float32)
        self.n, self.d = self.x.shape
        self.row0 = row0

    def assign_vectors(self, metric, centroids, boundary_factor=1.1):
        cv = np.ascontiguousarray(centroids, np.float32)
        aug = np.concatenate([self.x, cv], axis=0)            # centroids appended as extra rows
        cent = np.arange(self.n, self.n + cv.shape[0], dtype=np.uint64)
        res = oracle.assign(aug, metric, cent, point_idx=np.arange(self.n, dtype=np.uint64),
                            boundary_factor=boundary_factor)
        return OracleAssign(res, cv.shape[0])

    def cluster_sums(self, res):
        sums = np.zeros((res.k, self.d), np.float32)
        counts = np.zeros(res.k, np.uint64)
        for c in range(res.k):
            mem = res.members[int(res.offsets[c]):int(res.offsets[c + 1])].astype(np.int64)
            counts[c] = mem.size
            acc = np.zeros(self.d, np.float32)
            for r in mem:                                      # row by row, member order (utils.rs:13)
                acc = (acc + self.x[r]).astype(np.float32)
            sums[c] = acc
        return sums, counts

    def medoid_candidates(self, metric, res, means):
        dist = np.full(res.k, np.inf, np.float32)
        row = np.full(res.k, np.iinfo(np.uint64).max, np.uint64)
        for c in range(res.k):
            mem = res.members[int(res.offsets[c]):int(res.offsets[c + 1])].astype(np.int64)
            for r in mem:                                      # strict <, leftmost wins (:155-171)
                dv = np.float32(oracle.distance(metric, self.x[r], means[c]))
                if dv < dist[c]:
                    dist[c], row[c] = dv, r
        return dist, row

    def rows(self, local_rows):
        return self.x[np.asarray(local_rows, np.int64)]
