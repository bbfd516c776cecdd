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

    def assign_vectors(self, metric, centroids, boundary_factor=1.1, point_idx=None):
        cv = np.ascontiguousarray(centroids, np.float32)
        aug = np.concatenate([self.x, cv], axis=0)            # centroids appended as extra rows
        cent = np.arange(self.n, self.n + cv.shape[0], dtype=np.uint64)
        pts = np.arange(self.n, dtype=np.uint64) if point_idx is None else np.asarray(point_idx, np.uint64)
        res = oracle.assign(aug, metric, cent, point_idx=pts, boundary_factor=boundary_factor)
        return OracleAssign(res, cv.shape[0])

    def member_lists(self, res):
        return [res.members[int(res.offsets[c]):int(res.offsets[c + 1])].copy() for c in range(res.k)]

    def farthest_from(self, metric, c1_vector, members, skip_row=None):
        best_d, best_row = np.float32(0.0), None              # strict >, identity (0, 0.0) (:112-126)
        v = np.asarray(c1_vector, np.float32)
        for r in np.asarray(members, np.int64):
            if skip_row is not None and int(r) == int(skip_row):
                continue
            dv = np.float32(oracle.distance(metric, v, self.x[r]))
            if dv > best_d:
                best_d, best_row = dv, int(r)
        return float(best_d), best_row

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


class OracleKmpp:
    """The shard-local half of the sharded k-means++ (same interface as KmppShardSession)."""

    def __init__(self, shard, metric):
        self.x, self.metric = shard.x, metric
        self.mind = None
        self.w = None

    def fold_vector(self, centroid):
        c = np.ascontiguousarray(centroid, np.float32)
        d = np.array([oracle.distance(self.metric, self.x[i], c) for i in range(self.x.shape[0])], np.float32)
        self.mind = d if self.mind is None else np.where(d < self.mind, d, self.mind)
        acc = np.float32(0.0)
        for v in self.mind:                                   # sequential f32 fold (:278)
            acc = np.float32(acc + v)
        return float(acc)

    def weight_total(self, global_sum):
        denom = np.float32(max(np.float32(global_sum), np.float32(1e-10)))
        self.w = ((self.mind * self.mind).astype(np.float32) / denom).astype(np.float32).astype(np.float64)
        ok = bool(np.all(self.w >= 0.0))
        # the device adds 1024-weight blocks (256 threads x 4 consecutive weights, tree over threads),
        # then the blocks in order: mirror it so the f64 totals agree bit for bit
        tot = np.float64(0.0)
        self.block_sums = []
        for b0 in range(0, self.w.size, 1024):
            blk = np.zeros(1024, np.float64)
            seg = self.w[b0:b0 + 1024]
            blk[:seg.size] = seg
            t = blk.reshape(256, 4)
            per_thread = np.zeros(256, np.float64)
            for e in range(4):
                per_thread = per_thread + t[:, e]
            sarr = per_thread.copy()
            step = 128
            while step > 0:
                sarr[:step] = sarr[:step] + sarr[step:2 * step]
                step //= 2
            self.block_sums.append(np.float64(sarr[0]))
            tot = np.float64(tot + sarr[0])
        return float(tot), ok

    def pick_local(self, target):
        u = np.float64(target)
        cum = np.float64(0.0)
        b = 0
        nb = len(self.block_sums)
        while b + 1 < nb:
            if not (cum + self.block_sums[b] <= u):
                break
            cum = np.float64(cum + self.block_sums[b])
            b += 1
        n = self.w.size
        for i in range(b * 1024, n - 1):
            cum = np.float64(cum + self.w[i])
            if not (cum <= u):
                return i
        return n - 1

    def free(self):
        pass


def _kmpp_session(self, metric):
    return OracleKmpp(self, metric)


OracleShard.kmpp_session = _kmpp_session
