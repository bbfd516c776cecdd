"""Slow, independent pure-Python/numpy-scalar restatement of the reference semantics, used
only to cross-check the C oracle on tiny inputs (a second opinion written separately from
oracle/spf_oracle.c).  Cites the same reference lines."""
import numpy as np

F = np.float32


def dist(metric, a, b):
    # distance.rs:16-43 — sequential accumulation in f32
    acc = F(0)
    for x, y in zip(a, b):
        df = F(F(x) - F(y))
        if metric == 0:
            acc = F(acc + F(df * df))
        elif metric == 1:
            acc = F(acc + F(abs(df)))
        else:
            ad = F(abs(df))
            if ad > acc:
                acc = ad
    return acc


def mean(data, idx):
    # utils.rs:5-15
    d = data.shape[1]
    res = np.zeros(d, F)
    if len(idx) == 0:
        return res
    for r in idx:
        for j in range(d):
            res[j] = F(res[j] + F(data[r, j]))
    m = F(len(idx))
    return np.array([F(v / m) for v in res], F)


def assign(data, metric, point_idx, centroid_rows, factor=1.1):
    # hierarchical.rs:295-364
    k = len(centroid_rows)
    lists = [[] for _ in range(k)]
    best, dmin = [], []
    for p in point_idx:
        ds = [dist(metric, data[p], data[c]) for c in centroid_rows]
        bj, bd = 0, F(np.inf)
        for j, dj in enumerate(ds):
            if dj < bd:
                bj, bd = j, dj
        thr = F(bd * F(factor))
        take = [False] * k
        take[bj] = True
        for j, dj in enumerate(ds):
            if j != bj and dj < thr:
                cc = dist(metric, data[centroid_rows[bj]], data[centroid_rows[j]])
                if cc >= dj:
                    take[j] = True
        for j in range(k):
            if take[j]:
                lists[j].append(int(p))
        best.append(bj)
        dmin.append(bd)
    return lists, best, dmin


def update_medoids(data, metric, lists, old_rows):
    # hierarchical.rs:138-181
    out = []
    for pts, old in zip(lists, old_rows):
        if len(pts) == 0:
            out.append(int(old))
            continue
        mu = mean(data, pts)
        bi, bd = 0, F(np.inf)
        for p in pts:
            dd = dist(metric, data[p], mu)
            if dd < bd:
                bi, bd = int(p), dd
        out.append(bi)
    return out


def farthest(data, metric, c1, pts):
    # hierarchical.rs:112-126
    mi, md = 0, F(0)
    for p in pts:
        if p == c1:
            continue
        dd = dist(metric, data[c1], data[p])
        if dd > md:
            mi, md = int(p), dd
    return mi


def fit(data, metric, init_rows, desired, pick, max_splits=1000):
    # hierarchical.rs:65-135
    n = data.shape[0]
    lists, _, _ = assign(data, metric, range(n), init_rows)
    rows = update_medoids(data, metric, lists, init_rows)
    clusters = [[rows[j], lists[j], 0] for j in range(len(init_rows))]
    i, splits = 0, 0
    while i < len(clusters):
        if len(clusters[i][1]) > desired:
            splits += 1
            assert splits <= max_splits
            pts, depth = clusters[i][1], clusters[i][2] + 1
            c1 = pts[pick(len(pts))]
            c2 = farthest(data, metric, c1, pts)
            sub, _, _ = assign(data, metric, pts, [c1, c2])
            clusters[i] = [c1, sub[0], depth]
            clusters.append([c2, sub[1], depth])
        else:
            i += 1
    return clusters


def search(data, lists, centroid_rows, q, k, nprobe=0, prune=1.2):
    # spann_index.rs:148-197
    if nprobe == 0:
        nprobe = k
    cd = sorted(((dist(0, q, data[c]), j) for j, c in enumerate(centroid_rows)))
    cd = cd[:nprobe]
    thr = F(F(prune) * F(cd[0][0] + np.finfo(F).eps))
    cands = []
    for _, j in cd:
        for p in lists[j]:
            dd = dist(0, q, data[p])
            if dd <= thr:
                cands.append((dd, int(p)))
    if not cands:
        return []
    order = sorted(range(len(cands)), key=lambda t: (cands[t][0], t))
    return [cands[t] for t in order[:k]]
