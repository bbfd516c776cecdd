"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on
the same seeded inputs.  Integer / index outputs and all distances must be BIT-EXACT (the
north-star tolerance is 1e-5 relative; the design delivers identity, so that is what is tested).
"""
import itertools
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

METRICS = [0, 1, 2]


@pytest.fixture(scope="module")
def spf():
    import spfresh_b200 as s
    from spfresh_b200 import build as b
    b.build()
    if s._capi.lib().spf_device_count() == 0:
        pytest.fail("no CUDA device: the gpu-marked tests need a B200 (there is no CPU fallback)")
    return s


@pytest.fixture(scope="module")
def ctx(spf):
    return spf.Context.default()


def gauss(n, d, seed):
    return np.random.Generator(np.random.Philox(key=seed)).standard_normal((n, d), dtype=np.float32)


def clustered(n, d, ncent, seed):
    g = np.random.Generator(np.random.Philox(key=seed))
    cen = 2.0 * g.standard_normal((ncent, d), dtype=np.float32)
    lab = g.integers(0, ncent, n)
    return (cen[lab] + 0.5 * g.standard_normal((n, d), dtype=np.float32)).astype(np.float32)


def check_assign(got, ref):
    assert np.array_equal(got.offsets, ref.offsets)
    assert np.array_equal(got.members, ref.members)
    assert np.array_equal(got.best, ref.best)
    assert np.array_equal(got.dmin.view(np.uint32), ref.dmin.view(np.uint32))


# ----------------------------------------------------------------------------------------------
# distances (distance.rs KATs) through the device
# ----------------------------------------------------------------------------------------------
def test_distance_kats_on_device(spf, ctx, kats, oracle):
    name = {"Euclidean": 0, "Manhattan": 1, "Chebyshev": 2}
    for kat in kats["distance"]:
        got = ctx.distance_pairs(name[kat["metric"]], [kat["a"]], [kat["b"]])[0]
        assert abs(float(got) - kat["expected"]) < kat["tol"], kat["cite"]
    rng = np.random.default_rng(1)
    for d in (1, 3, 31, 32, 33, 100, 128, 960):
        a = rng.standard_normal((257, d)).astype(np.float32)
        b = rng.standard_normal((257, d)).astype(np.float32)
        for m in METRICS:
            got = ctx.distance_pairs(m, a, b)
            ref = np.array([oracle.distance(m, a[i], b[i]) for i in range(257)], np.float32)
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (m, d)
    with pytest.raises(ValueError):           # the reference panics on a shape mismatch
        spf.SquaredEuclideanDistance().compute([1.0, 2.0], [1.0, 2.0, 3.0])


# ----------------------------------------------------------------------------------------------
# assignment — exact CUDA-core kernel (all metrics)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("n,d,k", [(6, 2, 2), (500, 7, 9), (3000, 33, 70), (4097, 128, 300)])
def test_assign_exact_matches_oracle(spf, ctx, oracle, metric, n, d, k):
    data = gauss(n, d, 10 + n)
    data[n // 2] = data[1]                       # duplicate rows → exact ties
    rng = np.random.default_rng(n + k)
    cent = rng.choice(n, k, replace=False)
    if k > 5:
        cent[3], cent[4] = 1, n // 2             # identical centroids: lowest slot must win
    ds = spf.Dataset(ctx, data)
    got = ds.assign(metric, cent, flags=spf.ASSIGN_FORCE_EXACT).fetch()
    check_assign(got, oracle.assign(data, metric, cent))


def test_assign_toy_kat_all_inits(spf, ctx, kats):
    """hierarchical.rs:466-486 on the device, for every ordered pair of distinct centroids."""
    data = np.array(kats["toy_data"]["rows"], np.float32)
    ds = spf.Dataset(ctx, data)
    for c in itertools.permutations(range(6), 2):
        r = ds.assign(0, c).fetch()
        sizes = np.diff(r.offsets.astype(np.int64))
        assert sizes.sum() == 6 and (sizes > 0).all()


@pytest.mark.parametrize("metric", METRICS)
def test_assign_subset_and_order(spf, ctx, oracle, metric):
    data = clustered(5000, 24, 40, 3)
    rng = np.random.default_rng(5)
    sub = rng.permutation(5000)[:1777]           # arbitrary order must be preserved in the lists
    cent = rng.choice(5000, 2, replace=False)    # the bisect shape (k = 2)
    ds = spf.Dataset(ctx, data)
    got = ds.assign(metric, cent, point_idx=sub).fetch()
    check_assign(got, oracle.assign(data, metric, cent, point_idx=sub))


def test_assign_candidate_overflow_path(spf, oracle):
    """More candidates than slots → the brute-force rows must give the same answer."""
    c2 = spf.Context(0)
    try:
        c2.set_param("cand_cap", 4)
        data = gauss(2000, 16, 77)
        cent = np.random.default_rng(2).choice(2000, 200, replace=False)
        ds = spf.Dataset(c2, data)
        for flags in (spf.ASSIGN_FORCE_EXACT, spf.ASSIGN_DEFAULT):
            got = ds.assign(0, cent, boundary_factor=1.6, flags=flags).fetch()
            check_assign(got, oracle.assign(data, 0, cent, boundary_factor=1.6))
        ds.free()
    finally:
        c2.close()


def test_assign_degenerate_inputs(spf, ctx, oracle):
    # all points identical: every distance 0, thr 0, best = slot 0, nothing replicated
    same = np.ones((300, 8), np.float32)
    ds = spf.Dataset(ctx, same)
    check_assign(ds.assign(0, [5, 9, 200]).fetch(), oracle.assign(same, 0, [5, 9, 200]))
    # k == 1
    data = gauss(100, 5, 1)
    ds = spf.Dataset(ctx, data)
    check_assign(ds.assign(1, [7]).fetch(), oracle.assign(data, 1, [7]))
    # argument errors surface as status codes, not crashes
    with pytest.raises(spf.SpfError):
        ds.assign(0, [100])                      # centroid row out of range
    with pytest.raises(spf.SpfError):
        ds.assign(0, [])                         # k == 0
    with pytest.raises(spf.SpfError):
        ds.assign(0, [1], point_idx=[5, 1000])   # point row out of range


# ----------------------------------------------------------------------------------------------
# assignment — tcgen05 TF32 candidate GEMM + exact resolve (Euclidean)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d,k,kind", [
    (4096, 128, 256, "gauss"), (5000, 128, 300, "gauss"), (20000, 96, 1000, "clustered"),
    (3000, 64, 513, "gauss"), (2500, 100, 64, "clustered"), (30000, 128, 4096, "clustered"),
    (4000, 200, 300, "clustered"), (3000, 960, 256, "clustered"), (2048, 516, 128, "gauss"),   # streamed point K blocks
])
def test_assign_tensor_path_matches_oracle(spf, ctx, oracle, n, d, k, kind):
    data = gauss(n, d, n + d) if kind == "gauss" else clustered(n, d, max(k // 4, 2), n + d)
    data[17] = data[3]
    cent = np.random.default_rng(k).choice(n, k, replace=False)
    cent[0], cent[1] = 3, 17                     # identical centroids
    ds = spf.Dataset(ctx, data)
    ctx.set_profiling(True)
    got = ds.assign(0, cent).fetch()
    assert ctx.kernel_ms("assign_tc") > 0, "the tcgen05 path did not run"
    ctx.set_profiling(False)
    check_assign(got, oracle.assign(data, 0, cent))


def test_assign_tensor_path_large_norms(spf, ctx, oracle):
    """SIFT-like data: all-positive, norms >> distances — the worst case for the TF32 slack."""
    g = np.random.Generator(np.random.Philox(key=9))
    data = np.floor(g.random((8000, 128), dtype=np.float32) * 128).astype(np.float32)
    cent = np.random.default_rng(1).choice(8000, 512, replace=False)
    ds = spf.Dataset(ctx, data)
    check_assign(ds.assign(0, cent).fetch(), oracle.assign(data, 0, cent))


def test_assign_tensor_equals_exact_at_scale(spf, ctx):
    """Size-independent property at a size the oracle is too slow for: the tensor path and the
    exact CUDA-core path (itself oracle-checked above) agree bit for bit."""
    data = gauss(200_000, 128, 42)
    cent = np.random.Generator(np.random.Philox(key=7)).choice(200_000, 2048, replace=False)
    ds = spf.Dataset(ctx, data)
    a = ds.assign(0, cent).fetch()
    b = ds.assign(0, cent, flags=spf.ASSIGN_FORCE_EXACT).fetch()
    check_assign(a, b)
    sizes = np.diff(a.offsets.astype(np.int64))
    assert sizes.sum() == a.members.size and sizes.sum() >= 200_000
    for j in (0, 1, 1000, 2047):                 # members stay in input order inside a cluster
        seg = a.members[int(a.offsets[j]):int(a.offsets[j + 1])]
        assert np.all(np.diff(seg.astype(np.int64)) > 0)
    nearest = np.bincount(a.best, minlength=2048)
    assert nearest.sum() == 200_000


@pytest.mark.parametrize("param,value", [("short_cap", 2), ("short_cap", 8), ("work_cap", 1), ("work_cap", 100),
                                         ("cand_cap", 16)])
def test_assign_small_buffers_fall_back_exactly(spf, oracle, param, value):
    """Short list / work list / record buffers too small for the data: the dense fallback and the
    inline recomputation in finalize must reproduce the oracle bit for bit (tensor path)."""
    c2 = spf.Context(0)
    try:
        c2.set_param(param, value)
        data = gauss(6000, 64, 91)
        cent = np.random.default_rng(6).choice(6000, 192, replace=False)
        ds = spf.Dataset(c2, data)
        c2.set_profiling(True)
        got = ds.assign(0, cent, boundary_factor=1.15).fetch()
        assert c2.kernel_ms("assign_tc") > 0, "the tcgen05 path did not run"
        check_assign(got, oracle.assign(data, 0, cent, boundary_factor=1.15))
        if param != "work_cap":
            assert c2.last_overflow_rows() > 0          # the fallback was exercised
        ds.free()
    finally:
        c2.close()


def test_assign_vectors_and_fetch_rows(spf, ctx, oracle):
    """spf_assign_vectors with vectors that are not dataset rows (the sharded case) == oracle on the
    augmented matrix; spf_dataset_fetch_rows returns the uploaded rows bit for bit."""
    data = clustered(5000, 96, 24, 33)
    g = np.random.default_rng(9)
    cvec = (data[g.choice(5000, 80, replace=False)] + 0.05 * g.standard_normal((80, 96))).astype(np.float32)
    ds = spf.Dataset(ctx, data)
    aug = np.concatenate([data, cvec])
    ref = oracle.assign(aug, 0, np.arange(5000, 5080, dtype=np.uint64), point_idx=np.arange(5000, dtype=np.uint64))
    check_assign(ds.assign_vectors(0, cvec).fetch(), ref)
    pick = g.choice(5000, 300)
    assert np.array_equal(ds.fetch_rows(pick).view(np.uint32), data[pick].view(np.uint32))
    with pytest.raises(spf.SpfError):
        ds.fetch_rows([5000])


def test_assign_is_chunk_invariant(spf, oracle):
    """The point list is resolved in chunks; any chunk size must give the oracle's answer (tensor
    and exact path, full list and subset)."""
    c2 = spf.Context(0)
    try:
        data = clustered(9000, 64, 40, 5)
        cent = np.random.default_rng(3).choice(9000, 130, replace=False)
        ref = oracle.assign(data, 0, cent)
        sub = np.random.default_rng(4).permutation(9000)[:5000]
        ref_sub = oracle.assign(data, 0, cent, point_idx=sub)
        ds = spf.Dataset(c2, data)
        for chunk in (1000, 4096, 8999):
            c2.set_param("chunk_rows", chunk)
            for flags in (spf.ASSIGN_DEFAULT, spf.ASSIGN_FORCE_EXACT):
                check_assign(ds.assign(0, cent, flags=flags).fetch(), ref)
            check_assign(ds.assign(0, cent, point_idx=sub).fetch(), ref_sub)
        ds.free()
    finally:
        c2.close()


@pytest.mark.parametrize("metric", METRICS)
def test_assign_host_streamed_matches_oracle(spf, oracle, metric):
    """spf_assign_host: chunked upload overlapped with compute, rows with a stride, odd d (padding)."""
    c2 = spf.Context(0)
    try:
        c2.set_param("chunk_rows", 3000)
        wide = clustered(10000, 72, 30, 11 + metric)
        data = wide[:, :67]                          # row stride 72 > d = 67, ld = 68
        cent = np.random.default_rng(5).choice(10000, 96, replace=False)
        ds, res = spf.Dataset.assign_from_host(c2, data, metric, cent)
        ref = oracle.assign(np.ascontiguousarray(data), metric, cent)
        check_assign(res.fetch(), ref)
        # the returned dataset is fully resident and usable by the other entry points
        check_assign(ds.assign(metric, cent).fetch(), ref)
        rows = ds.update_medoids_from(metric, res, cent)
        assert np.array_equal(rows, oracle.update_medoids(np.ascontiguousarray(data), metric, ref.offsets, ref.members, cent))
        _, res2 = spf.Dataset.assign_from_host(c2, data, metric, cent, keep_dataset=False)
        check_assign(res2.fetch(), ref)
        with pytest.raises(spf.SpfError):
            spf.Dataset.assign_from_host(c2, data, metric, [10000])
        for obj in (res, res2, ds):              # handles must go before their context
            obj.free()
    finally:
        c2.close()


def test_pageable_host_buffers_go_through_the_staging_ring(spf):
    """Ordinary heap memory (what the reference's ndarray callers hold) is staged by worker threads
    through the context's pinned ring (more blocks than slots: the ring wraps), in both directions;
    results equal the driver-staged copies bit for bit."""
    rng = np.random.default_rng(21)
    wide = rng.standard_normal((260_000, 68), dtype=np.float32)
    data = wide[:, :66]                                # stride 68 > d = 66, ld = 68: rows re-pitched by the workers
    cent = rng.choice(data.shape[0], 64, replace=False)
    got = {}
    for staging in (1, 0):
        c2 = spf.Context(0)
        try:
            c2.set_param("no_host_staging", 0 if staging else 1)
            ds0 = spf.Dataset(c2, data)                # spf_dataset_upload, 68 MB
            pick = rng.choice(data.shape[0], 5000, replace=False)
            assert np.array_equal(ds0.fetch_rows(pick), data[pick])
            ds0.free()
            ds, res = spf.Dataset.assign_from_host(c2, data, 0, cent, boundary_factor=1.3)
            f = res.fetch()                            # members: tens of MB into pageable numpy arrays
            assert f.members.nbytes > (8 << 20)
            got[staging] = f
            res.free()
            ds.free()
        finally:
            c2.close()
    a, b = got[1], got[0]
    assert np.array_equal(a.best, b.best) and np.array_equal(a.dmin.view(np.uint32), b.dmin.view(np.uint32))
    assert np.array_equal(a.offsets, b.offsets) and np.array_equal(a.members, b.members)


@pytest.mark.parametrize("metric", METRICS)
def test_sharded_build_step_matches_oracle_shards(spf, oracle, metric):
    """§8(e) build: three row shards on one GPU exchanging partial sums / medoid candidates
    in-process; every rank's device ops (spf_assign_vectors, spf_cluster_sums,
    spf_medoid_candidates) against the oracle-backed shard, bit for bit."""
    import sys
    import threading
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from shard_ref import OracleShard
    from spfresh_b200.sharded import DeviceShard, ShardedKMeans, ThreadComm
    data = clustered(2600, 20, 16, 77 + metric)
    init = np.random.default_rng(8).choice(2600, 16, replace=False)
    bounds = [0, 700, 1500, 2600]

    def run_all(make_shard):
        grp = ThreadComm.Group(3)
        out = [None] * 3

        def run(r):
            km = ShardedKMeans(make_shard(r), ThreadComm(grp, r), metric)
            km.init_rows(init)
            rows = [np.array(km.step(), copy=True) for _ in range(2)]
            out[r] = (rows, np.array(km.vectors, copy=True))
        th = [threading.Thread(target=run, args=(r,)) for r in range(3)]
        [t.start() for t in th]
        [t.join() for t in th]
        return out
    ctxs = [spf.Context(0) for _ in range(3)]
    try:
        dss = [spf.Dataset(ctxs[r], data[bounds[r]:bounds[r + 1]]) for r in range(3)]
        got = run_all(lambda r: DeviceShard(dss[r], bounds[r], data[bounds[r]:bounds[r + 1]]))
        ref = run_all(lambda r: OracleShard(data[bounds[r]:bounds[r + 1]], bounds[r]))
        for r in range(3):
            assert all(np.array_equal(a, b) for a, b in zip(got[r][0], ref[r][0]))
            assert np.array_equal(got[r][1].view(np.uint32), ref[r][1].view(np.uint32))
        for d_ in dss:
            d_.free()
    finally:
        for c_ in ctxs:
            c_.close()


@pytest.mark.parametrize("metric", METRICS)
def test_sharded_bisect_matches_oracle(spf, oracle, metric):
    """§8(e) bisect: three device shards run assign + update + the bisect work-list
    (spf_farthest_from, spf_assign_vectors on the local member slices) == the oracle-backed shards
    == the single-process oracle fit."""
    import sys
    import threading
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_multirank_cpu import run_sharded_fit
    from spfresh_b200.sharded import DeviceShard, ThreadComm
    data = clustered(3000, 14, 10, 91 + metric)
    init = np.random.default_rng(5).choice(3000, 4, replace=False)
    bounds = [0, 800, 1900, 3000]

    def run_all(make_shard):
        grp = ThreadComm.Group(3)
        out = [None] * 3

        def run(r):
            out[r] = run_sharded_fit(ThreadComm(grp, r), data, init, bounds, metric, 250, make_shard)
        th = [threading.Thread(target=run, args=(r,)) for r in range(3)]
        [t.start() for t in th]
        [t.join() for t in th]
        return out
    ctxs = [spf.Context(0) for _ in range(3)]
    try:
        dss = [spf.Dataset(ctxs[r], data[bounds[r]:bounds[r + 1]]) for r in range(3)]
        got = run_all(lambda r: DeviceShard(dss[r], bounds[r], data[bounds[r]:bounds[r + 1]]))
        ref = run_all(None)
        assert got[0] == got[1] == got[2] and got[0] == ref[0]
        full = oracle.fit(data, metric, init, 250, pick=lambda m: (m * 5) // 7)
        assert got[0] == [(int(c.centroid_idx), np.asarray(c.points).tolist(), int(c.depth)) for c in full]
        assert len(full) > 4
        for d_ in dss:
            d_.free()
    finally:
        for c_ in ctxs:
            c_.close()


@pytest.mark.parametrize("metric", METRICS)
def test_sharded_kmeanspp_matches_oracle_shards(spf, oracle, metric):
    """Sharded k-means++ on three device shards == the oracle-backed shards, and (one shard) == the
    single-process device k-means++ session."""
    import sys
    import threading
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from shard_ref import OracleShard
    from spfresh_b200.sharded import DeviceShard, SingleComm, ThreadComm, kmeans_plus_plus
    data = clustered(2300, 12, 9, 55 + metric)
    bounds = [0, 900, 1500, 2300]
    u01 = [0.37, 0.91, 0.08, 0.55, 0.73, 0.21, 0.66]

    def run_all(make_shard):
        grp = ThreadComm.Group(3)
        out = [None] * 3

        def run(r):
            out[r] = kmeans_plus_plus(make_shard(r), ThreadComm(grp, r), metric, 8,
                                      spf.ScriptedRandomSource(index=[1234], u01=u01))
        th = [threading.Thread(target=run, args=(r,)) for r in range(3)]
        [t.start() for t in th]
        [t.join() for t in th]
        return out
    ctxs = [spf.Context(0) for _ in range(3)]
    try:
        dss = [spf.Dataset(ctxs[r], data[bounds[r]:bounds[r + 1]]) for r in range(3)]
        got = run_all(lambda r: DeviceShard(dss[r], bounds[r], data[bounds[r]:bounds[r + 1]]))
        ref = run_all(lambda r: OracleShard(data[bounds[r]:bounds[r + 1]], bounds[r]))
        for r in range(3):
            assert np.array_equal(got[r], ref[r])
        whole = spf.Dataset(ctxs[0], data)
        one = kmeans_plus_plus(DeviceShard(whole, 0, data), SingleComm(), metric, 8,
                               spf.ScriptedRandomSource(index=[1234], u01=u01))
        sess = whole.kmeanspp(metric, 1234)
        rows = [1234] + [sess.round(u) for u in u01]
        sess.free()
        assert np.array_equal(one, np.array(rows, np.uint64))
        whole.free()
        for d_ in dss:
            d_.free()
    finally:
        for c_ in ctxs:
            c_.close()


# ----------------------------------------------------------------------------------------------
# update_centroids / farthest / k-means++
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric", METRICS)
def test_update_medoids_matches_oracle(spf, ctx, oracle, metric):
    data = clustered(6000, 40, 30, 21)
    cent = np.random.default_rng(3).choice(6000, 50, replace=False)
    ds = spf.Dataset(ctx, data)
    res = ds.assign(metric, cent)
    f = res.fetch()
    ref_rows, ref_means = oracle.update_medoids(data, metric, f.offsets, f.members, cent, want_means=True)
    rows, means = ds.update_medoids_from(metric, res, cent, want_means=True)
    assert np.array_equal(rows, ref_rows)
    assert np.array_equal(means.view(np.uint32), ref_means.view(np.uint32))
    rows2 = ds.update_medoids(metric, f.offsets, f.members, cent)
    assert np.array_equal(rows2, ref_rows)
    # an empty cluster keeps its old centroid (hierarchical.rs:146-149)
    off = np.array([0, 0, 3], np.uint64)
    mem = np.array([5, 9, 5], np.uint64)
    assert np.array_equal(ds.update_medoids(metric, off, mem, [77, 1]),
                          oracle.update_medoids(data, metric, off, mem, [77, 1]))


@pytest.mark.parametrize("direct", [1, 0])
@pytest.mark.parametrize("d", [3, 130, 300])
@pytest.mark.parametrize("metric", METRICS)
def test_medoid_staging_variants_match_oracle(spf, ctx, oracle, metric, d, direct):
    """medoid_kernel with the member rows staged alone and the cluster mean read in place (medoid_direct = 1,
    128 dimensions per step: one, two and three steps here) and with both rows of every pair staged (0)."""
    data = clustered(3000, d, 12, 77 + d)
    cent = np.random.default_rng(5).choice(3000, 24, replace=False)
    ds = spf.Dataset(ctx, data)
    res = ds.assign(metric, cent)
    f = res.fetch()
    ref_rows = oracle.update_medoids(data, metric, f.offsets, f.members, cent)
    ctx.set_param("medoid_direct", direct)
    try:
        assert np.array_equal(ds.update_medoids_from(metric, res, cent), ref_rows)
        assert np.array_equal(ds.update_medoids(metric, f.offsets, f.members, cent), ref_rows)
    finally:
        ctx.set_param("medoid_direct", 1)
        res.free()
        ds.free()


@pytest.mark.parametrize("fast", [1, 0])
@pytest.mark.parametrize("d,hub", [(128, 64), (126, 0), (40, 64), (260, 64)])
def test_cluster_mean_producer_variants_match_oracle(spf, ctx, oracle, d, hub, fast):
    """compute_mean with the suspended-wait / short-loop producers (sum_fast = 1; rows of 32 128-bit columns take
    the short loop) and with the polling producers (0); sum_hub = 64 sends every cluster of >= 64 members through
    the deep 12-producer configuration that real data reserves for hub clusters."""
    data = clustered(7000, d, 20, 300 + d)
    cent = np.random.default_rng(11).choice(7000, 40, replace=False)
    ds = spf.Dataset(ctx, data)
    res = ds.assign(0, cent)
    f = res.fetch()
    ref_rows, ref_means = oracle.update_medoids(data, 0, f.offsets, f.members, cent, want_means=True)
    ctx.set_param("sum_fast", fast)
    ctx.set_param("sum_hub", hub)
    try:
        rows, means = ds.update_medoids_from(0, res, cent, want_means=True)
        assert np.array_equal(means.view(np.uint32), ref_means.view(np.uint32))
        assert np.array_equal(rows, ref_rows)
    finally:
        ctx.set_param("sum_fast", 1)
        ctx.set_param("sum_hub", 0)
        res.free()
        ds.free()


def test_mean_kat_on_device(spf, ctx, kats):
    kat = kats["mean"][0]                          # utils.rs:24-32
    data = np.array(kat["data"], np.float32)
    ds = spf.Dataset(ctx, data)
    _, means = ds.update_medoids(0, [0, 2], kat["indices"], [0], want_means=True)
    assert np.all(np.abs(means[0] - np.array(kat["expected"])) < kat["tol"])


@pytest.mark.parametrize("metric", METRICS)
def test_farthest_matches_oracle(spf, ctx, oracle, metric):
    data = gauss(3000, 20, 4)
    data[10] = data[11]
    ds = spf.Dataset(ctx, data)
    mem = np.random.default_rng(1).permutation(3000)[:1500]
    for c1 in (int(mem[0]), int(mem[700]), 2999):
        assert ds.farthest(metric, c1, mem) == oracle.farthest(data, metric, c1, mem)
    assert ds.farthest(metric, 10, [10, 11]) == 0          # all distances 0 → identity row 0
    assert ds.farthest(metric, 10, [10]) == 0


@pytest.mark.parametrize("metric", METRICS)
def test_farthest_and_kmeanspp_long_rows_match_oracle(spf, ctx, oracle, metric):
    """Rows of 600 floats: five staging steps of the row-staged distance (farthest_kernel, and the generic
    kmpp_update_kernel that serves rows too long for the tiled update)."""
    data = clustered(2500, 600, 10, 31 + metric)
    ds = spf.Dataset(ctx, data)
    mem = np.random.default_rng(2).permutation(2500)[:1200]
    for c1 in (int(mem[3]), 2499):
        assert ds.farthest(metric, c1, mem) == oracle.farthest(data, metric, c1, mem)
    k = 9
    u = np.random.default_rng(40 + metric).random(k - 1)
    ref, fell = oracle.kmeanspp(data, metric, k, 77, u)
    assert not fell.any()
    sess = ds.kmeanspp(metric, 77)
    got = [77] + [sess.round(u[r]) for r in range(k - 1)]
    sess.free()
    ds.free()
    assert got == ref.tolist()


@pytest.mark.parametrize("metric", METRICS)
def test_kmeanspp_matches_oracle(spf, ctx, oracle, metric):
    data = clustered(20000, 32, 25, 8)
    k = 24
    u = np.random.default_rng(metric).random(k - 1)
    ref, fell = oracle.kmeanspp(data, metric, k, 123, u)
    assert not fell.any()
    ds = spf.Dataset(ctx, data)
    sess = ds.kmeanspp(metric, 123)
    got = [123]
    for r in range(k - 1):
        got.append(sess.round(u[r]))
    sess.free()
    assert got == ref.tolist()
    # degenerate: identical points → weighted pick impossible → host fallback is requested
    same = np.ones((64, 4), np.float32)
    sess = spf.Dataset(ctx, same).kmeanspp(0, 3)
    assert sess.round(0.5) is None
    sess.push(9)
    assert sess.round(0.25) is None
    sess.free()


@pytest.mark.parametrize("metric", METRICS)
def test_kmeanspp_batched_rounds_match_oracle(spf, ctx, oracle, metric):
    """spf_kmpp_rounds (device-resident rounds, one host sync per batch, cluster scan for the fold:
    40 000 rows > one single-CTA window) against the oracle and against single rounds."""
    data = clustered(40000, 24, 30, 9)
    k = 41
    u = np.random.default_rng(50 + metric).random(k - 1)
    ref, fell = oracle.kmeanspp(data, metric, k, 77, u)
    assert not fell.any()
    ds = spf.Dataset(ctx, data)
    sess = ds.kmeanspp(metric, 77)
    got = [77]
    for lo, hi in ((0, 1), (1, 18), (18, 40)):             # uneven batches, the state carries over
        rows, failed = sess.rounds(u[lo:hi])
        assert not failed and len(rows) == hi - lo
        got += [int(r) for r in rows]
    sums = sess.last_sums()
    sess.free()
    assert got == ref.tolist()
    one = ds.kmeanspp(metric, 77)
    for r in range(k - 1):
        assert one.round(u[r]) == got[r + 1]
    assert one.last_sums() == sums
    one.free()
    # degenerate: identical points -> the first round of the batch cannot pick, nothing else runs
    same = np.ones((64, 4), np.float32)
    sess = spf.Dataset(ctx, same).kmeanspp(0, 3)
    rows, failed = sess.rounds([0.5, 0.25, 0.125])
    assert failed and len(rows) == 0
    sess.push(9)
    rows, failed = sess.rounds([0.25])
    assert failed and len(rows) == 0
    sess.free()
    # a batch that fails in the middle: two distinct points, the third round has all-zero weights
    two = np.concatenate([np.zeros((40, 4), np.float32), np.ones((40, 4), np.float32)])
    sess = spf.Dataset(ctx, two).kmeanspp(0, 0)
    rows, failed = sess.rounds([0.5, 0.5, 0.5, 0.5])
    assert failed and len(rows) == 1 and rows[0] >= 40
    sess.push(5)
    rows, failed = sess.rounds([0.3])
    assert failed and len(rows) == 0
    sess.free()


@pytest.mark.parametrize("metric", METRICS)
def test_kmeanspp_device_resident_sharded_rounds_one_rank(spf, ctx, oracle, metric):
    """spf_kmpp_set_vector + spf_kmpp_rounds_sharded with one rank (comm NULL: the all-gathers are device
    copies): the picks of the oracle; two ranks over NCCL are checked by tools/check_sharded_nccl.py."""
    from spfresh_b200.device import KmppShardSession
    data = clustered(30000, 20, 30, 19)
    k = 33
    u = np.random.default_rng(70 + metric).random(k - 1)
    ref, fell = oracle.kmeanspp(data, metric, k, 4321, u)
    assert not fell.any()
    ds = spf.Dataset(ctx, data)
    sess = KmppShardSession(ds, metric)
    with pytest.raises(spf.SpfError):
        sess.rounds_sharded(None, 0, u[:2])                 # no centroid yet
    sess.set_vector(data[4321])
    got = [4321]
    for lo, hi in ((0, 5), (5, 32)):
        rows, failed = sess.rounds_sharded(None, 0, u[lo:hi])
        assert not failed and len(rows) == hi - lo
        got += [int(r) for r in rows]
    sess.free()
    assert got == ref.tolist()
    same = np.ones((64, 4), np.float32)
    sess = KmppShardSession(spf.Dataset(ctx, same), 0)
    sess.set_vector(same[3])
    rows, failed = sess.rounds_sharded(None, 1000, [0.5, 0.25])
    assert failed and len(rows) == 0
    with pytest.raises(spf.SpfError):
        sess.rounds_sharded(None, 1000, [0.5])              # the host must set the uniformly drawn row first
    sess.free()


def test_sequential_sum_scan_is_bit_exact(spf, ctx):
    """hierarchical.rs:278 is a strictly sequential f32 fold.  The scan-based kernel (two-state
    transducers per binade) must return the bits of the serial add chain and of numpy's sequential
    cumsum for every input: ties, zeros, denormals, huge dynamic range, overflow, invalid values."""
    rng = np.random.default_rng(7)
    cases = []
    for n in (0, 1, 2, 3, 7, 8, 9, 1023, 8192, 8193, 100_000, 1_000_003):
        cases.append(rng.random(n, dtype=np.float32) * 300.0)                       # distances of the bench shape
    cases.append((rng.integers(0, 4096, 300_000) * 0.25).astype(np.float32))           # many exact ties
    cases.append((rng.integers(0, 3, 200_000)).astype(np.float32))                      # zeros and small integers
    cases.append(np.exp(rng.normal(0, 12, 200_000)).astype(np.float32))                 # 30 binades of range
    cases.append(np.concatenate([np.zeros(5000, np.float32), rng.random(50_000, dtype=np.float32) * 1e-41]))   # denormals
    cases.append(np.concatenate([rng.random(10_000, dtype=np.float32) * 1e-40, rng.random(10_000, dtype=np.float32)]))
    cases.append(np.full(70_000, 16777216.0, np.float32))                               # 2^24: every add is a tie or exact
    cases.append(np.concatenate([[1.0], np.full(100_000, 2.0 ** -24, np.float32)]).astype(np.float32))   # half-ulp ties: never grows
    cases.append(np.concatenate([[1.0], np.full(100_000, 2.0 ** -24 * 1.5, np.float32)]).astype(np.float32))
    cases.append(np.full(3000, 3.0e38, np.float32))                                     # overflows to +inf
    cases.append(np.concatenate([rng.random(20_000, dtype=np.float32), [-1.0], rng.random(20_000, dtype=np.float32)]).astype(np.float32))
    cases.append(np.concatenate([rng.random(9000, dtype=np.float32), [np.inf], rng.random(100, dtype=np.float32)]).astype(np.float32))
    cases.append(np.concatenate([rng.random(9000, dtype=np.float32), [np.nan], rng.random(100, dtype=np.float32)]).astype(np.float32))
    cases.append(np.concatenate([[0.0, -0.0, 0.0], rng.random(5000, dtype=np.float32)]).astype(np.float32))
    for i, v in enumerate(cases):
        v = np.ascontiguousarray(v, np.float32)
        with np.errstate(over="ignore", invalid="ignore"):
            want = np.cumsum(v, dtype=np.float32)[-1] if v.size else np.float32(0.0)
        serial = ctx.seq_sum_f32(v, 2)
        scan = ctx.seq_sum_f32(v, 1)
        cluster = ctx.seq_sum_f32(v, 3)                    # thread-block-cluster scan (the k-means++ rounds' kernel)
        assert np.array_equal(np.array([serial]).view(np.uint32), np.array([want]).view(np.uint32)) or \
            (np.isnan(serial) and np.isnan(want)), (i, serial, want)
        assert np.array_equal(np.array([scan]).view(np.uint32), np.array([serial]).view(np.uint32)) or \
            (np.isnan(scan) and np.isnan(serial)), (i, v.size, scan, serial)
        assert np.array_equal(np.array([cluster]).view(np.uint32), np.array([serial]).view(np.uint32)) or \
            (np.isnan(cluster) and np.isnan(serial)), (i, v.size, cluster, serial)


# ----------------------------------------------------------------------------------------------
# fit() end to end through the host mirror
# ----------------------------------------------------------------------------------------------
def as_tuples(clusters):
    return [(int(c.centroid_idx), np.asarray(c.points).tolist(), int(c.depth)) for c in clusters]


@pytest.mark.parametrize("metric_cls", ["SquaredEuclideanDistance", "ManhattanDistance", "ChebyshevDistance"])
def test_fit_matches_oracle(spf, ctx, oracle, metric_cls):
    data = clustered(4000, 16, 12, 31)
    init = np.random.default_rng(4).choice(4000, 5, replace=False).tolist()
    pick = lambda n: (n * 5) // 7  # noqa: E731
    metric = getattr(spf, metric_cls)()
    params = spf.ClusteringParams(metric, spf.InitializationMethod.Random, 300, 5,
                                  random_source=spf.ScriptedRandomSource(multiple=init, index=pick))
    hc = spf.HierarchicalClustering(params, data, ctx=ctx)
    hc.fit()
    ref = oracle.fit(data, metric.kind, init, 300, pick=pick)
    assert as_tuples(hc.clusters) == as_tuples(ref)
    assert len(hc.clusters) > 5                       # the bisect did fire
    lab = hc.labels()
    assert lab.shape == (4000,) and lab.min() >= 0 and lab.max() < len(hc.clusters)


def test_fit_kmeanspp_end_to_end(spf, ctx, oracle):
    data = clustered(3000, 12, 9, 5)
    u = np.random.default_rng(9).random(7).tolist()
    params = spf.ClusteringParams(spf.SquaredEuclideanDistance(), spf.InitializationMethod.KMeansPlusPlus, 600, 8,
                                  random_source=spf.ScriptedRandomSource(index=[17] + [0] * 50, u01=u))
    hc = spf.HierarchicalClustering(params, data, ctx=ctx)
    hc.fit()
    init, _ = oracle.kmeanspp(data, 0, 8, 17, u)
    ref = oracle.fit(data, 0, init, 600, pick=lambda n: 0)
    assert as_tuples(hc.clusters) == as_tuples(ref)


def test_reference_unit_tests_on_device(spf, ctx, kats):
    """The reference's own hierarchical.rs tests, run against the device implementation."""
    data = np.array(kats["toy_data"]["rows"], np.float32)
    # test_subdivide_clusters (:444-463)
    for init in range(6):
        p = spf.ClusteringParams(spf.SquaredEuclideanDistance(), spf.InitializationMethod.Random, 2, 1,
                                 random_source=spf.ScriptedRandomSource(multiple=[init], index=lambda n: n // 2))
        hc = spf.HierarchicalClustering(p, data, ctx=ctx)
        hc.initialize_clusters(1)
        hc.assign_points()
        hc.update_centroids()
        hc.subdivide_clusters()
        assert len(hc.clusters) > 1 and all(len(c.points) <= 2 for c in hc.clusters)
    # test_initialize_clusters_randomly / kmeans_plus_plus (:405-441) with the default RNG
    for meth in (spf.InitializationMethod.Random, spf.InitializationMethod.KMeansPlusPlus):
        p = spf.ClusteringParams(spf.SquaredEuclideanDistance(), meth, 3, 2, rng_seed=42)
        hc = spf.HierarchicalClustering(p, data, ctx=ctx)
        hc.initialize_clusters(2)
        assert len(hc.clusters) == 2 and all(c.centroid_idx is not None for c in hc.clusters)
    # test_fit (:489-507) invariants with the default RNG: every cluster <= desired size
    p = spf.ClusteringParams(spf.SquaredEuclideanDistance(), spf.InitializationMethod.KMeansPlusPlus, 2, 3, rng_seed=42)
    hc = spf.HierarchicalClustering(p, data, ctx=ctx)
    hc.fit()
    assert all(len(c.points) <= 2 for c in hc.clusters)


def test_example_build_index_kat_on_device(spf, ctx, kats, tmp_path):
    """examples/build_index.rs through SpannIndexBuilder: point_id 0 / [1.0, 2.0] for any draws."""
    kat = kats["example_query"][0]
    data = np.array(kats["toy_data"]["rows"], np.float32)
    for init in [(0, 1, 2, 3), (5, 3, 1, 0), (2, 4, 5, 1)]:
        cfg = spf.Config(spf.ClusteringParamsConfig("Euclidean", "Random", 4), None, str(tmp_path / "idx"))
        b = spf.SpannIndexBuilder(cfg, ctx=ctx,
                                  random_source=spf.ScriptedRandomSource(multiple=list(init), index=lambda n: 0))
        index = b.with_data(data).build(2)
        res = index.find_k_nearest_neighbor_spann(np.array(kat["query"], np.float32), kat["k"])
        assert res == [spf.PointData(kat["expected_point_id"], kat["expected_vector"])]
        # load::<N>() from the files just written gives the same answer
        loaded = spf.SpannIndexBuilder(cfg, ctx=ctx).load(2)
        res2 = loaded.find_k_nearest_neighbor_spann(np.array(kat["query"], np.float32), kat["k"])
        assert res2 == res
    with pytest.raises(ValueError):
        spf.SpannIndexBuilder(cfg, ctx=ctx).with_data(data).build(3)      # dimension mismatch


# ----------------------------------------------------------------------------------------------
# query path
# ----------------------------------------------------------------------------------------------
def build_lists(oracle, data, k, seed):
    cent = np.random.default_rng(seed).choice(data.shape[0], k, replace=False)
    r = oracle.assign(data, 0, cent)
    return cent, r.offsets, r.members


@pytest.mark.parametrize("n,d,nlists,topk,nprobe", [
    (3000, 16, 40, 10, 0), (5000, 128, 64, 10, 0), (5000, 33, 50, 5, 20), (4000, 96, 30, 40, 8),
    (2000, 8, 300, 100, 0), (6000, 16, 1500, 10, 40), (40000, 16, 30, 10, 6),
])
def test_search_matches_oracle(spf, ctx, oracle, n, d, nlists, topk, nprobe):
    data = clustered(n, d, 20, n + d)
    cent, off, mem = build_lists(oracle, data, nlists, d)
    ds = spf.Dataset(ctx, data)
    idx = spf.DeviceIndex.pack(ds, off, mem, cent)
    q = clustered(200, d, 20, n + d)              # same centres → realistic probes
    q[0] = data[cent[3]]                          # exact hit on a centroid: thr = 1.2 * eps
    q[1] = 100.0                                  # far away
    ids, dists, counts, vec = idx.search(q, topk, nprobe, want_vectors=True)
    rid, rd, rc = oracle.search_batch(data, off, mem, cent, q, topk, nprobe)
    assert np.array_equal(counts, rc)
    for i in range(q.shape[0]):
        c = int(rc[i])
        assert np.array_equal(ids[i, :c], rid[i, :c]), i
        assert np.array_equal(dists[i, :c].view(np.uint32), rd[i, :c].view(np.uint32)), i
        assert np.array_equal(vec[i, :c], data[rid[i, :c].astype(np.int64)])
    assert idx.last_scan_bytes() > 0
    # no pruning at all
    ids2, d2, c2 = idx.search(q, topk, nprobe, prune_factor=float("inf"))
    rid2, rd2, rc2 = oracle.search_batch(data, off, mem, cent, q, topk, nprobe, prune_factor=float("inf"))
    assert np.array_equal(c2, rc2)
    for i in range(q.shape[0]):
        assert np.array_equal(ids2[i, :rc2[i]], rid2[i, :rc2[i]])


@pytest.mark.parametrize("nprobe", [17, 64, 256, 300, 1024])
def test_search_probe_selection_large_nprobe(spf, oracle, nprobe):
    """probe_topn_kernel (select the nprobe smallest (distance, list id) keys, sort only those) ==
    the oracle's full stable sort of the centroid distances, for 1024 < nlists <= 4096."""
    c2 = spf.Context(0)
    try:
        data = clustered(9000, 8, 30, 404)
        data[100:140] = data[60:100]                       # duplicated centroid vectors: equal distances, id order
        nlists = 3000
        cent = np.arange(nlists, dtype=np.uint64) * 3 % 9000
        r = oracle.assign(data, 0, cent)
        ds = spf.Dataset(c2, data)
        idx = spf.DeviceIndex.pack(ds, r.offsets, r.members, cent)
        q = clustered(64, 8, 30, 404)
        rid, rd, rc = oracle.search_batch(data, r.offsets, r.members, cent, q, 10, nprobe, prune_factor=float("inf"))
        for mode in (0, 2):              # exact selection kernel / tensor-core probe (dense variant for nprobe > 32)
            c2.set_param("scan_tc", mode)
            c2.set_profiling(True)
            ids, dists, counts, keys = idx.search(q, 10, nprobe, prune_factor=float("inf"), want_keys=True)
            assert (c2.kernel_ms("probe_tc_select") > 0) == (mode == 2 and nprobe > 32)
            c2.set_profiling(False)
            assert np.array_equal(counts, rc)
            for i in range(64):
                assert np.array_equal(ids[i, :rc[i]], rid[i, :rc[i]]), (mode, i)
                assert np.array_equal(dists[i, :rc[i]].view(np.uint32), rd[i, :rc[i]].view(np.uint32)), (mode, i)
        idx.free()
        ds.free()
    finally:
        c2.close()


def test_search_list_major_equals_query_major(spf, oracle):
    """The list-major scan (lists shared by query batches) and the query-major scan are the same
    function: identical ids, distance bits, counts and merge keys; the first is oracle-checked."""
    c2 = spf.Context(0)
    try:
        data = clustered(20000, 64, 50, 21)
        cent, off, mem = build_lists(oracle, data, 120, 3)
        ds = spf.Dataset(c2, data)
        idx = spf.DeviceIndex.pack(ds, off, mem, cent)
        q = clustered(1500, 64, 50, 22)
        out = {}
        for mode in (0, 2):
            c2.set_param("scan_list_major", mode)
            out[mode] = idx.search(q, 10, 16, want_keys=True)
        for x, y in zip(out[0], out[2]):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
        rid, rd, rc = oracle.search_batch(data, off, mem, cent, q[:200], 10, 16)
        assert np.array_equal(out[2][2][:200], rc)
        for i in range(200):
            assert np.array_equal(out[2][0][i, :rc[i]], rid[i, :rc[i]])
        idx.free()
        ds.free()
    finally:
        c2.close()


@pytest.mark.parametrize("n,d,nlists,topk,nprobe,kind", [
    (6000, 16, 12, 10, 4, "clustered"), (20000, 128, 40, 10, 8, "clustered"), (8000, 33, 20, 5, 20, "gauss"),
    (12000, 96, 30, 16, 6, "gauss"), (5000, 8, 10, 1, 3, "gauss"), (30000, 64, 150, 10, 32, "clustered"),
    (3000, 100, 7, 3, 0, "gauss"), (4000, 200, 16, 10, 6, "gauss"), (3000, 960, 12, 5, 4, "clustered"),
    (2500, 516, 9, 8, 40, "gauss"),
])
def test_search_tensor_scan_matches_oracle(spf, oracle, n, d, nlists, topk, nprobe, kind):
    """scan_tc.cu: the TF32 candidate scan + exact refinement is the same function as the exact
    scans — identical ids, distance bits, counts and merge keys — and matches the oracle."""
    c2 = spf.Context(0)
    try:
        data = clustered(n, d, 20, n + d) if kind == "clustered" else gauss(n, d, n + d)
        cent, off, mem = build_lists(oracle, data, nlists, d)
        ds = spf.Dataset(c2, data)
        idx = spf.DeviceIndex.pack(ds, off, mem, cent)
        q = (clustered(700, d, 20, n + d) if kind == "clustered" else gauss(700, d, 5)).copy()
        q[0] = data[cent[3]]                          # exact hit on a centroid: thr = 1.2 * eps
        q[1] = 100.0                                  # far away
        q[2] = data[int(mem[0])]                      # exact hit on a member
        for pf in (1.2, float("inf")):
            out = {}
            # exact scans / one GEMM pass + group refinement / two GEMM passes
            for mode, cmax_mb in ((0, 0), (2, 16384), (2, 0)):
                c2.set_param("scan_tc", mode)
                c2.set_param("scan_tc_cmax_mb", cmax_mb)
                c2.set_profiling(True)
                out[(mode, cmax_mb)] = idx.search(q, topk, nprobe, prune_factor=pf, want_keys=True)
                assert (c2.kernel_ms("scan_tc_b") > 0) == (mode == 2)
                assert (c2.kernel_ms("probe_tc_b") > 0 or c2.kernel_ms("probe_tc_select") > 0) == (mode == 2)   # tensor-core probe
                assert (c2.kernel_ms("scan_tc_groups") > 0) == (mode == 2 and cmax_mb > 0 and d <= 256)
                c2.set_profiling(False)
            for key in ((2, 16384), (2, 0)):
                for x, y in zip(out[(0, 0)], out[key]):
                    assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), key
            out[2] = out[(2, 16384)]
            rid, rd, rc = oracle.search_batch(data, off, mem, cent, q[:150], topk, nprobe, prune_factor=pf)
            assert np.array_equal(out[2][2][:150], rc)
            for i in range(150):
                assert np.array_equal(out[2][0][i, :rc[i]], rid[i, :rc[i]]), i
                assert np.array_equal(out[2][1][i, :rc[i]].view(np.uint32), rd[i, :rc[i]].view(np.uint32)), i
        idx.free()
        ds.free()
    finally:
        c2.close()


def test_search_tensor_scan_random_shapes(spf, oracle):
    """Randomised shapes (tiny and empty lists, k larger than a list, one list, nprobe = all lists,
    duplicate rows, ragged dimensions): tensor scan + tensor probe == exact kernels, bit for bit."""
    rng = np.random.default_rng(2024)
    c2 = spf.Context(0)
    try:
        for case in range(14):
            n = int(rng.integers(40, 4000))
            d = int(rng.choice([1, 2, 3, 7, 16, 31, 64, 100, 128]))
            nlists = int(rng.integers(1, min(n, 70) + 1))
            topk = int(rng.integers(1, 17))
            nprobe = int(rng.choice([0, 1, 2, 5, nlists]))
            data = clustered(n, d, 5, 1000 + case) if case % 2 else gauss(n, d, 1000 + case)
            if case % 3 == 0:
                data[n // 2:] = data[:n - n // 2]            # duplicate rows: exact ties everywhere
            cent = rng.choice(n, nlists, replace=False)
            r = oracle.assign(data, 0, cent)
            off, mem = r.offsets.copy(), r.members.copy()
            if nlists > 2:                                    # empty a list: its members vanish from the index
                lo, hi = int(off[1]), int(off[2])
                mem = np.concatenate([mem[:lo], mem[hi:]])
                off[2:] -= np.uint64(hi - lo)
            ds = spf.Dataset(c2, data)
            idx = spf.DeviceIndex.pack(ds, off, mem, cent)
            q = np.concatenate([data[rng.integers(0, n, 150)], gauss(150, d, 77 + case)]).astype(np.float32)
            pf = float(rng.choice([1.0, 1.2, 3.0, np.inf]))
            out = {}
            for mode, cmax_mb in ((0, 0), (2, 16384), (2, 0)):
                c2.set_param("scan_tc", mode)
                c2.set_param("scan_tc_cmax_mb", cmax_mb)
                out[(mode, cmax_mb)] = idx.search(q, topk, nprobe, prune_factor=pf, want_keys=True)
            for key in ((2, 16384), (2, 0)):
                for x, y in zip(out[(0, 0)], out[key]):
                    assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), (case, n, d, nlists, topk, nprobe, pf, key)
            idx.free()
            ds.free()
    finally:
        c2.close()


def test_search_tensor_scan_fallbacks(spf, oracle):
    """Bucket overflow, a bound pass over a subset of the probes, non-finite queries and huge norms
    all end in the exact result (flagged queries re-run on the exact query-major kernel)."""
    c2 = spf.Context(0)
    try:
        data = gauss(9000, 48, 77)
        data[17] *= 3.0e4                             # large norm: loose bound for everybody
        cent, off, mem = build_lists(oracle, data, 16, 5)
        ds = spf.Dataset(c2, data)
        idx = spf.DeviceIndex.pack(ds, off, mem, cent)
        q = gauss(400, 48, 78).copy()
        q[3, 5] = np.inf
        q[4, 7] = np.nan
        q[5] = 1.0e18                                 # |q|^2 overflows
        q[6] = 0.0
        c2.set_param("scan_tc", 0)
        want = idx.search(q, 10, 6, want_keys=True)
        for cmax_mb in (16384, 0):
            for name, value in (("scan_tc_bucket", 4), ("scan_tc_bucket", 256), ("scan_tc_tau_probes", 1),
                                ("scan_tc_tau_probes", 3)):
                c2.set_param("scan_tc", 2)
                c2.set_param("scan_tc_cmax_mb", cmax_mb)
                c2.set_param(name, value)
                got = idx.search(q, 10, 6, want_keys=True)
                for x, y in zip(want, got):
                    assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), (cmax_mb, name, value)
        idx.free()
        ds.free()
    finally:
        c2.close()


@pytest.mark.parametrize("cmax_mb", [16384, 0])
def test_search_tensor_scan_split_launch(spf, oracle, cmax_mb):
    """The bound pass as two concurrent launches (units of multi-unit lists on some SMs, single-unit lists
    on the others; scan_tc.cu: scan_tc_run) returns what the single launch and the exact scans return."""
    c2 = spf.Context(0)
    try:
        # 60 lists, 3000 queries x 4 probes on clustered data: popular lists get several units of 128
        # pairs, the rest a single one
        data = clustered(24000, 32, 40, 11)
        cent, off, mem = build_lists(oracle, data, 60, 13)
        ds = spf.Dataset(c2, data)
        idx = spf.DeviceIndex.pack(ds, off, mem, cent)
        q = clustered(3000, 32, 40, 11)[::-1].copy()
        c2.set_param("scan_tc", 0)
        want = idx.search(q, 10, 4, want_keys=True)
        c2.set_param("scan_tc", 2)
        c2.set_param("scan_tc_cmax_mb", cmax_mb)
        for split in (0, 3, 100):
            c2.set_param("scan_tc_split", split)
            c2.set_profiling(True)
            got = idx.search(q, 10, 4, want_keys=True)
            hub, units = c2.kernel_ms("scan_tc_hub_units"), c2.kernel_ms("scan_tc_units")
            assert 0 < hub < units, (hub, units)          # both classes of lists are present
            assert (c2.kernel_ms("scan_tc_split") > 0) == (split > 0), split
            c2.set_profiling(False)
            for x, y in zip(want, got):
                assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), split
        rid, rd, rc = oracle.search_batch(data, off, mem, cent, q[:100], 10, 4)
        assert np.array_equal(got[2][:100], rc)
        for i in range(100):
            assert np.array_equal(got[0][i, :rc[i]], rid[i, :rc[i]]), i
            assert np.array_equal(got[1][i, :rc[i]].view(np.uint32), rd[i, :rc[i]].view(np.uint32)), i
        idx.free()
        ds.free()
    finally:
        c2.close()


def test_search_tensor_scan_sharded_lists(spf, oracle):
    """List shards scanned by the tensor path merge to the unsharded exact answer."""
    c2 = spf.Context(0)
    try:
        data = clustered(15000, 32, 25, 9)
        cent, off, mem = build_lists(oracle, data, 24, 7)
        ds = spf.Dataset(c2, data)
        q = clustered(900, 32, 25, 9)
        c2.set_param("scan_tc", 0)
        full = spf.DeviceIndex.pack(ds, off, mem, cent)
        ids, dists, counts = full.search(q, 10, 8)
        c2.set_param("scan_tc", 2)
        parts, handles = [], [full]
        for lb, le in ((0, 5), (5, 17), (17, 24)):
            part = spf.DeviceIndex.pack(ds, off, mem, cent, list_range=(lb, le))
            handles.append(part)
            parts.append(part.search(q, 10, 8, want_keys=True))
        mids, md, mc = spf.topk_merge(np.stack([p[3] for p in parts]), np.stack([p[0] for p in parts]),
                                      np.stack([p[1] for p in parts]), np.stack([p[2] for p in parts]))
        assert np.array_equal(mc, counts)
        for i in range(q.shape[0]):
            assert np.array_equal(mids[i, :mc[i]], ids[i, :mc[i]])
            assert np.array_equal(md[i, :mc[i]], dists[i, :mc[i]])
        for h in handles:
            h.free()
        ds.free()
    finally:
        c2.close()


def test_search_duplicates_are_kept(spf, ctx, oracle):
    """F7: a boundary-replicated point appears once per probed list it lives in."""
    data = gauss(2000, 8, 12)
    cent, off, mem = build_lists(oracle, data, 6, 1)
    assert mem.size > 2000                         # replication did happen
    ds = spf.Dataset(ctx, data)
    idx = spf.DeviceIndex.pack(ds, off, mem, cent)
    q = gauss(300, 8, 13)
    ids, dists, counts = idx.search(q, 6, 0, prune_factor=float("inf"))
    rid, rd, rc = oracle.search_batch(data, off, mem, cent, q, 6, 0, prune_factor=float("inf"))
    assert np.array_equal(ids, rid) and np.array_equal(counts, rc)
    assert any(len(set(ids[i].tolist())) < 6 for i in range(300))


def test_search_list_sharding_and_merge(spf, ctx, oracle):
    """§8(e): lists sharded over ranks, per-rank top-k merged on the stable key == unsharded."""
    data = clustered(6000, 24, 30, 2)
    cent, off, mem = build_lists(oracle, data, 48, 7)
    ds = spf.Dataset(ctx, data)
    q = clustered(150, 24, 30, 2)
    full = spf.DeviceIndex.pack(ds, off, mem, cent)
    ids, dists, counts = full.search(q, 10)
    parts = []
    for lb, le in ((0, 11), (11, 30), (30, 48)):
        part = spf.DeviceIndex.pack(ds, off, mem, cent, list_range=(lb, le))
        parts.append(part.search(q, 10, want_keys=True))
    mids, md, mc = spf.topk_merge(np.stack([p[3] for p in parts]), np.stack([p[0] for p in parts]),
                                  np.stack([p[1] for p in parts]), np.stack([p[2] for p in parts]))
    assert np.array_equal(mc, counts)
    for i in range(150):
        assert np.array_equal(mids[i, :mc[i]], ids[i, :mc[i]])
        assert np.array_equal(md[i, :mc[i]], dists[i, :mc[i]])


def test_index_files_interoperate_with_reference_layout(spf, ctx, oracle, tmp_path):
    """posting_lists.rs:64-129: files written by the library are byte-identical to the oracle's
    bincode writer, and an index loaded from oracle-written files answers identically."""
    data = gauss(1500, 12, 3)
    cent, off, mem = build_lists(oracle, data, 9, 2)
    ds = spf.Dataset(ctx, data)
    idx = spf.DeviceIndex.pack(ds, off, mem, cent)
    a, b = tmp_path / "ours", tmp_path / "ref"
    b.mkdir()
    idx.save_dir(str(a))
    for j in range(9):
        oracle.posting_list_write(str(b), j, data, mem[int(off[j]):int(off[j + 1])])
        assert open(a / f"posting_list_{j}.bin", "rb").read() == open(b / f"posting_list_{j}.bin", "rb").read()
    oracle.cluster_ids_write(str(b), [8, 3, 0, 1, 2, 4, 5, 6, 7])      # arbitrary hash-map order
    loaded = spf.DeviceIndex.load_dir(ctx, str(b), data[cent.astype(np.int64)])
    q = gauss(64, 12, 4)
    r1, r2 = idx.search(q, 7), loaded.search(q, 7)
    for x, y in zip(r1, r2):
        assert np.array_equal(x, y)


def test_index_shards_saved_into_one_directory_reload_complete(spf, ctx, oracle, tmp_path):
    """two list shards saved one after the other into the same directory must leave a
    complete cluster_ids.bin; the reloaded index answers like the unsharded one."""
    data = gauss(2500, 16, 31)
    cent, off, mem = build_lists(oracle, data, 11, 2)
    ds = spf.Dataset(ctx, data)
    full = spf.DeviceIndex.pack(ds, off, mem, cent)
    d = tmp_path / "shards"
    for lo, hi in ((0, 4), (4, 11)):
        part = spf.DeviceIndex.pack(ds, off, mem, cent, list_range=(lo, hi))
        part.save_dir(str(d))
        part.free()
    ids = np.fromfile(d / "cluster_ids.bin", "<u8")
    assert ids[0] == 11 and sorted(ids[1:].tolist()) == list(range(11))
    loaded = spf.DeviceIndex.load_dir(ctx, str(d), data[cent.astype(np.int64)])
    q = gauss(80, 16, 32)
    for x, y in zip(full.search(q, 9), loaded.search(q, 9)):
        assert np.array_equal(x, y)


def test_index_load_rejects_corrupt_files_and_finds_centroids(spf, ctx, oracle, tmp_path):
    """A truncated / corrupt posting-list file is an error code, never an abort (the count in the file
    is checked against the file size before anything is allocated); a directory without the
    centroids.bin sidecar loads with caller-supplied or recomputed centroids."""
    data = clustered(1200, 8, 6, 77)
    cent = np.random.default_rng(3).choice(1200, 6, replace=False).astype(np.uint64)
    a = oracle.assign(data, 0, cent)
    med = oracle.update_medoids(data, 0, a.offsets, a.members, cent)
    ds = spf.Dataset(ctx, data)
    idx = spf.DeviceIndex.pack(ds, a.offsets, a.members, med)
    good = tmp_path / "good"
    idx.save_dir(str(good))
    bad = tmp_path / "bad"
    bad.mkdir()
    for f in os.listdir(good):
        raw = open(good / f, "rb").read()
        if f == "posting_list_2.bin":
            raw = (2 ** 31).to_bytes(8, "little") + raw[8:]          # an absurd vector count
        open(bad / f, "wb").write(raw)
    with pytest.raises(spf.SpfError):
        spf.DeviceIndex.load_dir(ctx, str(bad), data[med.astype(np.int64)])
    open(bad / "posting_list_2.bin", "wb").write(open(good / "posting_list_2.bin", "rb").read()[:-5])
    with pytest.raises(spf.SpfError):
        spf.DeviceIndex.load_dir(ctx, str(bad), data[med.astype(np.int64)])
    # no sidecar: explicit error, then the two fallbacks
    index = spf.SpannIndex(str(good), ctx)
    with pytest.raises(FileNotFoundError):
        index.load_posting_list(str(good))
    index.load_posting_list(str(good), centroids=data[med.astype(np.int64)])
    q = clustered(50, 8, 6, 78)
    want = idx.search(q, 5)
    got = index.device_index.search(q, 5)
    assert all(np.array_equal(x, y) for x, y in zip(want, got))
    index2 = spf.SpannIndex(str(good), ctx)
    index2.load_posting_list(str(good), recompute_centroids=True)     # medoids of the lists == update_centroids' output
    assert np.array_equal(index2.centroids.view(np.uint32), data[med.astype(np.int64)].view(np.uint32))


# ----------------------------------------------------------------------------------------------
# LIRE operations on the hot-path kernels (src/spann/lire/operations.rs)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric_cls,kind", [("SquaredEuclideanDistance", 0), ("ManhattanDistance", 1),
                                             ("ChebyshevDistance", 2)])
def test_lire_split_and_reassign(spf, ctx, oracle, metric_cls, kind):
    """Split: c1 = vectors[0], c2 = LAST maximum of d(c1, v) over vectors[1..] (Rust max_by),
    `dist1 <= dist2` -> partition 1.  Reassign: FIRST minimum over the candidate centroids (min_by).
    Compared with a per-pair restatement on the oracle's distance function."""
    metric = getattr(spf, metric_cls)()
    rng = np.random.default_rng(11 + kind)
    vecs = clustered(700, 24, 3, 50 + kind)
    vecs[400] = vecs[0]                                   # distance 0 to c1
    vecs[650] = vecs[123]                                 # a duplicated vector: equal distances, ties on both rules
    ids = rng.permutation(10_000)[:700]
    vectors = [(int(i), v) for i, v in zip(ids, vecs)]
    op = spf.Split(7, vectors, metric, (8, 9), ctx=ctx)
    assert op.validate() and not spf.Split(7, vectors, metric, (7, 9), ctx=ctx).validate()
    c1, c2 = op.select_initial_centroids()
    d1 = np.array([oracle.distance(kind, vecs[0], v) for v in vecs], np.float32)
    far = 1 + int(np.flatnonzero(d1[1:] == d1[1:].max())[-1])            # last maximum over skip(1)
    assert np.array_equal(c1, vecs[0]) and np.array_equal(c2, vecs[far])
    p1, p2 = op.assign_vectors(c1, c2)
    d2 = np.array([oracle.distance(kind, vecs[far], v) for v in vecs], np.float32)
    want1 = [int(i) for i, a, b in zip(ids, d1, d2) if a <= b]
    want2 = [int(i) for i, a, b in zip(ids, d1, d2) if not a <= b]
    assert [i for i, _ in p1] == want1 and [i for i, _ in p2] == want2
    assert op.execute() == {7, 8, 9} and op.get_affected_partitions() == {7, 8, 9}
    # all vectors identical: every distance is 0 and max_by returns the last element
    same = spf.Split(1, [(k, vecs[5]) for k in range(6)], metric, (2, 3), ctx=ctx)
    s1, s2 = same.select_initial_centroids()
    assert np.array_equal(s1, vecs[5]) and np.array_equal(s2, vecs[5])
    assert [i for i, _ in same.assign_vectors(s1, s2)[0]] == list(range(6))
    with pytest.raises(spf.LireError):
        spf.Split(1, vectors[:1], metric, (2, 3), ctx=ctx).select_initial_centroids()
    op.free()
    same.free()
    # Reassign
    cands = [(100 + j, vecs[50 + 7 * j]) for j in range(9)] + [(300, vecs[50])]   # posting 300 duplicates posting 100
    for t in (3, 50, 57, 600):
        r = spf.Reassign(t, vecs[t], 5, cands, metric, 1, ctx=ctx)
        dc = np.array([oracle.distance(kind, vecs[t], c) for _, c in cands], np.float32)
        assert r.find_best_posting() == cands[int(np.argmin(dc))][0]
        assert r.get_affected_partitions() == {5, cands[int(np.argmin(dc))][0]}
    with pytest.raises(spf.LireError):
        spf.Reassign(1, vecs[1], 5, [], metric, 1, ctx=ctx).find_best_posting()
    best = spf.reassign_batch(ctx, metric, vecs, np.stack([c for _, c in cands]))
    full = np.array([[oracle.distance(kind, v, c) for _, c in cands] for v in vecs[:200]], np.float32)
    assert np.array_equal(best[:200], full.argmin(axis=1))


# ----------------------------------------------------------------------------------------------
# BASELINE-size parity (VERDICT r1 "parity holes"): the configurations the numbers are quoted on
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cc_cache", [1, 0])
def test_assign_tensor_kernel_variants_match_oracle(spf, oracle, cc_cache):
    """The cached centroid matrix (second call with the same centroids, third with different ones,
    then a point subset) — all bit-identical to the oracle, with and without the cache."""
    c2 = spf.Context(0)
    try:
        c2.set_param("cc_cache", cc_cache)
        for n, d, k, kind in [(20000, 128, 1024, "gauss"), (9000, 96, 300, "clustered"), (6000, 64, 4096, "gauss")]:
            data = gauss(n, d, n + d + 1) if kind == "gauss" else clustered(n, d, max(k // 4, 2), n + d + 1)
            data[17] = data[3]
            rng = np.random.default_rng(k + cc_cache)
            ds = spf.Dataset(c2, data)
            cent = rng.choice(n, k, replace=False)
            cent[0], cent[1] = 3, 17
            ref = oracle.assign(data, 0, cent)
            c2.set_profiling(True)
            got = ds.assign(0, cent).fetch()
            assert c2.kernel_ms("assign_tc") > 0, "the tcgen05 path did not run"
            c2.set_profiling(False)
            check_assign(got, ref)
            check_assign(ds.assign(0, cent).fetch(), ref)            # same centroids again (cache hit)
            cent2 = rng.choice(n, k, replace=False)                   # different centroids (cache miss)
            check_assign(ds.assign(0, cent2).fetch(), oracle.assign(data, 0, cent2))
            sub = rng.permutation(n)[: n // 3]                        # subset in arbitrary order
            check_assign(ds.assign(0, cent2, point_idx=sub).fetch(), oracle.assign(data, 0, cent2, point_idx=sub))
            ds.free()
    finally:
        c2.close()


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("n,d,k", [(700, 7, 9), (5000, 33, 300), (4097, 128, 513), (3000, 200, 130)])
def test_assign_exact_both_kernels_match_oracle(spf, oracle, metric, n, d, k):
    """The CUDA-core kernels (64 x 64 __ldg kernel; TMA-staged 8 x 8 register tile, scalar and with packed
    FADD2 differences) for every metric, candidate mode and the dense modes behind the centroid matrix /
    overflow fallback."""
    data = gauss(n, d, 40 + n)
    data[n // 2] = data[1]
    cent = np.random.default_rng(n + k).choice(n, k, replace=False)
    cent[3], cent[4] = 1, n // 2
    ref = oracle.assign(data, metric, cent)
    for mask, packed, one_cta, three_cta in ((0, 0, 0, 0), (7, 0, 0, 0), (7, 7, 0, 0), (7, 7, 7, 0), (7, 7, 0, 7)):
        c2 = spf.Context(0)
        try:
            c2.set_param("exact_tma", mask)
            c2.set_param("exact_packed", packed)
            c2.set_param("exact_one_cta", one_cta)
            c2.set_param("exact_three_cta", three_cta)
            c2.set_param("exact_tma_min_pairs", 1)
            c2.set_param("cc_cache", 0)
            ds = spf.Dataset(c2, data)
            check_assign(ds.assign(metric, cent, flags=spf.ASSIGN_FORCE_EXACT).fetch(), ref)
            ds.free()
        finally:
            c2.close()


@pytest.mark.parametrize("metric", [1, 2])
def test_assign_l1_linf_gist_dimension_matches_oracle(spf, ctx, oracle, metric):
    """Config 3's row length (d = 960) on the CUDA-core direct-form kernel, Manhattan and Chebyshev."""
    data = clustered(6000, 960, 64, 960 + metric)
    data[100] = data[7]
    cent = np.random.default_rng(960).choice(6000, 256, replace=False)
    cent[0], cent[1] = 7, 100
    ds = spf.Dataset(ctx, data)
    ref = oracle.assign(data, metric, cent)
    gdata = gauss(3000, 960, 961)                # N(0,1): the wide boundary band of BASELINE.md 4.3
    gcent = np.random.default_rng(961).choice(3000, 300, replace=False)
    ds2 = spf.Dataset(ctx, gdata)
    gref = oracle.assign(gdata, metric, gcent)
    try:
        # unseeded, and with the running minimum seeded by the tensor-core squared-L2 pre-pass
        for seed_mode in (0, 2):
            ctx.set_param("exact_seed", seed_mode)
            ctx.set_profiling(True)
            check_assign(ds.assign(metric, cent).fetch(), ref)
            assert (ctx.kernel_ms("exact_seed") > 0) == (seed_mode == 2)
            ctx.set_profiling(False)
            check_assign(ds2.assign(metric, gcent).fetch(), gref)
            sub = np.random.default_rng(5).permutation(6000)[:2500]
            check_assign(ds.assign(metric, cent, point_idx=sub).fetch(), oracle.assign(data, metric, cent, point_idx=sub))
    finally:
        ctx.set_profiling(False)
        ctx.set_param("exact_seed", 1)


@pytest.fixture(scope="module")
def config2(spf, ctx):
    """BASELINE config 2: 1M x 128 N(0,1), k = 4096 explicit centroid rows."""
    n, d, k = 1_000_000, 128, 4096
    data = np.random.Generator(np.random.Philox(key=42)).standard_normal((n, d), dtype=np.float32)
    cent = np.random.Generator(np.random.Philox(key=7)).choice(n, k, replace=False).astype(np.uint64)
    ds = spf.Dataset(ctx, data)
    return data, cent, ds


def test_config2_assign_full_size_matches_oracle_sample(spf, ctx, oracle, config2):
    """The headline configuration at its real size: the full 1M x 4096 tensor-path assign against the
    oracle on a 120k-row sample — nearest slot, distance bits, member lists restricted to the sample
    (order included) — and the same rows assigned as a point_idx subset."""
    data, cent, ds = config2
    n, k = data.shape[0], cent.size
    sample = np.sort(np.random.default_rng(11).choice(n, 120_000, replace=False)).astype(np.uint64)
    ref = oracle.assign(data, 0, cent, point_idx=sample)
    full = ds.assign(0, cent).fetch()
    assert np.array_equal(full.best[sample], ref.best)
    assert np.array_equal(full.dmin[sample].view(np.uint32), ref.dmin.view(np.uint32))
    insample = np.zeros(n, bool)
    insample[sample] = True
    mask = insample[full.members.astype(np.int64)]
    assert np.array_equal(full.members[mask], ref.members)
    cum = np.concatenate([[0], np.cumsum(mask, dtype=np.int64)])
    assert np.array_equal(cum[full.offsets.astype(np.int64)], ref.offsets.astype(np.int64))
    check_assign(ds.assign(0, cent, point_idx=sample).fetch(), ref)
    assert full.members.size >= n and int(full.offsets[k]) == full.members.size


@pytest.mark.parametrize("nprobe", [8, 32, 256])
def test_config5_query_sweep_matches_oracle_sample(spf, ctx, oracle, config2, nprobe):
    """BASELINE config 5: 100k queries, top-10 over the config-2 index on the tensor-core scan; a
    2k-query sample of every batch against the oracle (ids, distance bits, counts)."""
    data, cent, ds = config2
    if not hasattr(test_config5_query_sweep_matches_oracle_sample, "_idx"):
        res = ds.assign(0, cent)
        f = res.fetch(best=False, dmin=False)
        med = ds.update_medoids_from(0, res, cent)
        res.free()
        test_config5_query_sweep_matches_oracle_sample._idx = (spf.DeviceIndex.pack(ds, f.offsets, f.members, med), f, med)
    idx, f, med = test_config5_query_sweep_matches_oracle_sample._idx
    q = np.random.Generator(np.random.Philox(key=46)).standard_normal((100_000, 128), dtype=np.float32)
    ctx.set_profiling(True)
    ids, dists, counts = idx.search(q, 10, nprobe=nprobe)
    assert ctx.kernel_ms("scan_tc_a") > 0, "the tensor-core scan did not run"
    ctx.set_profiling(False)
    pick = np.sort(np.random.default_rng(nprobe).choice(100_000, 2000, replace=False))
    rid, rd, rc = oracle.search_batch(data, f.offsets, f.members, med, q[pick], 10, nprobe=nprobe)
    assert np.array_equal(counts[pick], rc)
    for i, qi in enumerate(pick):
        c = int(rc[i])
        assert np.array_equal(ids[qi, :c], rid[i, :c]), (nprobe, qi)
        assert np.array_equal(dists[qi, :c].view(np.uint32), rd[i, :c].view(np.uint32)), (nprobe, qi)


# ----------------------------------------------------------------------------------------------
# device-resident k-means iterations (spf_kmeans) and the NCCL group
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric,seeded", [(0, True), (0, False), (1, True), (2, True)])
def test_kmeans_session_single_gpu_matches_oracle(spf, ctx, oracle, metric, seeded):
    """One rank, no exchange: three iterations of assign_points + update_centroids stay on the device
    and reproduce the oracle bit for bit — centroid rows, vectors, means, cluster sizes, and the
    assignment itself (seeded from the second iteration on for the tensor path)."""
    n, d, k = (30000, 128, 512) if metric == 0 else (6000, 24, 40)
    data = clustered(n, d, 64, 500 + metric)
    rows = np.random.default_rng(metric + 3).choice(n, k, replace=False).astype(np.uint64)
    ds = spf.Dataset(ctx, data)
    sess = spf.KMeansSession(ds, None, metric, 0, k, seeded=seeded)
    sess.set_centroids(rows, data[rows.astype(np.int64)])
    cur = rows.copy()
    for it in range(3):
        sess.step()
        ref_a = oracle.assign(data, metric, cur)
        check_assign(sess.assignment().fetch(), ref_a)
        new_rows, ref_means = oracle.update_medoids(data, metric, ref_a.offsets, ref_a.members, cur, want_means=True)
        got_rows, got_vec, got_means, got_cnt = sess.fetch(means=True)
        assert np.array_equal(got_rows, new_rows), it
        assert np.array_equal(got_vec.view(np.uint32), data[new_rows.astype(np.int64)].view(np.uint32))
        assert np.array_equal(got_means.view(np.uint32), ref_means.view(np.uint32))
        assert np.array_equal(got_cnt, np.diff(ref_a.offsets.astype(np.int64)).astype(np.uint64))
        cur = new_rows
    sess.free()
    ds.free()


def test_kmeans_session_empty_cluster_and_duplicates(spf, ctx, oracle):
    """A centroid vector far from every point (it is not a row of this shard, as in the sharded
    build) attracts nothing: its cluster is empty and keeps its centroid (hierarchical.rs:146-149).
    Two identical centroids exercise the lowest-slot-wins tie."""
    data = gauss(5000, 16, 321)
    data[40] = data[7]
    far = np.full((1, 16), 1.0e3, np.float32)
    aug = np.concatenate([data, far])                       # the oracle sees the far vector as row 5000
    rows = np.array([7, 40, 100, 5000, 900, 2500], np.uint64)
    ds = spf.Dataset(ctx, data)
    sess = spf.KMeansSession(ds, None, 0, 0, rows.size)
    sess.set_centroids(rows, aug[rows.astype(np.int64)])
    sess.step()
    ref_a = oracle.assign(aug, 0, rows, point_idx=np.arange(5000, dtype=np.uint64))
    new_rows = oracle.update_medoids(aug, 0, ref_a.offsets, ref_a.members, rows)
    got_rows, got_vec, _, got_cnt = sess.fetch()
    assert got_cnt[3] == 0 and got_rows[3] == 5000
    assert np.array_equal(got_cnt, np.diff(ref_a.offsets.astype(np.int64)).astype(np.uint64))
    assert np.array_equal(got_rows, new_rows)
    assert np.array_equal(got_vec.view(np.uint32), aug[new_rows.astype(np.int64)].view(np.uint32))
    check_assign(sess.assignment().fetch(), ref_a)
    sess.free()
    ds.free()


def test_kmeans_session_two_gpus_nccl_matches_host_staged_shards(spf):
    """N = 2 ranks under torch.distributed.run: the device-resident NCCL iteration == the host-staged
    exchange over the same shards (tools/check_sharded_nccl.py does the comparison on rank 0)."""
    import subprocess
    import sys
    if spf._capi.lib().spf_device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29571",
                        os.path.join(root, "tools", "check_sharded_nccl.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "SHARDED_NCCL_OK" in r.stdout


def test_search_sharded_single_rank_equals_search_batch(spf, ctx, oracle):
    """comm == NULL: the sharded entry point (probe / scan phases on device buffers + device merge)
    returns exactly what spf_search_batch and the oracle return."""
    data = clustered(20000, 64, 64, 4242)
    cent = np.random.default_rng(42).choice(20000, 256, replace=False)
    ds = spf.Dataset(ctx, data)
    res = ds.assign(0, cent)
    f = res.fetch(best=False, dmin=False)
    med = ds.update_medoids_from(0, res, cent)
    res.free()
    idx = spf.DeviceIndex.pack(ds, f.offsets, f.members, med)
    q = clustered(3000, 64, 64, 4243)
    for nprobe in (4, 10, 40):
        a = idx.search_sharded(None, q, 10, nprobe=nprobe)
        b = idx.search(q, 10, nprobe=nprobe)
        assert np.array_equal(a[2], b[2]) and np.array_equal(a[0], b[0])
        assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    rid, rd, rc = oracle.search_batch(data, f.offsets, f.members, med, q[:500], 10, nprobe=10)
    a = idx.search_sharded(None, q[:500], 10, nprobe=10)
    assert np.array_equal(a[2], rc)
    for i in range(500):
        assert np.array_equal(a[0][i, :rc[i]], rid[i, :rc[i]])
    idx.free()
    ds.free()


def test_search_more_than_128_results_matches_oracle(spf, ctx, oracle):
    """The reference has no cap on k (spann_index.rs:188-193): k > 128 runs the exact query-major scan in
    passes of 128 results (each pass takes the keys behind the previous pass's last one); nprobe = 0
    means nprobe = k.  ids, distance bits and counts against the oracle, single call and sharded entry."""
    data = clustered(12000, 48, 40, 515)
    cent = np.random.default_rng(51).choice(12000, 300, replace=False)
    ds = spf.Dataset(ctx, data)
    res = ds.assign(0, cent)
    f = res.fetch(best=False, dmin=False)
    med = ds.update_medoids_from(0, res, cent)
    res.free()
    idx = spf.DeviceIndex.pack(ds, f.offsets, f.members, med)
    q = clustered(300, 48, 40, 516)
    for k, nprobe, prune in ((129, 6, 1.2), (200, 0, 1.2), (300, 20, 3.0), (1000, 12, 50.0)):
        rid, rd, rc = oracle.search_batch(data, f.offsets, f.members, med, q, k, nprobe=nprobe, prune_factor=prune)
        for got in (idx.search(q, k, nprobe=nprobe, prune_factor=prune),
                    idx.search_sharded(None, q, k, nprobe=nprobe, prune_factor=prune)):
            assert np.array_equal(got[2], rc), (k, nprobe)
            for i in range(q.shape[0]):
                assert np.array_equal(got[0][i, :rc[i]], rid[i, :rc[i]]), (k, nprobe, i)
                assert np.array_equal(got[1][i, :rc[i]].view(np.uint32), rd[i, :rc[i]].view(np.uint32))
        assert rc.max() > 128 or k == 129
    with pytest.raises(spf.SpfError):
        idx.search(q, 1025)
    idx.free()
    ds.free()


def test_search_very_long_rows_stay_on_the_query_major_scan(spf, ctx, oracle):
    """a mid-size batch of rows too long for the list-major kernel's shared-memory tiles
    (ld ~ 6500) must fall back to the query-major scan instead of failing."""
    rng = np.random.default_rng(77)
    n, d, nl = 1200, 6500, 12
    data = rng.standard_normal((n, d), dtype=np.float32)
    owner = rng.integers(0, nl, n)
    members = np.concatenate([np.flatnonzero(owner == j) for j in range(nl)]).astype(np.uint64)
    offsets = np.concatenate([[0], np.cumsum(np.bincount(owner, minlength=nl))]).astype(np.uint64)
    med = np.array([members[int(offsets[j])] for j in range(nl)], np.uint64)
    ds = spf.Dataset(ctx, data)
    idx = spf.DeviceIndex.pack(ds, offsets, members, med)
    q = rng.standard_normal((64, d), dtype=np.float32)
    rid, rd, rc = oracle.search_batch(data, offsets, members, med, q, 10, nprobe=4, prune_factor=5.0)
    ids, dists, counts = idx.search(q, 10, nprobe=4, prune_factor=5.0)
    assert np.array_equal(counts, rc)
    for i in range(q.shape[0]):
        assert np.array_equal(ids[i, :rc[i]], rid[i, :rc[i]])
        assert np.array_equal(dists[i, :rc[i]].view(np.uint32), rd[i, :rc[i]].view(np.uint32))
    idx.free()
    ds.free()


# ----------------------------------------------------------------------------------------------
# EXTENSION: balanced assignment / Lloyd iterations (no reference counterpart; the oracle's
# orc_assign_balanced is the specification — parity unpinned)
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric,n,d,k", [(0, 20000, 128, 512), (0, 6000, 96, 64), (0, 900, 10, 7), (1, 3000, 24, 40),
                                          (2, 3000, 24, 40)])
def test_assign_balanced_matches_oracle(spf, ctx, oracle, metric, n, d, k):
    """cost = fl(d + penalty[j]): nearest slot, cost bits and the one-cluster-per-point CSR equal the
    oracle extension — tensor path (penalty folded into the K extension of the GEMM) and CUDA-core
    path, explicit centroid vectors that are no dataset rows, zero and large penalties, a subset."""
    data = clustered(n, d, max(k // 4, 2), 900 + n + metric)
    rng = np.random.default_rng(n + k)
    cen = np.stack([data[rng.choice(n, 5, replace=False)].mean(0) for _ in range(k)]).astype(np.float32)
    cen[1] = cen[0]                                       # identical centroids: the lower slot wins unless penalised
    typical = float(np.median(((data[:200, None, :] - cen[None, :8, :]) ** 2).sum(-1))) if metric == 0 else 1.0
    for pen in (None, (rng.random(k) * 0.2 * typical).astype(np.float32),
                np.where(np.arange(k) % 3 == 0, np.float32(1e6), np.float32(0)).astype(np.float32)):
        ds = spf.Dataset(ctx, data)
        got = ds.assign_balanced(metric, cen, pen).fetch()
        ref = oracle.assign_balanced(data, metric, cen, pen)
        check_assign(got, ref)
        sub = rng.permutation(n)[: n // 4]
        check_assign(ds.assign_balanced(metric, cen, pen, point_idx=sub).fetch(),
                     oracle.assign_balanced(data, metric, cen, pen, point_idx=sub))
        ds.free()
    with pytest.raises(spf.SpfError):
        spf.Dataset(ctx, data).assign_balanced(metric, cen, -np.ones(k, np.float32))


@pytest.mark.parametrize("metric,lloyd", [(0, False), (0, True), (1, False)])
def test_kmeans_balanced_session_matches_oracle(spf, ctx, oracle, metric, lloyd):
    """Iterated balanced k-means on the device == the same loop on the oracle: penalty = lambda * size
    of the previous iteration, assignment by cost, medoid (or mean, Lloyd mode) update."""
    n, d, k = (30000, 128, 256) if metric == 0 else (5000, 16, 24)
    data = clustered(n, d, 16, 700 + metric)              # few true clusters: sizes are very uneven without a penalty
    rows = np.random.default_rng(5).choice(n, k, replace=False).astype(np.uint64)
    dm = float(np.median(oracle.assign(data, metric, rows, boundary_factor=1.0).dmin))
    lam = np.float32(0.5 * dm / (n / k))
    ds = spf.Dataset(ctx, data)
    sess = spf.KMeansSession(ds, None, metric, 0, k, balance_lambda=float(lam), lloyd_means=lloyd)
    sess.set_centroids(rows, data[rows.astype(np.int64)])
    vec = data[rows.astype(np.int64)].copy()
    cur_rows = rows.copy()
    counts = np.zeros(k, np.uint64)
    spread = []
    for it in range(3):
        sess.step()
        pen = (lam * counts.astype(np.float32)).astype(np.float32)
        ref = oracle.assign_balanced(data, metric, vec, pen)
        check_assign(sess.assignment().fetch(), ref)
        counts = np.diff(ref.offsets.astype(np.int64)).astype(np.uint64)
        med, means = oracle.update_medoids(data, metric, ref.offsets, ref.members, np.zeros(k, np.uint64), want_means=True)
        nz = counts > 0
        if lloyd:
            vec[nz] = means[nz]
        else:
            cur_rows = np.where(nz, med, cur_rows)
            vec[nz] = data[med[nz].astype(np.int64)]
        g_rows, g_vec, _, g_cnt = sess.fetch()
        assert np.array_equal(g_cnt, counts)
        assert np.array_equal(g_vec.view(np.uint32), vec.view(np.uint32))
        if not lloyd:
            assert np.array_equal(g_rows, cur_rows)
        spread.append(int(counts.max()))
    assert len(spread) == 3
    sess.free()
    ds.free()
