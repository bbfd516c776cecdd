"""CPU tests (-m "not gpu"): the oracle against the reference's own known-answer vectors
(tests/golden/reference_kats.json) and against an independently written Python restatement."""
import itertools
import os

import numpy as np
import pytest

import pyref

METRIC = {"Euclidean": 0, "Manhattan": 1, "Chebyshev": 2}


def toy(kats, dtype=np.float32):
    return np.array(kats["toy_data"]["rows"], dtype)


def test_distance_kats(oracle, kats):
    for kat in kats["distance"]:
        a = np.array(kat["a"], np.float64)
        b = np.array(kat["b"], np.float64)
        m = METRIC[kat["metric"]]
        assert abs(oracle.distance(m, a, b) - kat["expected"]) < kat["tol"], kat["cite"]
        # same vectors through the f32 instantiation (what the GPU path computes in)
        assert abs(float(oracle.distance(m, a.astype(np.float32), b.astype(np.float32)))
                   - kat["expected"]) < kat["tol"]


def test_mean_kat(oracle, kats):
    for kat in kats["mean"]:
        for dt in (np.float64, np.float32):
            got = oracle.compute_mean(np.array(kat["data"], dt), kat["indices"])
            assert np.all(np.abs(got - np.array(kat["expected"])) < kat["tol"]), kat["cite"]
    # empty selection → zeros (utils.rs:10-12)
    assert np.all(oracle.compute_mean(np.ones((3, 2), np.float32), []) == 0)


def test_assign_kat_all_random_inits(oracle, kats):
    """hierarchical.rs:466-486 holds for every ordered pair of distinct centroids."""
    data = toy(kats)
    for c in itertools.permutations(range(6), 2):
        r = oracle.assign(data, 0, c)
        sizes = np.diff(r.offsets.astype(np.int64))
        assert sizes.sum() == 6 and (sizes > 0).all(), c
        assert sorted(r.members.tolist()) == list(range(6))


def test_subdivide_kat(oracle, kats):
    """hierarchical.rs:444-463: k=1, desired=2 → >1 clusters, each ≤ 2; the partition is always
    {0,1},{2,3},{4,5} whatever the random draws are."""
    data = toy(kats)
    for init in range(6):
        for pick_mode in range(3):
            pick = [lambda n: 0, lambda n: n - 1, lambda n: n // 2][pick_mode]
            cl = oracle.fit(data, 0, [init], 2, pick=pick)
            assert len(cl) > 1
            assert all(len(c.points) <= 2 for c in cl)
            assert sorted(tuple(sorted(c.points.tolist())) for c in cl) == [(0, 1), (2, 3), (4, 5)]


def test_fit_kat_structure(oracle, kats):
    """hierarchical.rs:489-507 asserts 3 clusters each ≤ 2 for rand's seed-42 draws, which can
    not be regenerated here; check the assertion holds for the init triples where it can
    (one centroid per natural pair) and that the oracle keeps the invariants otherwise."""
    data = toy(kats)
    good = 0
    for init in itertools.permutations(range(6), 3):
        cl = oracle.fit(data, 0, init, 2, pick=lambda n: 0)
        assert all(len(c.points) <= 2 for c in cl)
        covered = set()
        for c in cl:
            covered |= set(c.points.tolist())
        assert covered == set(range(6))
        if len(cl) == 3:
            good += 1
    assert good == 48   # true for 48 of the 120 ordered distinct triples


def test_example_build_index_kat(oracle, kats):
    """examples/build_index.rs:9-25: whatever the 4 random centroids are, the 6x2 toy set ends
    as 6 singleton clusters and the query (1,2), k=1 returns point_id 0."""
    data = toy(kats)
    kat = kats["example_query"][0]
    q = np.array([kat["query"]], np.float32)
    for init in itertools.permutations(range(6), 4):
        for pick in (lambda n: 0, lambda n: n - 1):
            cl = oracle.fit(data, 0, init, kat["desired_cluster_size"], pick=pick, max_splits=64)
            assert len(cl) == 6 and all(len(c.points) == 1 for c in cl)
            off, mem, rows = oracle.clusters_to_csr(cl)
            ids, dists, counts = oracle.search_batch(data, off, mem, rows, q, kat["k"])
            assert counts[0] == 1 and ids[0, 0] == kat["expected_point_id"]
            assert np.array_equal(data[ids[0, 0]], np.array(kat["expected_vector"], np.float32))


@pytest.mark.parametrize("metric", [0, 1, 2])
def test_oracle_matches_python_restatement(oracle, metric):
    rng = np.random.default_rng(100 + metric)
    n, d, k = 120, 7, 9
    data = rng.standard_normal((n, d)).astype(np.float32)
    data[17] = data[3]            # duplicates → exact ties
    data[40] = data[3]
    cent = rng.choice(n, k, replace=False)
    cent[2] = 3
    cent[5] = 17                  # two identical centroids: lowest slot must win
    r = oracle.assign(data, metric, cent)
    lists, best, dmin = pyref.assign(data, metric, range(n), cent)
    assert [l.tolist() for l in r.lists()] == lists
    assert r.best.tolist() == best
    assert np.array_equal(r.dmin, np.array(dmin, np.float32))
    # subset + order preserved
    sub = rng.permutation(n)[:50]
    r2 = oracle.assign(data, metric, cent[:3], point_idx=sub)
    l2, _, _ = pyref.assign(data, metric, sub, cent[:3])
    assert [l.tolist() for l in r2.lists()] == l2
    # medoid update
    rows = oracle.update_medoids(data, metric, r.offsets, r.members, cent)
    assert rows.tolist() == pyref.update_medoids(data, metric, lists, cent)
    # farthest
    assert oracle.farthest(data, metric, 3, np.arange(n)) == pyref.farthest(data, metric, 3, range(n))
    assert oracle.farthest(data, metric, 3, [3, 17, 40]) == 0     # all-zero distances → row 0


def test_fit_matches_python_restatement(oracle):
    rng = np.random.default_rng(7)
    data = rng.standard_normal((90, 5)).astype(np.float32)
    init = rng.choice(90, 3, replace=False)
    picks = lambda n: (n * 7) // 11  # noqa: E731
    cl = oracle.fit(data, 0, init, 12, pick=picks)
    ref = pyref.fit(data, 0, init, 12, picks)
    assert [(c.centroid_idx, c.points.tolist(), c.depth) for c in cl] == [(a, b, c) for a, b, c in ref]


def test_kmeanspp_running_min_equals_naive(oracle):
    """H5: the running-minimum restatement picks exactly what the reference's O(n k^2 d)
    recomputation (hierarchical.rs:260-276) picks."""
    rng = np.random.default_rng(5)
    data = rng.standard_normal((300, 6)).astype(np.float32)
    for metric in (0, 1, 2):
        u = rng.random(11)
        a, fa = oracle.kmeanspp(data, metric, 12, 17, u, naive=True)
        b, fb = oracle.kmeanspp(data, metric, 12, 17, u, naive=False)
        assert np.array_equal(a, b) and not fa.any() and not fb.any()
        assert a[0] == 17 and len(set(a.tolist())) == 12
    # degenerate: all points identical → all weights zero → fallback rows used
    same = np.ones((10, 3), np.float32)
    rows, fell = oracle.kmeanspp(same, 0, 3, 4, [0.5, 0.5], fallback_rows=[7, 2])
    assert rows.tolist() == [4, 7, 2] and fell.tolist() == [1, 1]


def test_search_semantics(oracle):
    """spann_index.rs:148-197 quirks: nprobe == k, 1.2x point-level filter, duplicates kept,
    None when nothing survives, fewer than k allowed."""
    rng = np.random.default_rng(11)
    data = rng.standard_normal((400, 8)).astype(np.float32)
    cent = rng.choice(400, 16, replace=False)
    r = oracle.assign(data, 0, cent)
    q = rng.standard_normal((25, 8)).astype(np.float32)
    q[0] = data[cent[3]]          # query on a centroid: d0 = 0 → thr = 1.2*eps → only exact hits
    ids, dists, counts = oracle.search_batch(data, r.offsets, r.members, cent, q, 5)
    lists = [l.tolist() for l in r.lists()]
    for i in range(25):
        ref = pyref.search(data, lists, cent, q[i], 5)
        assert counts[i] == len(ref)
        assert ids[i, :counts[i]].tolist() == [p for _, p in ref]
        assert np.array_equal(dists[i, :counts[i]], np.array([d for d, _ in ref], np.float32))
    assert counts[0] >= 1 and dists[0, 0] == 0
    assert (counts < 5).any()     # the point-level prune does return fewer than k


def test_posting_list_bincode_roundtrip(oracle, tmp_path):
    """posting_lists.rs:64-113 byte layout: u64 n | n x (u64 id, u64 d, d x f32 LE)."""
    rng = np.random.default_rng(3)
    data = rng.standard_normal((20, 4)).astype(np.float32)
    members = np.array([5, 1, 19, 5], np.uint64)
    oracle.posting_list_write(str(tmp_path), 7, data, members)
    raw = open(os.path.join(tmp_path, "posting_list_7.bin"), "rb").read()
    assert len(raw) == 8 + 4 * (8 + 8 + 4 * 4)
    assert int.from_bytes(raw[:8], "little") == 4
    assert int.from_bytes(raw[8:16], "little") == 5 and int.from_bytes(raw[16:24], "little") == 4
    assert np.array_equal(np.frombuffer(raw[24:40], "<f4"), data[5])
    ids, vec = oracle.posting_list_read(str(tmp_path), 7)
    assert np.array_equal(ids, members) and np.array_equal(vec, data[members.astype(np.int64)])
    oracle.cluster_ids_write(str(tmp_path), [7, 3])
    raw = open(os.path.join(tmp_path, "cluster_ids.bin"), "rb").read()
    assert np.array_equal(np.frombuffer(raw, "<u8"), [2, 7, 3])


def test_posting_list_reader_and_centroid_fallback_match_oracle_files(oracle, tmp_path):
    """spfresh_b200.spann.read_posting_list_file / centroids_from_lists (the loader's fallback when a
    directory has no centroids.bin) on files written by the oracle's bincode writer."""
    from spfresh_b200.spann import centroids_from_lists, read_posting_list_file
    rng = np.random.default_rng(5)
    data = rng.standard_normal((400, 6)).astype(np.float32)
    cent = rng.choice(400, 5, replace=False).astype(np.uint64)
    a = oracle.assign(data, 0, cent)
    med = oracle.update_medoids(data, 0, a.offsets, a.members, cent)
    for j in range(5):
        oracle.posting_list_write(str(tmp_path), j, data, a.members[int(a.offsets[j]):int(a.offsets[j + 1])])
    oracle.cluster_ids_write(str(tmp_path), [3, 0, 4, 1, 2])
    ids, vec = read_posting_list_file(str(tmp_path / "posting_list_3.bin"))
    want = a.members[int(a.offsets[3]):int(a.offsets[4])]
    assert np.array_equal(ids, want) and np.array_equal(vec, data[want.astype(np.int64)])
    cen = centroids_from_lists(str(tmp_path))
    assert np.array_equal(cen.view(np.uint32), data[med.astype(np.int64)].view(np.uint32))
    open(tmp_path / "posting_list_1.bin", "ab").write(b"x")
    with pytest.raises(OSError):
        read_posting_list_file(str(tmp_path / "posting_list_1.bin"))
