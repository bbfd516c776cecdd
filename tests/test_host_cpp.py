"""The C++ host layer (host/spfresh.hpp) above the C ABI: the reference's operator interface
(HierarchicalClustering::fit, SpannIndexBuilder, find_k_nearest_neighbor_spann) driven from C++
with scripted random decisions, compared with the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "host")
BIN = os.path.join(HOST, "build", "host_check")


@pytest.fixture(scope="module")
def host_bin():
    from spfresh_b200 import build as b
    b.build()                                           # the .so the programs link against
    r = subprocess.run(["make", "-C", HOST], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return BIN


def test_host_layer_builds_and_links(host_bin):
    """No GPU needed: the programs link against the C ABI and the library loads."""
    r = subprocess.run([host_bin, "--abi"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "1"
    for prog in ("build_index", "load_index"):
        assert os.access(os.path.join(HOST, "build", prog), os.X_OK)


def clustered(n, d, ncent, seed):
    g = np.random.default_rng(seed)
    cen = 2.0 * g.standard_normal((ncent, d)).astype(np.float32)
    return (cen[g.integers(0, ncent, n)] + 0.5 * g.standard_normal((n, d)).astype(np.float32)).astype(np.float32)


def run_scenario(host_bin, tmp_path, lines):
    sc, out = tmp_path / "scenario.txt", tmp_path / "out.txt"
    sc.write_text("\n".join(lines) + "\n")
    r = subprocess.run([host_bin, str(sc), str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    rows = [ln.split() for ln in out.read_text().splitlines()]
    res, i = {}, 0
    while i < len(rows):
        key = rows[i][0]
        if key == "clusters":
            c = int(rows[i][1])
            res["clusters"] = [(int(r[0]), [int(x) for x in r[3:]], int(r[1])) for r in rows[i + 1:i + 1 + c]]
            i += 1 + c
        elif key == "search":
            nq = int(rows[i][1])
            res["search"] = [[int(x) for x in r[1:]] for r in rows[i + 1:i + 1 + nq]]
            i += 1 + nq
        elif key == "labels":
            res["labels"] = np.array([int(x) for x in rows[i][2:]])
            i += 1
        else:
            res[key] = int(rows[i][1])
            i += 1
    return res


def as_tuples(clusters):
    return [(int(c.centroid_idx), np.asarray(c.points).tolist(), int(c.depth)) for c in clusters]


@pytest.mark.gpu
@pytest.mark.parametrize("metric,kind", [("Euclidean", 0), ("Manhattan", 1), ("Chebyshev", 2)])
def test_cpp_fit_and_search_match_oracle(host_bin, tmp_path, metric, kind):
    import oracle
    oracle.build()
    n, d, k0, desired = 4000, 16, 5, 300
    data = clustered(n, d, 12, 31)
    q = clustered(120, d, 12, 31)
    data.tofile(tmp_path / "data.f32")
    q.tofile(tmp_path / "q.f32")
    init = np.random.default_rng(4).choice(n, k0, replace=False).tolist()
    res = run_scenario(host_bin, tmp_path, [
        f"data {tmp_path / 'data.f32'} {n} {d}", f"metric {metric}", "init Random", f"initial_k {k0}",
        f"desired {desired}", "multiple %d %s" % (k0, " ".join(map(str, init))), "pick 5 7",
        f"queries {tmp_path / 'q.f32'} 120", "k 10", f"out_dir {tmp_path / 'idx'}"])
    ref = oracle.fit(data, kind, init, desired, pick=lambda m: (m * 5) // 7)
    assert res["clusters"] == as_tuples(ref)
    assert len(ref) > k0                                  # the bisect did fire
    assert res["labels"].shape == (n,) and res["labels"].max() < len(ref)
    # query path over the index built from those clusters (search is always squared L2, F8)
    offsets = np.zeros(len(ref) + 1, np.uint64)
    for i, c in enumerate(ref):
        offsets[i + 1] = offsets[i] + len(c.points)
    members = np.concatenate([np.asarray(c.points, np.uint64) for c in ref])
    cent = np.array([c.centroid_idx for c in ref], np.uint64)
    rid, rd, rc = oracle.search_batch(data, offsets, members, cent, q, 10, 0)
    for i in range(120):
        assert res["search"][i] == rid[i, :rc[i]].tolist(), i
    assert res["single_query_same"] == 1 and res["loaded_same"] == 1


@pytest.mark.gpu
def test_cpp_fit_kmeanspp_matches_oracle(host_bin, tmp_path):
    import oracle
    oracle.build()
    n, d = 3000, 12
    data = clustered(n, d, 9, 5)
    data.tofile(tmp_path / "data.f32")
    u = np.random.default_rng(9).random(7).tolist()
    res = run_scenario(host_bin, tmp_path, [
        f"data {tmp_path / 'data.f32'} {n} {d}", "metric Euclidean", "init KMeansPlusPlus", "initial_k 8", "desired 600",
        "first 17", "pick 0 1", "u01 7 " + " ".join(repr(x) for x in u)])
    init, _ = oracle.kmeanspp(data, 0, 8, 17, u)
    ref = oracle.fit(data, 0, init, 600, pick=lambda m: 0)
    assert res["clusters"] == as_tuples(ref)


@pytest.mark.gpu
def test_cpp_example_build_and_load_index(host_bin):
    """examples/build_index.rs + load_index.rs on the C++ layer: point_id 0 / [1.0, 2.0]."""
    r = subprocess.run([os.path.join(HOST, "build", "build_index")], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "PointData { point_id: 0, vector: [1.0, 2.0] }"
    r = subprocess.run([os.path.join(HOST, "build", "load_index")], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "Nearest neighbour: point_id:0 and vector:[1.0, 2.0]"


@pytest.mark.gpu
def test_cpp_lire_split_and_reassign(host_bin, tmp_path):
    """lire::Split / lire::Reassign on the C++ layer against a per-pair restatement."""
    import oracle
    oracle.build()
    n, d = 500, 20
    vecs = clustered(n, d, 3, 77)
    vecs[300] = vecs[0]
    vecs[450] = vecs[200]
    vecs.tofile(tmp_path / "v.f32")
    out = tmp_path / "lire.txt"
    r = subprocess.run([host_bin, "--lire", str(tmp_path / "v.f32"), str(n), str(d), str(out)], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    rows = {ln.split()[0]: [int(x) for x in ln.split()[1:]] for ln in out.read_text().splitlines()}
    d1 = np.array([oracle.distance(0, vecs[0], v) for v in vecs], np.float32)
    far = 1 + int(np.flatnonzero(d1[1:] == d1[1:].max())[-1])
    assert rows["far"] == [far]
    d2 = np.array([oracle.distance(0, vecs[far], v) for v in vecs], np.float32)
    assert rows["partition1"][1:] == [i for i in range(n) if d1[i] <= d2[i]]
    assert rows["partition2"][1:] == [i for i in range(n) if not d1[i] <= d2[i]]
    dc = np.array([oracle.distance(0, vecs[1], vecs[j]) for j in range(2, 10)], np.float32)
    assert rows["best"] == [2 + int(np.argmin(dc))]
