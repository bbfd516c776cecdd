"""CPU tests (-m "not gpu"): the C-ABI library loads, exports every symbol include/spfresh_b200.h
declares, fails loudly without a device, and the host-side logic (config, merge) works."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from spfresh_b200 import build as b
    b.build()
    from spfresh_b200 import _capi
    return _capi


def declared_functions():
    src = open(os.path.join(ROOT, "include", "spfresh_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spf_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(capi):
    names = declared_functions()
    assert len(names) >= 40
    L = capi.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
        assert n in capi.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(capi.SIGNATURES) == names
    assert L.spf_abi_version() == 1


def test_rust_ffi_matches_header_and_ctypes(capi):
    """spann-cuda-sys/src/ffi.rs is generated from the header (tools/gen_rust_ffi.py): it must be up to
    date, declare exactly the exported functions, and agree with the ctypes table on arity."""
    import subprocess
    import sys
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_rust_ffi.py"), "--check"])
    rs = open(os.path.join(ROOT, "spann-cuda-sys", "src", "ffi.rs")).read()
    fns = dict(re.findall(r"pub fn (spf_\w+)\(([^)]*)\)", rs))
    assert sorted(fns) == declared_functions()
    for name, args in fns.items():
        arity = 0 if not args.strip() else args.count(":")
        assert arity == len(capi.SIGNATURES[name][1]), name
    for f in ("Cargo.toml", "build.rs", "src/lib.rs", "integration/hierarchical_gpu.rs", "integration/spann_index_gpu.rs",
              "integration/distance_kind.rs"):
        assert os.path.exists(os.path.join(ROOT, "spann-cuda-sys", f)), f


def test_header_compiles_as_c(tmp_path):
    import subprocess
    c = tmp_path / "t.c"
    c.write_text('#include "spfresh_b200.h"\nint main(void){return SPF_OK;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", str(c), "-o", str(tmp_path / "t.o")])


def test_no_cpu_fallback(capi):
    import spfresh_b200 as s
    if capi.lib().spf_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(s.SpfError) as e:
        s.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """The product path must not import / link / call anything under oracle/."""
    pkg = os.path.join(ROOT, "spfresh_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "spf_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f


def test_topk_merge_host(capi):
    import spfresh_b200 as s
    rng = np.random.default_rng(0)
    parts, nq, k = 3, 50, 5
    # global candidate pool per query, split across parts; keys unique
    keys = np.full((parts, nq, k), np.iinfo(np.uint64).max, np.uint64)
    ids = np.zeros((parts, nq, k), np.uint64)
    dists = np.full((parts, nq, k), np.inf, np.float32)
    counts = np.zeros((parts, nq), np.uint32)
    expect = []
    for q in range(nq):
        pool = []
        for p in range(parts):
            c = int(rng.integers(0, k + 1))
            d = np.sort(rng.random(c).astype(np.float32))
            seq = rng.choice(1000, c, replace=False)
            kk = sorted((int(np.float32(x).view(np.uint32)) << 32) | int(sq) for x, sq in zip(d, seq))
            counts[p, q] = c
            for i, key in enumerate(kk):
                keys[p, q, i] = key
                ids[p, q, i] = key & 0xffff
                dists[p, q, i] = np.uint32(key >> 32).view(np.float32)
            pool += kk
        expect.append(sorted(pool)[:k])
    o_ids, o_d, o_c = s.topk_merge(keys, ids, dists, counts)
    for q in range(nq):
        assert o_c[q] == len(expect[q])
        assert o_ids[q, :o_c[q]].tolist() == [e & 0xffff for e in expect[q]]
        assert np.all(np.isinf(o_d[q, o_c[q]:]))


def test_config_mirror(tmp_path):
    """config.rs:59-113 / F11: 'KMeansPlusPlus' is accepted, the README's 'KMeans++' is rejected."""
    import spfresh_b200 as s
    p = tmp_path / "c.yaml"
    p.write_text('clustering_params:\n  distance_metric: "Euclidean"\n  initialization_method: "Random"\n'
                 '  initial_k: 4\noutput_path: "data"\n')
    cfg = s.Config.from_file(str(p))
    assert cfg.clustering_params.initial_k == 4 and cfg.output_path == "data"
    cp = cfg.to_clustering_params()
    assert cp.distance_metric.kind == s.METRIC_EUCLIDEAN and cp.rng_seed is None and cp.desired_cluster_size is None
    for bad in ('"KMeans++"', '"kmeans"'):
        p.write_text(f'clustering_params:\n  distance_metric: "Euclidean"\n  initialization_method: {bad}\n  initial_k: 4\n')
        with pytest.raises(ValueError):
            s.Config.from_file(str(p))
    p.write_text('clustering_params:\n  distance_metric: "Cosine"\n  initialization_method: "Random"\n  initial_k: 4\n')
    with pytest.raises(ValueError):
        s.Config.from_file(str(p))
    p.write_text('clustering_params:\n  distance_metric: "Euclidean"\n  initialization_method: "Random"\n  initial_k: 0\n')
    with pytest.raises(ValueError):
        s.Config.from_file(str(p))


def test_lire_bookkeeping_needs_no_device():
    """validate() / get_affected_partitions() of the LIRE mirror are host logic (operations.rs:103-120)."""
    import spfresh_b200 as s
    vecs = [(3, [0.0, 1.0]), (9, [2.0, 2.0])]
    op = s.Split(7, vecs, s.SquaredEuclideanDistance(), (8, 9))
    assert op.validate() and op.get_affected_partitions() == {7, 8, 9}
    assert not s.Split(7, vecs, s.SquaredEuclideanDistance(), (8, 8)).validate()
    assert not s.Split(7, vecs[:1], s.SquaredEuclideanDistance(), (8, 9)).validate()
    with pytest.raises(s.LireError):
        s.Split(7, vecs[:1], s.SquaredEuclideanDistance(), (8, 9)).select_initial_centroids()
    with pytest.raises(s.LireError):
        s.Reassign(1, [0.0, 1.0], 5, [], s.SquaredEuclideanDistance(), 1).find_best_posting()
