"""Config 2 end to end on one B200: k-means++ init (k = 4096) -> assign -> update_centroids -> assign
-> posting lists in HBM -> 10k-query top-10, through the host mirror of the reference interface.
usage: python tools/build_prof.py [k] [kind]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
import spfresh_b200 as s  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else bench.K_CENT
kind = sys.argv[2] if len(sys.argv) > 2 else "gauss"
rows = bench.make_rows(0)
if kind == "clustered":
    g = np.random.Generator(np.random.Philox(key=44))
    cen = 2.0 * g.standard_normal((1024, bench.DIM), dtype=np.float32)
    rows = (cen[g.integers(0, 1024, rows.shape[0])] + 0.5 * rows).astype(np.float32)
ctx = s.Context(0)
t0 = time.perf_counter()
ds = s.Dataset(ctx, rows)
t_up = time.perf_counter() - t0
u = np.random.Generator(np.random.Philox(key=9)).random(k)
params = s.ClusteringParams(s.SquaredEuclideanDistance(), s.InitializationMethod.KMeansPlusPlus,
                            int(round(rows.shape[0] * 0.18)), k,
                            random_source=s.ScriptedRandomSource(index=lambda n: n // 3, u01=u.tolist()))
hc = s.HierarchicalClustering(params, rows, ctx=ctx, dataset=ds)
t = {}
t0 = time.perf_counter(); hc.initialize_clusters(k); t["kmeanspp_init_s"] = time.perf_counter() - t0
t0 = time.perf_counter(); hc.assign_points(); t["assign_points_s"] = time.perf_counter() - t0
t0 = time.perf_counter(); hc.update_centroids(); t["update_centroids_s"] = time.perf_counter() - t0
t0 = time.perf_counter(); hc.subdivide_clusters(); t["subdivide_clusters_s"] = time.perf_counter() - t0
index = s.SpannIndex("/tmp/spf_build_prof", ctx)
t0 = time.perf_counter(); index.create_posting_lists(ds, hc.clusters); t["create_posting_lists_s"] = time.perf_counter() - t0
q = bench.make_queries(10000)
index.device_index.search(q, 10)
t0 = time.perf_counter(); index.device_index.search(q, 10); t["query_10k_top10_s"] = time.perf_counter() - t0
t["upload_s"] = t_up
t["fit_total_s"] = sum(t[x] for x in ("kmeanspp_init_s", "assign_points_s", "update_centroids_s", "subdivide_clusters_s"))
print(json.dumps({"config": "build", "data": kind, "n": int(rows.shape[0]), "d": bench.DIM, "k": k,
                  "clusters": len(hc.clusters), "members": int(sum(len(c.points) for c in hc.clusters)), **t}))
