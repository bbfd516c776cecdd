"""Writes tests/golden/reference_kats.json: every known-answer vector the reference's own
tests / examples hold for the hot path.  The reference is Rust and cannot be
executed in this image, so the vectors are transcribed from its test sources; each entry
cites the file:line it was read from.  Re-run:  python tests/golden/make_reference_kats.py
"""
import json
import os

TOY = [[1.0, 2.0], [1.5, 2.5], [8.0, 8.0], [8.5, 8.5], [4.0, 4.0], [4.5, 4.5]]

KATS = {
    "source": "jairad26/spfresh (mounted at /root/reference when this file was written)",
    "distance": [
        {"cite": "src/distances/distance.rs:52-61", "metric": "Euclidean", "dtype": "f64",
         "a": [1.0, 2.0, 3.0], "b": [4.0, 5.0, 6.0], "expected": 27.0, "tol": 1e-6},
        {"cite": "src/distances/distance.rs:64-73", "metric": "Manhattan", "dtype": "f64",
         "a": [1.0, 2.0, 3.0], "b": [4.0, 5.0, 6.0], "expected": 9.0, "tol": 1e-6},
        {"cite": "src/distances/distance.rs:76-85", "metric": "Chebyshev", "dtype": "f64",
         "a": [1.0, 2.0, 3.0], "b": [4.0, 5.0, 6.0], "expected": 3.0, "tol": 1e-6},
        {"cite": "src/distances/distance.rs:88-104", "metric": "Euclidean", "dtype": "f64",
         "a": [1.0, 2.0, 3.0], "b": [1.0, 2.0, 3.0], "expected": 0.0, "tol": 1e-6},
        {"cite": "src/distances/distance.rs:88-104", "metric": "Manhattan", "dtype": "f64",
         "a": [1.0, 2.0, 3.0], "b": [1.0, 2.0, 3.0], "expected": 0.0, "tol": 1e-6},
        {"cite": "src/distances/distance.rs:88-104", "metric": "Chebyshev", "dtype": "f64",
         "a": [1.0, 2.0, 3.0], "b": [1.0, 2.0, 3.0], "expected": 0.0, "tol": 1e-6},
    ],
    "mean": [
        {"cite": "src/clustering/utils.rs:24-32", "data": [[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]],
         "indices": [0, 2], "expected": [3.0, 4.0], "tol": 1e-6},
    ],
    "toy_data": {"cite": "src/clustering/hierarchical.rs:400-402, examples/build_index.rs:9-12",
                 "rows": TOY},
    "assign": [
        # test_assign_points: for ANY pair of distinct random centroids the sizes sum to 6
        # (no boundary replication on this data) and no cluster is empty.
        {"cite": "src/clustering/hierarchical.rs:466-486", "metric": "Euclidean", "initial_k": 2,
         "init": "Random", "expect": {"sum_sizes": 6, "no_empty": True}},
    ],
    "subdivide": [
        # test_subdivide_clusters: k=1, desired 2 → more than one cluster, each <= 2 points.
        {"cite": "src/clustering/hierarchical.rs:444-463", "metric": "Euclidean", "initial_k": 1,
         "desired_cluster_size": 2, "expect": {"min_clusters": 2, "max_size": 2}},
    ],
    "fit": [
        # test_fit: KMeans++ k=3 desired 2 seed 42 → exactly 3 clusters each <= 2.  Depends on
        # rand 0.9's seed-42 stream (not reproducible here): usable only as a property over the
        # init triples for which it can hold.
        {"cite": "src/clustering/hierarchical.rs:489-507", "metric": "Euclidean", "initial_k": 3,
         "desired_cluster_size": 2, "init": "KMeansPlusPlus", "seed": 42,
         "expect": {"clusters": 3, "max_size": 2}, "rng_dependent": True},
    ],
    "example_query": [
        # examples/build_index.rs: Random init k=4, desired = round(0.18*6) = 1, query (1,2), k=1
        {"cite": "examples/build_index.rs:9-25, src/spann/spann_builder.rs:48-49",
         "metric": "Euclidean", "initial_k": 4, "desired_cluster_size": 1, "query": [1.0, 2.0],
         "k": 1, "expected_point_id": 0, "expected_vector": [1.0, 2.0]},
    ],
}

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")
    with open(out, "w") as f:
        json.dump(KATS, f, indent=1)
    print("wrote", out)
