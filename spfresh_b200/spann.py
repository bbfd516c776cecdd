"""Host-side mirror of src/spann/{config,spann_builder,spann_index,posting_lists}.rs for the hot
path: same names and semantics, posting lists resident in HBM, queries batched."""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import yaml

from .clustering import (ChebyshevDistance, ClusteringParams, HierarchicalClustering, InitializationMethod,
                         ManhattanDistance, RandomSource, SquaredEuclideanDistance)
from .device import Context, Dataset, DeviceIndex


@dataclass
class PointData:                                  # posting_lists.rs:7-11
    point_id: int
    vector: List[float]


@dataclass
class ClusteringParamsConfig:                     # config.rs:7-12
    distance_metric: str
    initialization_method: str
    initial_k: int


@dataclass
class Config:                                     # config.rs:14-19
    clustering_params: ClusteringParamsConfig
    data_file: Optional[str] = None
    output_path: Optional[str] = None

    @staticmethod
    def from_file(file_path: str) -> "Config":    # config.rs:52-57
        with open(file_path) as f:
            raw = yaml.safe_load(f)
        cp = raw["clustering_params"]
        cfg = Config(ClusteringParamsConfig(str(cp["distance_metric"]), str(cp["initialization_method"]),
                                            int(cp["initial_k"])),
                     raw.get("data_file"), raw.get("output_path"))
        cfg.validate()
        return cfg

    def validate(self):                           # config.rs:59-87
        if self.clustering_params.distance_metric not in ("Euclidean", "Manhattan", "Chebyshev"):
            raise ValueError(f"Unsupported distance metric: {self.clustering_params.distance_metric}")
        if self.clustering_params.initialization_method not in ("Random", "KMeansPlusPlus"):
            raise ValueError(f"Unsupported initialization method: {self.clustering_params.initialization_method}")
        if self.clustering_params.initial_k == 0:
            raise ValueError("initial_k must be greater than 0")

    def to_clustering_params(self) -> ClusteringParams:   # config.rs:90-113
        metric = {"Euclidean": SquaredEuclideanDistance, "Manhattan": ManhattanDistance,
                  "Chebyshev": ChebyshevDistance}[self.clustering_params.distance_metric]()
        init = InitializationMethod(self.clustering_params.initialization_method)
        return ClusteringParams(metric, init, None, self.clustering_params.initial_k, None)


def read_posting_list_file(path: str):
    """posting_list_{id}.bin = bincode 1.x of Vec<PointData{point_id: usize, vector: Vec<f32>}>
    (posting_lists.rs:7-11, 64-90): u64 n, then n x (u64 id, u64 d, d x f32), little endian."""
    raw = np.fromfile(path, np.uint8)
    if raw.size < 8:
        raise OSError(f"{path} is truncated")
    n = int(raw[:8].view("<u8")[0])
    if n == 0:
        return np.zeros(0, np.uint64), np.zeros((0, 0), np.float32)
    if raw.size < 24:
        raise OSError(f"{path} is truncated")
    d = int(raw[16:24].view("<u8")[0])
    rec = 16 + 4 * d
    if raw.size != 8 + n * rec:
        raise OSError(f"{path}: {raw.size} bytes, expected {8 + n * rec} for {n} vectors of {d} floats")
    body = raw[8:].reshape(n, rec)
    ids = np.ascontiguousarray(body[:, :8]).view("<u8").reshape(n)
    vec = np.ascontiguousarray(body[:, 16:]).view("<f4").reshape(n, d)
    return ids.astype(np.uint64), vec.astype(np.float32)


def centroids_from_lists(path: str) -> np.ndarray:
    """Fallback navigator for a directory without centroids.bin: per list the medoid of its vectors
    (mean by row-by-row f32 sum, utils.rs:13-14; nearest member by squared L2, strict <, leftmost,
    hierarchical.rs:155-171).  Lists are numbered by cluster_ids.bin."""
    raw = np.fromfile(os.path.join(path, "cluster_ids.bin"), "<u8")
    ids = np.sort(raw[1:1 + int(raw[0])])
    nlists = int(ids.max()) + 1 if ids.size else 0
    cen = None
    for l in ids:
        _, vec = read_posting_list_file(os.path.join(path, f"posting_list_{int(l)}.bin"))
        if vec.shape[0] == 0:
            continue
        if cen is None:
            cen = np.zeros((nlists, vec.shape[1]), np.float32)
        acc = np.zeros(vec.shape[1], np.float32)
        for row in vec:                                   # sequential f32 row sum
            acc = (acc + row).astype(np.float32)
        mean = (acc / np.float32(vec.shape[0])).astype(np.float32)
        diff = (vec - mean).astype(np.float32)
        dist = np.zeros(vec.shape[0], np.float32)
        for j in range(vec.shape[1]):                     # sequential un-fused f32 chain per member
            dist = (dist + (diff[:, j] * diff[:, j]).astype(np.float32)).astype(np.float32)
        cen[int(l)] = vec[int(np.argmin(dist))]           # argmin returns the first minimum
    if cen is None:
        raise OSError(f"{path} holds no posting lists")
    return cen


class SpannIndex:
    """spann_index.rs:17-197.  The kd-tree over centroids is replaced by an exact batched probe
    (same result: exact k-NN by squared L2, ascending) and the per-cluster files by lists in HBM."""

    def __init__(self, posting_lists_dir: str, ctx: Optional[Context] = None):
        self.posting_list_dir = posting_lists_dir
        self.ctx = ctx or Context.default()
        self.device_index: Optional[DeviceIndex] = None
        self.centroids: Optional[np.ndarray] = None   # dense nlists x d (stands in for the kd-tree)

    def create_posting_lists(self, dataset: Dataset, clusters, list_range=None):   # spann_index.rs:56-114
        offsets = np.zeros(len(clusters) + 1, np.uint64)
        for i, c in enumerate(clusters):
            offsets[i + 1] = offsets[i] + np.uint64(len(c.points))
        members = (np.concatenate([np.asarray(c.points, np.uint64) for c in clusters])
                   if clusters else np.zeros(0, np.uint64))
        rows = np.array([c.centroid_idx for c in clusters], np.uint64)
        self.device_index = DeviceIndex.pack(dataset, offsets, members, rows, list_range)
        return self.device_index

    def save_posting_list(self):                  # spann_index.rs:45-53 → posting_lists.rs:108-113
        if self.device_index is None:
            raise RuntimeError("Posting list is not available")
        self.device_index.save_dir(self.posting_list_dir)

    def save_centroids(self, path: str):
        """Sidecar for the dense centroid matrix (the reference keeps centroids only inside
        output.kdtree, kiddo's private layout — SURVEY.md §8(f) rank 2)."""
        c = np.ascontiguousarray(self.centroids, np.float32)
        with open(path, "wb") as f:
            f.write(np.array(c.shape, "<u8").tobytes())
            f.write(c.astype("<f4").tobytes())

    def load_posting_list(self, path: str, centroids_path: Optional[str] = None, centroids=None,
                          recompute_centroids: bool = False):                          # spann_index.rs:32-43
        """Loads posting_list_{id}.bin + cluster_ids.bin (the reference's layout, both directions
        compatible) and the dense centroid matrix the GPU probe needs.  The reference keeps the
        centroids only inside output.kdtree (gzip + bincode of kiddo's private tree layout, which
        is not read here), so the matrix comes from, in this order:
          1. `centroids` (nlists x d, list-id order) given by the caller;
          2. the `centroids.bin` sidecar this builder writes next to the lists;
          3. `recompute_centroids=True`: per list the member nearest (squared L2, leftmost on ties)
             to the list's mean, i.e. what update_centroids (hierarchical.rs:138-181) produced for
             a Euclidean build whose clusters were not bisected afterwards.  For other builds this
             is only an approximation of the reference's navigator: results can differ from it.
        A directory written by the reference alone therefore needs 1. or 3.; a directory written
        here lacks output.kdtree, so the reference's own load() cannot query it (one-way format for
        the navigator, two-way for the lists)."""
        cpath = centroids_path or os.path.join(path, "centroids.bin")
        if centroids is not None:
            cen = np.ascontiguousarray(centroids, np.float32)
        elif os.path.exists(cpath):
            with open(cpath, "rb") as f:
                shape = np.frombuffer(f.read(16), "<u8")
                cen = np.frombuffer(f.read(), "<f4").reshape(int(shape[0]), int(shape[1]))
        elif recompute_centroids:
            cen = centroids_from_lists(path)
        else:
            raise FileNotFoundError(
                f"{cpath} not found: the reference stores centroids only inside output.kdtree (kiddo's private layout). "
                "Pass centroids=<nlists x d array>, or recompute_centroids=True to derive them from the lists "
                "(exact for Euclidean builds without bisected clusters), or write the sidecar with "
                "SpannIndex.save_centroids().")
        self.centroids = cen
        self.device_index = DeviceIndex.load_dir(self.ctx, path, cen)

    def find_k_nearest_neighbors_batch(self, queries, k: int, nprobe: int = 0, prune_factor: float = 1.2):
        """Batched sibling of find_k_nearest_neighbor_spann: list (per query) of Optional[List[PointData]]."""
        if self.device_index is None:
            raise RuntimeError("Posting list is not available")   # .expect() in the reference
        ids, dists, counts, vec = self.device_index.search(queries, k, nprobe, prune_factor, want_vectors=True)
        out = []
        for q in range(ids.shape[0]):
            n = int(counts[q])
            out.append(None if n == 0 else
                       [PointData(int(ids[q, i]), vec[q, i].tolist()) for i in range(n)])
        return out

    def find_k_nearest_neighbor_spann(self, query, k: int):   # spann_index.rs:148-197
        query = np.asarray(query, np.float32)
        if query.ndim != 1 or query.shape[0] != self.device_index.d:
            raise ValueError("Query length mismatch")          # .expect() in the reference
        return self.find_k_nearest_neighbors_batch(query[None, :], k)[0]


class SpannIndexBuilder:
    """spann_builder.rs:8-75."""

    def __init__(self, config: Config, ctx: Optional[Context] = None,
                 random_source: Optional[RandomSource] = None):
        self.config = config
        self.data = None
        self.ctx = ctx
        self.random_source = random_source

    def with_data(self, data) -> "SpannIndexBuilder":
        self.data = np.asarray(data, np.float32)
        return self

    def build(self, N: Optional[int] = None) -> SpannIndex:   # spann_builder.rs:25-64
        if self.data is None:
            raise ValueError("No data provided (in-memory or file)")
        if N is not None and self.data.shape[1] != N:
            raise ValueError(f"Data dimension mismatch: expected {N}, got {self.data.shape[1]}")
        params = self.config.to_clustering_params()
        params.desired_cluster_size = int(math.floor(self.data.shape[0] * 0.18 + 0.5))   # :48-49 f64::round
        params.random_source = self.random_source
        ctx = self.ctx or Context.default()
        clustering = HierarchicalClustering(params, self.data, ctx=ctx)
        clustering.fit()
        if self.config.output_path is None:
            raise ValueError("Output path is not specified")
        index = SpannIndex(self.config.output_path, ctx)
        index.create_posting_lists(clustering.dataset, clustering.clusters)
        index.centroids = self.data[[c.centroid_idx for c in clustering.clusters]]
        index.clusters = clustering.clusters
        try:                                       # `let _ =` in the reference: errors are dropped
            index.save_posting_list()
            index.save_centroids(os.path.join(self.config.output_path, "centroids.bin"))
        except Exception:
            pass
        return index

    def load(self, N: Optional[int] = None, centroids=None, recompute_centroids: bool = False) -> SpannIndex:
        """spann_builder.rs:66-75.  See SpannIndex.load_posting_list for where the centroid matrix comes
        from (the sidecar written by build(), the caller, or the lists themselves)."""
        if self.config.output_path is None:
            raise ValueError("Output path is not specified")
        index = SpannIndex(self.config.output_path, self.ctx or Context.default())
        index.load_posting_list(self.config.output_path, centroids=centroids, recompute_centroids=recompute_centroids)
        return index
