"""ctypes binding of the CPU oracle (oracle/spf_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under spfresh_b200/ may import this module.
The oracle restates the reference CPU path (see spf_oracle.h for the file:line map and the
pinning status).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libspf_oracle.so")

EUCLIDEAN, MANHATTAN, CHEBYSHEV = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (-O2 -ffp-contract=off).  Returns the .so path."""
    src = os.path.join(_HERE, "spf_oracle.c")
    hdr = os.path.join(_HERE, "spf_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libspf_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Assign(C.Structure):
    _fields_ = [("k", C.c_uint64), ("m", C.c_uint64), ("offsets", C.POINTER(C.c_uint64)),
                ("members", C.POINTER(C.c_uint64)), ("best", C.POINTER(C.c_uint32)),
                ("dmin", C.POINTER(C.c_float))]


class _Cluster(C.Structure):
    _fields_ = [("centroid", C.c_uint64), ("points", C.POINTER(C.c_uint64)),
                ("len", C.c_uint64), ("depth", C.c_uint64)]


class _Clusters(C.Structure):
    _fields_ = [("c", C.POINTER(_Cluster)), ("count", C.c_size_t), ("cap", C.c_size_t)]


PICK_FN = C.CFUNCTYPE(C.c_uint64, C.c_void_p, C.c_uint64)

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    fp, dp, up = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint64)
    L.orc_distance_f32.restype = C.c_float
    L.orc_distance_f32.argtypes = [C.c_int, fp, fp, C.c_size_t]
    L.orc_distance_f64.restype = C.c_double
    L.orc_distance_f64.argtypes = [C.c_int, dp, dp, C.c_size_t]
    L.orc_compute_mean_f32.restype = None
    L.orc_compute_mean_f32.argtypes = [fp, C.c_size_t, up, C.c_size_t, fp]
    L.orc_compute_mean_f64.restype = None
    L.orc_compute_mean_f64.argtypes = [dp, C.c_size_t, up, C.c_size_t, dp]
    L.orc_assign.restype = C.c_int
    L.orc_assign.argtypes = [fp, C.c_size_t, C.c_size_t, C.c_int, up, C.c_size_t, up, C.c_size_t,
                             C.c_float, C.c_int, C.POINTER(_Assign)]
    L.orc_assign_free.argtypes = [C.POINTER(_Assign)]
    L.orc_update_medoids.restype = C.c_int
    L.orc_update_medoids.argtypes = [fp, C.c_size_t, C.c_size_t, C.c_int, up, up, C.c_size_t, up, up,
                                     fp, C.c_int]
    L.orc_kmeanspp.restype = C.c_int
    L.orc_kmeanspp.argtypes = [fp, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_uint64, dp, up,
                               C.c_int, C.c_int, up, C.POINTER(C.c_uint8)]
    L.orc_farthest.restype = C.c_uint64
    L.orc_farthest.argtypes = [fp, C.c_size_t, C.c_int, C.c_uint64, up, C.c_size_t]
    L.orc_fit.restype = C.c_int
    L.orc_fit.argtypes = [fp, C.c_size_t, C.c_size_t, C.c_int, up, C.c_size_t, C.c_uint64, PICK_FN,
                          C.c_void_p, C.c_uint64, C.c_int, C.POINTER(_Clusters)]
    L.orc_clusters_free.argtypes = [C.POINTER(_Clusters)]
    L.orc_search_batch.restype = C.c_int
    L.orc_search_batch.argtypes = [fp, C.c_size_t, up, up, up, C.c_size_t, fp, C.c_size_t, C.c_size_t,
                                   C.c_size_t, C.c_float, C.c_int, up, fp, C.POINTER(C.c_uint32)]
    L.orc_posting_list_write.restype = C.c_int
    L.orc_posting_list_write.argtypes = [C.c_char_p, C.c_uint64, fp, C.c_size_t, up, C.c_size_t]
    L.orc_cluster_ids_write.restype = C.c_int
    L.orc_cluster_ids_write.argtypes = [C.c_char_p, up, C.c_size_t]
    L.orc_posting_list_read.restype = C.c_int
    L.orc_posting_list_read.argtypes = [C.c_char_p, C.c_uint64, up, up, C.POINTER(up), C.POINTER(fp)]
    L.orc_assign_balanced.restype = C.c_int
    L.orc_assign_balanced.argtypes = [fp, C.c_size_t, C.c_int, up, C.c_size_t, fp, fp, C.c_size_t, C.c_int,
                                      C.POINTER(C.c_uint32), fp]
    L.orc_free.argtypes = [C.c_void_p]
    L.orc_online_cpus.restype = C.c_int
    _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


def online_cpus() -> int:
    return int(lib().orc_online_cpus())


def distance(metric: int, a, b):
    a = np.asarray(a)
    if a.dtype == np.float64:
        a = np.ascontiguousarray(a, np.float64)
        b = np.ascontiguousarray(b, np.float64)
        return float(lib().orc_distance_f64(metric, _p(a, C.c_double), _p(b, C.c_double), a.size))
    a, b = _f32(a), _f32(b)
    return np.float32(lib().orc_distance_f32(metric, _p(a, C.c_float), _p(b, C.c_float), a.size))


def compute_mean(data, idx):
    data = np.asarray(data)
    idx = _u64(idx)
    if data.dtype == np.float64:
        data = np.ascontiguousarray(data, np.float64)
        out = np.zeros(data.shape[1], np.float64)
        lib().orc_compute_mean_f64(_p(data, C.c_double), data.shape[1], _p(idx, C.c_uint64), idx.size,
                                   _p(out, C.c_double))
        return out
    data = _f32(data)
    out = np.zeros(data.shape[1], np.float32)
    lib().orc_compute_mean_f32(_p(data, C.c_float), data.shape[1], _p(idx, C.c_uint64), idx.size,
                               _p(out, C.c_float))
    return out


@dataclass
class AssignResult:
    offsets: np.ndarray   # (k+1,) uint64
    members: np.ndarray   # (offsets[k],) uint64 dataset rows, cluster-major, input order
    best: np.ndarray      # (m,) uint32
    dmin: np.ndarray      # (m,) float32

    def lists(self):
        return [self.members[self.offsets[j]:self.offsets[j + 1]] for j in range(len(self.offsets) - 1)]


def assign(data, metric, centroid_rows, point_idx=None, boundary_factor=1.1, threads=0) -> AssignResult:
    data = _f32(data)
    n, d = data.shape
    cr = _u64(centroid_rows)
    if point_idx is None:
        pi, pip, m = None, None, n
    else:
        pi = _u64(point_idx)
        pip, m = _p(pi, C.c_uint64), pi.size
    res = _Assign()
    rc = lib().orc_assign(_p(data, C.c_float), n, d, metric, pip, m, _p(cr, C.c_uint64), cr.size,
                          np.float32(boundary_factor), threads, C.byref(res))
    if rc != 0:
        lib().orc_assign_free(C.byref(res))
        raise RuntimeError(f"orc_assign rc={rc}")
    k = cr.size
    offsets = np.ctypeslib.as_array(res.offsets, (k + 1,)).copy()
    total = int(offsets[k])
    members = np.ctypeslib.as_array(res.members, (max(total, 1),))[:total].copy()
    best = np.ctypeslib.as_array(res.best, (max(m, 1),))[:m].copy()
    dmin = np.ctypeslib.as_array(res.dmin, (max(m, 1),))[:m].copy()
    lib().orc_assign_free(C.byref(res))
    return AssignResult(offsets, members, best, dmin)


def assign_balanced(data, metric, centroids, penalty=None, point_idx=None, threads=0) -> AssignResult:
    """EXTENSION (parity unpinned — the reference has no balanced assignment): cost = fl(d + penalty[j]),
    argmin with the reference's fold, every point in exactly its best cluster (CSR in input order)."""
    data = _f32(data)
    n, d = data.shape
    cen = _f32(centroids).reshape(-1, d)
    k = cen.shape[0]
    pen = None if penalty is None else _f32(penalty).reshape(k)
    if point_idx is None:
        pi, pip, m = None, None, n
    else:
        pi = _u64(point_idx)
        pip, m = _p(pi, C.c_uint64), pi.size
    best = np.zeros(m, np.uint32)
    cost = np.zeros(m, np.float32)
    rc = lib().orc_assign_balanced(_p(data, C.c_float), d, metric, pip, m, _p(cen, C.c_float),
                                   None if pen is None else _p(pen, C.c_float), k, threads,
                                   _p(best, C.c_uint32), _p(cost, C.c_float))
    if rc != 0:
        raise RuntimeError(f"orc_assign_balanced rc={rc}")
    order = np.argsort(best, kind="stable")
    rows = np.arange(m, dtype=np.uint64) if pi is None else pi
    members = rows[order]
    offsets = np.concatenate([[0], np.cumsum(np.bincount(best, minlength=k))]).astype(np.uint64)
    return AssignResult(offsets, members, best, cost)


def update_medoids(data, metric, offsets, members, old_rows, want_means=False, threads=0):
    data = _f32(data)
    n, d = data.shape
    offsets, members, old_rows = _u64(offsets), _u64(members), _u64(old_rows)
    k = old_rows.size
    new_rows = np.zeros(k, np.uint64)
    means = np.zeros((k, d), np.float32) if want_means else None
    lib().orc_update_medoids(_p(data, C.c_float), n, d, metric, _p(offsets, C.c_uint64),
                             _p(members, C.c_uint64), k, _p(old_rows, C.c_uint64),
                             _p(new_rows, C.c_uint64),
                             _p(means, C.c_float) if want_means else None, threads)
    return (new_rows, means) if want_means else new_rows


def kmeanspp(data, metric, k, first_row, u01, fallback_rows=None, naive=False, threads=0):
    data = _f32(data)
    n, d = data.shape
    u01 = np.ascontiguousarray(u01, np.float64)
    fb = _u64(fallback_rows if fallback_rows is not None else np.zeros(max(k - 1, 1)))
    out = np.zeros(k, np.uint64)
    fell = np.zeros(max(k - 1, 1), np.uint8)
    rc = lib().orc_kmeanspp(_p(data, C.c_float), n, d, metric, k, int(first_row), _p(u01, C.c_double),
                            _p(fb, C.c_uint64), int(naive), threads, _p(out, C.c_uint64),
                            _p(fell, C.c_uint8))
    if rc != 0:
        raise RuntimeError(f"orc_kmeanspp rc={rc}")
    return out, fell[:max(k - 1, 0)]


def farthest(data, metric, c1, members):
    data = _f32(data)
    members = _u64(members)
    return int(lib().orc_farthest(_p(data, C.c_float), data.shape[1], metric, int(c1),
                                  _p(members, C.c_uint64), members.size))


@dataclass
class Cluster:
    """hierarchical.rs:26-30"""
    centroid_idx: int
    points: np.ndarray
    depth: int


def fit(data, metric, init_rows, desired_cluster_size, pick=None, max_splits=100000, threads=0):
    """hierarchical.rs:65-71 given the initial centroid rows; pick(len)->index stands in for
    points.choose(&mut rng) in create_subclusters."""
    data = _f32(data)
    n, d = data.shape
    init_rows = _u64(init_rows)
    if pick is None:
        pick = lambda ln: 0  # noqa: E731
    cb = PICK_FN(lambda ctx, ln: int(pick(int(ln))))
    cl = _Clusters()
    rc = lib().orc_fit(_p(data, C.c_float), n, d, metric, _p(init_rows, C.c_uint64), init_rows.size,
                       int(desired_cluster_size), cb, None, int(max_splits), threads, C.byref(cl))
    out = []
    for i in range(cl.count):
        c = cl.c[i]
        pts = np.ctypeslib.as_array(c.points, (max(int(c.len), 1),))[:int(c.len)].copy()
        out.append(Cluster(int(c.centroid), pts, int(c.depth)))
    lib().orc_clusters_free(C.byref(cl))
    if rc < 0:
        raise RuntimeError(f"orc_fit rc={rc}")
    return out


def clusters_to_csr(clusters):
    offsets = np.zeros(len(clusters) + 1, np.uint64)
    for i, c in enumerate(clusters):
        offsets[i + 1] = offsets[i] + np.uint64(len(c.points))
    members = (np.concatenate([np.asarray(c.points, np.uint64) for c in clusters])
               if clusters else np.zeros(0, np.uint64))
    rows = np.array([c.centroid_idx for c in clusters], np.uint64)
    return offsets, members.astype(np.uint64), rows


def search_batch(data, offsets, members, centroid_rows, queries, k, nprobe=0, prune_factor=1.2, threads=0):
    """spann_index.rs:148-197 for each query.  Returns ids (nq,k) uint64, dists (nq,k) f32,
    counts (nq,) uint32 (0 == None)."""
    data = _f32(data)
    d = data.shape[1]
    offsets, members, cr = _u64(offsets), _u64(members), _u64(centroid_rows)
    q = _f32(queries).reshape(-1, d)
    nq = q.shape[0]
    ids = np.zeros((nq, k), np.uint64)
    dists = np.zeros((nq, k), np.float32)
    counts = np.zeros(nq, np.uint32)
    lib().orc_search_batch(_p(data, C.c_float), d, _p(offsets, C.c_uint64), _p(members, C.c_uint64),
                           _p(cr, C.c_uint64), cr.size, _p(q, C.c_float), nq, k, nprobe,
                           np.float32(prune_factor), threads, _p(ids, C.c_uint64),
                           _p(dists, C.c_float), _p(counts, C.c_uint32))
    return ids, dists, counts


def posting_list_write(directory, cluster_id, data, members):
    data = _f32(data)
    members = _u64(members)
    rc = lib().orc_posting_list_write(os.fsencode(directory), int(cluster_id), _p(data, C.c_float),
                                      data.shape[1], _p(members, C.c_uint64), members.size)
    if rc:
        raise OSError(f"orc_posting_list_write rc={rc}")


def cluster_ids_write(directory, ids):
    ids = _u64(ids)
    rc = lib().orc_cluster_ids_write(os.fsencode(directory), _p(ids, C.c_uint64), ids.size)
    if rc:
        raise OSError(f"orc_cluster_ids_write rc={rc}")


def posting_list_read(directory, cluster_id):
    ln, d = C.c_uint64(), C.c_uint64()
    ids = C.POINTER(C.c_uint64)()
    vec = C.POINTER(C.c_float)()
    rc = lib().orc_posting_list_read(os.fsencode(directory), int(cluster_id), C.byref(ln), C.byref(d),
                                     C.byref(ids), C.byref(vec))
    if rc:
        raise OSError(f"orc_posting_list_read rc={rc}")
    n, dd = int(ln.value), int(d.value)
    out_ids = np.ctypeslib.as_array(ids, (max(n, 1),))[:n].copy()
    out_vec = (np.ctypeslib.as_array(vec, (max(n * dd, 1),))[:n * dd].copy().reshape(n, dd)
               if n else np.zeros((0, dd), np.float32))
    lib().orc_free(ids)
    if vec:
        lib().orc_free(vec)
    return out_ids, out_vec
