"""spfresh_b200 — B200-native (sm_100a) implementation of the SPFresh/SPANN data-parallel hot path
(batched point-to-centroid assignment, medoid update, k-means++ rounds, bisect seeds, centroid
probe + posting-list scan) behind a C ABI (include/spfresh_b200.h).

The Python layer only marshals numpy arrays into that C ABI and mirrors the reference's host
interface (HierarchicalClustering, SpannIndexBuilder, SpannIndex...).  There is no CPU fallback:
importing works without a GPU, any compute call raises.
"""
from . import _capi
from ._capi import (ASSIGN_DEFAULT, ASSIGN_FORCE_EXACT, ASSIGN_NO_CSR, METRIC_CHEBYSHEV, METRIC_EUCLIDEAN,
                    METRIC_MANHATTAN, SpfError)
from .clustering import (BOUNDARY_THRESHOLD, ChebyshevDistance, Cluster, ClusteringParams, DistanceMetric,
                         HierarchicalClustering, InitializationMethod, ManhattanDistance, NumpyRandomSource,
                         RandomSource, ScriptedRandomSource, SquaredEuclideanDistance)
from .device import (AssignResult, Context, Dataset, DeviceComm, DeviceIndex, KMeansSession, KmppSession,
                     KmppShardSession, topk_merge)
from .lire import LireError, Reassign, Split, reassign_batch
from .spann import ClusteringParamsConfig, Config, PointData, SpannIndex, SpannIndexBuilder

__all__ = [n for n in dir() if not n.startswith("_")]
