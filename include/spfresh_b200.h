/*
 * spfresh_b200.h — C ABI of libspfresh_b200.so, the B200 (sm_100a) implementation of the
 * SPFresh/SPANN data-parallel hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each entry
 * point replaces one batched seam of the reference (jairad26/spfresh, Rust); the reference
 * interface it stands in for is cited as file:line next to it.  The binding a maintainer adds
 * on the Rust side (an `extern "C"` block in a `spann-cuda-sys` crate) is in INTEGRATION.md.
 *
 * Conventions:
 *  - every function returns SPF_OK (0) or a negative SPF_E_* code; the message is available
 *    from spf_last_error() (thread-local).  No exception or abort crosses the boundary.
 *  - host pointers passed in are only read during the call; nothing is retained.
 *  - row / point ids are uint64_t (Rust usize); cluster slots are uint32_t.
 *  - all floating point is IEEE f32; distances are computed exactly as the reference does
 *    (sequential, un-fused accumulation), so results are bit-identical to it.
 *  - there is no CPU fallback: without a CUDA device every compute call fails with
 *    SPF_E_NO_DEVICE.
 *  - a context may be used from several host threads; calls on one context are serialised.
 */
#ifndef SPFRESH_B200_H
#define SPFRESH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define SPF_ABI_VERSION 1

enum {
  SPF_OK = 0,
  SPF_E_INVALID = -1,    /* bad argument (null pointer, out-of-range row, k == 0, ...)      */
  SPF_E_NO_DEVICE = -2,  /* no usable CUDA device / not an sm_100 part                      */
  SPF_E_CUDA = -3,       /* a CUDA runtime or driver call failed                            */
  SPF_E_OOM = -4,        /* host or device allocation failed                                */
  SPF_E_IO = -5,         /* posting-list directory could not be read / written              */
  SPF_E_STATE = -6       /* call order violated (e.g. k-means++ session exhausted)          */
};

/* src/spann/config.rs:92-100 — "Euclidean" | "Manhattan" | "Chebyshev"
 * (src/distances/distance.rs:14-43).  Euclidean is the SQUARED L2 distance, as there. */
enum { SPF_METRIC_EUCLIDEAN = 0, SPF_METRIC_MANHATTAN = 1, SPF_METRIC_CHEBYSHEV = 2 };

/* spf_assign flags */
enum {
  SPF_ASSIGN_DEFAULT = 0,
  SPF_ASSIGN_FORCE_EXACT = 1,   /* Euclidean only: skip the tcgen05 candidate GEMM and use the
                                   CUDA-core direct-form kernel for all n*k distances          */
  SPF_ASSIGN_NO_CSR = 2         /* only best/dmin are needed (no boundary replication, no CSR) */
};

typedef struct spf_ctx spf_ctx;                  /* one CUDA device + stream + scratch        */
typedef struct spf_dataset spf_dataset;          /* n x d f32 rows resident in HBM            */
typedef struct spf_assign_result spf_assign_result;
typedef struct spf_kmpp spf_kmpp;                /* k-means++ session state in HBM            */
typedef struct spf_index spf_index;              /* posting lists + centroids resident in HBM */
typedef struct spf_comm spf_comm;                /* one rank of a multi-GPU group (NCCL)      */
typedef struct spf_kmeans spf_kmeans;            /* device-resident row-sharded k-means state */

/* ---- library / context ---------------------------------------------------------------- */
int         spf_abi_version(void);
const char* spf_last_error(void);
int         spf_device_count(void);

int  spf_ctx_create(int device, spf_ctx** out);
void spf_ctx_destroy(spf_ctx* ctx);
int  spf_ctx_device(const spf_ctx* ctx);
/* cudaStream_t all work of this context is enqueued on (for CUDA-event timing by callers). */
void* spf_ctx_stream(spf_ctx* ctx);
int  spf_ctx_synchronize(spf_ctx* ctx);
/* The library keeps freed device temporaries in the stream-ordered pool for reuse; this returns
 * everything that is currently unused to the driver (after a large one-off call, or before another
 * allocator in the same process needs the memory). */
int  spf_ctx_trim(spf_ctx* ctx);
/* Per-kernel device timing: when enabled, the library brackets its dominant kernels with CUDA
 * events on its stream; spf_ctx_kernel_ms returns the duration of the named kernel's last
 * launch ("assign_tc", "assign_exact", "resolve", "csr", "scan", "probe", ...) or < 0. */
int   spf_ctx_set_profiling(spf_ctx* ctx, int enabled);
float spf_ctx_kernel_ms(spf_ctx* ctx, const char* name);
/* Number of kernels this library launched on this context since creation. */
uint64_t spf_ctx_launch_count(const spf_ctx* ctx);
/* Diagnostics: points of the last spf_assign whose candidate buffers overflowed and that were
 * therefore resolved by the dense exact fallback (results are identical, only slower). */
uint32_t spf_ctx_last_overflow_rows(const spf_ctx* ctx);

/* ---- dataset -------------------------------------------------------------------------- *
 * Stands in for the borrowed ArrayView2<F> held by SpannIndexBuilder / HierarchicalClustering
 * (src/spann/spann_builder.rs:10,20; src/clustering/hierarchical.rs:45).  `rows` is row-major
 * with `row_stride` elements between rows (>= d).  The rows are copied to the device. */
int  spf_dataset_upload(spf_ctx* ctx, const float* rows, uint64_t n, uint32_t d,
                        uint64_t row_stride, spf_dataset** out);
/* Same, but `dev_rows` already is device memory of this context's device (dense, stride d);
 * it is copied device-to-device into the library's padded layout. */
int  spf_dataset_from_device(spf_ctx* ctx, const void* dev_rows, uint64_t n, uint32_t d,
                             spf_dataset** out);
void spf_dataset_free(spf_dataset* ds);
/* Copies m dataset rows (m x d, row-major) back to the host: out[i] = row rows[i].  The reference
 * indexes its host ArrayView2 directly; shards created on the device need this instead. */
int  spf_dataset_fetch_rows(spf_dataset* ds, const uint64_t* rows, uint64_t m, float* out);
uint64_t spf_dataset_rows(const spf_dataset* ds);
uint32_t spf_dataset_dim(const spf_dataset* ds);

/* ---- single-pair distance (API-compat seam) --------------------------------------------- *
 * DistanceMetric::compute, src/distances/distance.rs:7-43 — evaluated on the device for
 * `pairs` pairs: out[i] = metric(a + i*d, b + i*d).  Present so the per-pair trait has a
 * device-backed equivalent; batched callers use the entry points below. */
int spf_distance_pairs(spf_ctx* ctx, int metric, const float* a, const float* b, uint32_t d,
                       uint64_t pairs, float* out);

/* ---- assignment ------------------------------------------------------------------------ *
 * HierarchicalClustering::assign_points_to_clusters, src/clustering/hierarchical.rs:295-364
 * (and assign_points :368-390 when point_idx == NULL).
 *   point_idx     m dataset rows to assign, or NULL for rows 0..m-1 (m must equal n then)
 *   centroid_rows k dataset rows acting as centroids (Cluster::centroid_idx)
 *   boundary_factor  BOUNDARY_THRESHOLD, 1.1 in the reference (:55)
 * Result (device resident until fetched): per listed point the nearest centroid slot and its
 * distance, plus the cluster-major CSR of members (nearest + boundary replicas, input order
 * preserved inside each cluster) — exactly the Vec<Vec<usize>> the reference returns. */
int  spf_assign(spf_dataset* ds, int metric, const uint64_t* point_idx, uint64_t m,
                const uint64_t* centroid_rows, uint32_t k, float boundary_factor, int flags,
                spf_assign_result** out);
/* The same with the k centroids given as explicit vectors (k x d, row-major host memory) instead
 * of dataset rows: the row-sharded build, where a centroid may be a row of
 * another rank's shard. */
int  spf_assign_vectors(spf_dataset* ds, int metric, const uint64_t* point_idx, uint64_t m,
                        const float* centroids, uint32_t k, float boundary_factor, int flags,
                        spf_assign_result** out);
/* The same for rows that are still in HOST memory (the borrowed ArrayView2 of
 * src/spann/spann_builder.rs:20): uploads the n x d rows and assigns all of them to the k
 * centroid rows in one call.  The upload is chunked and overlapped with the kernels of the
 * previous chunk, so the call costs little more than the host-to-device copy itself (use pinned
 * host memory for full PCIe rate).  `rows` is only read during the call.  When ds_out is not NULL
 * it receives the resident dataset (for further spf_assign / spf_update_medoids_from calls);
 * otherwise the device copy is dropped. */
int  spf_assign_host(spf_ctx* ctx, const float* rows, uint64_t n, uint32_t d, uint64_t row_stride,
                     int metric, const uint64_t* centroid_rows, uint32_t k, float boundary_factor,
                     int flags, spf_dataset** ds_out, spf_assign_result** out);
uint64_t spf_assign_points(const spf_assign_result* r);        /* m                         */
uint32_t spf_assign_clusters(const spf_assign_result* r);      /* k                         */
uint64_t spf_assign_total(const spf_assign_result* r);         /* sum of cluster sizes      */
/* Copies what is non-NULL: best[m], dmin[m], offsets[k+1], members[total] (dataset rows). */
int  spf_assign_fetch(const spf_assign_result* r, uint32_t* best, float* dmin,
                      uint64_t* offsets, uint64_t* members);
void spf_assign_free(spf_assign_result* r);

/* ---- centroid update ------------------------------------------------------------------- *
 * HierarchicalClustering::update_centroids, src/clustering/hierarchical.rs:138-181, with
 * compute_mean src/clustering/utils.rs:5-15: per cluster the f32 mean over all members
 * (row-by-row sum in member order, then division) and the member nearest to it (leftmost on
 * ties); an empty cluster keeps old_rows[c].  means_out (k*d) may be NULL.
 * The CSR is either host arrays or a previous spf_assign result (no re-upload). */
int spf_update_medoids(spf_dataset* ds, int metric, const uint64_t* offsets,
                       const uint64_t* members, uint32_t k, const uint64_t* old_rows,
                       uint64_t* new_rows, float* means_out);
int spf_update_medoids_from(spf_dataset* ds, int metric, const spf_assign_result* r,
                            const uint64_t* old_rows, uint64_t* new_rows, float* means_out);

/* Row-sharded update_centroids (hierarchical.rs:138-181 split at its two reductions): every rank
 * calls spf_cluster_sums on its shard's assignment, the sums / counts are all-reduced and divided
 * (compute_mean, utils.rs:13-14), every rank calls spf_medoid_candidates with the global means,
 * and the (distance, rank order) minimum over ranks names the new centroid (:155-171).
 *   sums   k x d per-cluster f32 sums of this shard's member rows (member order), counts[k]
 *   dist / row   per cluster the best local member for `means` (k x d): its distance and dataset
 *                row, or (+inf, UINT64_MAX) when the shard has no candidate */
int spf_cluster_sums(spf_dataset* ds, const spf_assign_result* r, float* sums, uint64_t* counts);
int spf_medoid_candidates(spf_dataset* ds, int metric, const spf_assign_result* r, const float* means,
                          float* dist, uint64_t* row);

/* ---- multi-GPU group -------------------------------------------------- *
 * One process per GPU.  The reference is single-process (rayon, hierarchical.rs:144-303); the
 * exchange steps below are what north_star adds on one NVLink / NVSwitch box.  libnccl is loaded
 * at run time (dlopen "libnccl.so.2"), only when a group with world > 1 is created.
 *   spf_comm_unique_id   rank 0 creates the 128-byte NCCL id; the host passes it to every rank by
 *                        whatever channel it has (MPI, a TCP store, torch.distributed, a file)
 *   spf_comm_create      collective over all ranks; world == 1 needs no id and no NCCL */
#define SPF_COMM_ID_BYTES 128
int  spf_comm_unique_id(uint8_t* id128);
int  spf_comm_create(spf_ctx* ctx, int world, int rank, const uint8_t* id128, spf_comm** out);
void spf_comm_destroy(spf_comm* comm);
int  spf_comm_world(const spf_comm* comm);
int  spf_comm_rank(const spf_comm* comm);

/* Device-resident row-sharded k-means iteration: HierarchicalClustering::assign_points +
 * update_centroids (hierarchical.rs:368-390, 138-181) with the rows sharded contiguously over the
 * ranks (rank r holds global rows [row0, row0 + n), rank 0 starts at 0) and the k centroid vectors
 * replicated.  Nothing leaves the device between iterations; per iteration two NCCL all-gathers
 * on the context stream carry (C1) the per-cluster partial sums + counts, added in rank order and
 * divided (compute_mean, utils.rs:13-14), and (C2) every rank's best member per cluster for the
 * new mean with its vector; the (distance, rank) minimum names the new centroid (:155-171), an
 * empty cluster keeps its centroid (:146-149).  comm == NULL: one GPU, no exchange.
 *   set_centroids  k global rows + their vectors (k x d, host), e.g. from k-means++
 *   step           one iteration (collective: every rank of the group must call it)
 *   fetch          what is non-NULL: centroid rows[k], vectors[k*d], last means[k*d], global
 *                  cluster sizes[k]
 *   assignment     the rank's last local assignment (owned by the session, valid until the next
 *                  step / free), for spf_assign_fetch or spf_index_pack */
enum { SPF_KMEANS_DEFAULT = 0,
       SPF_KMEANS_UNSEEDED = 1, /* do not seed the assignment with the previous iteration's result */
       SPF_KMEANS_BALANCED = 2, /* extension: size-balancing penalty in the assignment (see below)  */
       SPF_KMEANS_MEANS = 4     /* extension: Lloyd iterations, the centroid is the mean itself     */ };
int  spf_kmeans_create(spf_dataset* ds, spf_comm* comm, int metric, uint64_t row0, uint32_t k,
                       float boundary_factor, int flags, spf_kmeans** out);
int  spf_kmeans_set_centroids(spf_kmeans* s, const uint64_t* global_rows, const float* vectors);
int  spf_kmeans_step(spf_kmeans* s);
int  spf_kmeans_fetch(spf_kmeans* s, uint64_t* rows, float* vectors, float* means, uint64_t* counts);
const spf_assign_result* spf_kmeans_assignment(const spf_kmeans* s);
/* EXTENSION beyond the reference (north_star: "the size-balancing penalty is applied in the same
 * pass"; the reference's fit() is one assign + one update, hierarchical.rs:65-71, so there is no
 * reference behaviour: the specification is oracle/spf_oracle.h orc_assign_balanced, parity unpinned).
 * With SPF_KMEANS_BALANCED a point goes to argmin_j fl(d(x, c_j) + lambda * n_j), n_j the global size
 * of cluster j after the previous iteration, and belongs to that cluster only.  The same assignment
 * as a single call, with explicit centroid vectors and penalties (k floats >= 0, NULL = zeros): */
int  spf_kmeans_set_balance(spf_kmeans* s, float lambda);
int  spf_assign_balanced(spf_dataset* ds, int metric, const uint64_t* point_idx, uint64_t m,
                         const float* centroids, const float* penalty, uint32_t k, int flags,
                         spf_assign_result** out);
void spf_kmeans_free(spf_kmeans* s);

/* ---- k-means++ ------------------------------------------------------------------------- *
 * HierarchicalClustering::initialize_clusters_kmeans_plus_plus, hierarchical.rs:249-293.
 * The host keeps the RNG (rand::SmallRng in the reference).  begin() takes the uniformly
 * drawn first row (:253-256).  Each round() folds the newest centroid into the running
 * minimum distance (bit-identical to the reference's per-round recomputation :260-276),
 * forms sum and weights d^2/max(sum,1e-10) (:278-282) and performs the weighted pick of
 * rand 0.9 WeightedIndex for the uniform draw u01 in [0,1) (:285-286).
 * round() returns SPF_OK with *chosen set, or 1 (positive) when the weighted pick is
 * impossible (all-zero or non-finite weights, the Err arm :287-290): the host then draws a
 * uniform row itself and calls spf_kmpp_push(). */
int  spf_kmpp_begin(spf_dataset* ds, int metric, uint64_t first_row, spf_kmpp** out);
int  spf_kmpp_round(spf_kmpp* s, double u01, uint64_t* chosen);
int  spf_kmpp_push(spf_kmpp* s, uint64_t row);
/* `count` rounds (:259-291) back to back without a host round trip per round: the draws u01[0..count)
 * are uploaded once, every round's update kernel reads the row the previous round picked from
 * device memory, chosen[0..*done) receives the picked rows.  Returns SPF_OK with *done == count, or
 * 1 when round *done (0-based within the batch) could not pick (its centroid is folded, the later
 * rounds did not run, draws u01[*done+1..) are unused): the host draws uniformly, calls
 * spf_kmpp_push() and carries on.  spf_kmpp_round() is this call with count = 1. */
int  spf_kmpp_rounds(spf_kmpp* s, const double* u01, uint32_t count, uint64_t* chosen, uint32_t* done);
/* Diagnostics of the last round: the f32 sum (:278) and the f64 weight total. */
int  spf_kmpp_last_sums(const spf_kmpp* s, float* sum, double* total);
/* Row-sharded form (one session per rank over its shard): the newest centroid
 * arrives as an explicit vector, the reductions of :278 and of WeightedIndex are done by the host
 * over the ranks:
 *   fold_vector   folds the centroid into the shard's running minimum, returns the shard's f32
 *                 sum of minimum distances (:260-278); the ranks' sums are added in rank order
 *   weight_total  with the global sum as denominator (:279-282): the shard's f64 weight total;
 *                 returns 1 when a weight of this shard is invalid (the Err arm, :287-290)
 *   pick_local    partition point of the shard's cumulative weights for a target already made
 *                 relative to the shard (u * total - sum of the lower ranks' totals) */
int  spf_kmpp_begin_sharded(spf_dataset* ds, int metric, spf_kmpp** out);
int  spf_kmpp_fold_vector(spf_kmpp* s, const float* centroid, float* local_sum);
int  spf_kmpp_weight_total(spf_kmpp* s, float global_sum, double* local_total);
int  spf_kmpp_pick_local(spf_kmpp* s, double target, uint64_t* row);
/* Device-resident form of the same rounds over the ranks of `comm` (NULL: one rank): no host round
 * trip per round.  spf_kmpp_set_vector() places the newest centroid's vector (d floats, the same on
 * every rank); spf_kmpp_rounds_sharded() then runs `count` rounds: fold, local f32 sum, all-gather of
 * the sums and rank-ordered add (:278), f64 weight totals all-gathered, the rank whose weight range
 * holds u01[i] * total picks inside its shard, the picked row's vector reaches every rank by a third
 * all-gather and becomes the next round's centroid.  Same arithmetic and the same picks as the
 * fold_vector / weight_total / pick_local sequence.  chosen[0..*done) receives GLOBAL rows (row_base +
 * the owner's local row; row_base = first global row of this rank's shard).  Returns 1 when round
 * *done could not pick (invalid or all-zero weights; its draw u01[*done] is NOT consumed): the host
 * draws uniformly, sets that row's vector with spf_kmpp_set_vector() and carries on.  Collective:
 * every rank calls it with the same count and draws. */
int  spf_kmpp_set_vector(spf_kmpp* s, const float* centroid);
int  spf_kmpp_rounds_sharded(spf_kmpp* s, spf_comm* comm, uint64_t row_base, const double* u01, uint32_t count,
                             uint64_t* chosen, uint32_t* done);
void spf_kmpp_free(spf_kmpp* s);

/* ---- bisect seed ----------------------------------------------------------------------- *
 * The fold of create_subclusters, hierarchical.rs:112-126: argmax over members != c1 of
 * d(c1, member), strict >, identity (row 0, distance 0). */
int spf_farthest(spf_dataset* ds, int metric, uint64_t c1_row, const uint64_t* members,
                 uint64_t m, uint64_t* out_row);
/* Row-sharded form ("for bisects, every GPU processes its slice of the member
 * list"): c1 arrives as an explicit vector (it may be a row of another rank's shard); skip_row is
 * the shard-local row of c1 when this rank owns it, UINT64_MAX otherwise.  Returns the largest
 * distance over this shard's members and the earliest member attaining it, or (0, UINT64_MAX) when
 * no member has a distance > 0; the ranks' answers are combined with strict > in rank order. */
int spf_farthest_from(spf_dataset* ds, int metric, const float* c1_vector, uint64_t skip_row,
                      const uint64_t* members, uint64_t m, float* out_dist, uint64_t* out_row);

/* ---- index (posting lists in HBM) + query ---------------------------------------------- *
 * spf_index_pack stands in for SpannIndex::create_posting_lists + build_kdtree,
 * src/spann/spann_index.rs:56-114: list c holds the vectors of rows
 * members[offsets[c]..offsets[c+1]) in that order and is represented by the dataset row
 * centroid_rows[c].  Instead of one bincode file per cluster read per probe
 * (src/spann/posting_lists.rs:98-106) the lists stay resident in HBM.
 * Multi-GPU: each rank packs only the lists [list_begin, list_end) it owns but all centroids. */
int  spf_index_pack(spf_dataset* ds, const uint64_t* offsets, const uint64_t* members,
                    const uint64_t* centroid_rows, uint32_t nlists,
                    uint32_t list_begin, uint32_t list_end, spf_index** out);
/* Load an index saved in the reference's on-disk layout (posting_list_{id}.bin +
 * cluster_ids.bin, bincode 1.3; src/spann/posting_lists.rs:42-45,64-129).  The reference
 * keeps centroids only inside output.kdtree (kiddo internals); the dense centroid matrix is
 * therefore passed in (nlists x d, list id order) — see INTEGRATION.md. */
int  spf_index_load_dir(spf_ctx* ctx, const char* dir, const float* centroids, uint32_t nlists,
                        uint32_t d, spf_index** out);
/* Write the lists of this index in the reference's on-disk layout. */
int  spf_index_save_dir(const spf_index* idx, const char* dir);
void spf_index_free(spf_index* idx);
uint32_t spf_index_lists(const spf_index* idx);
uint64_t spf_index_vectors(const spf_index* idx);   /* total vectors stored on this rank */

/* Batched SpannIndex::find_k_nearest_neighbor_spann, src/spann/spann_index.rs:148-197, for nq
 * queries (row-major nq x d):
 *   probe   the nprobe nearest centroids by squared L2, ascending (kiddo nearest_n :164);
 *           nprobe == 0 means nprobe = k, the reference behaviour;
 *   prune   thr = prune_factor * (d(q, c_nearest) + FLT_EPSILON), 1.2 in the reference (:165);
 *   scan    every vector of every probed list, keep dist <= thr (:168-179);
 *   top-k   stable sort by distance in encounter order, truncate to k, no de-duplication
 *           (:188-193).
 * Outputs: counts[q] <= k results per query (0 == the reference's None), ids / dists are
 * nq x k (rows padded with UINT64_MAX / +inf), vectors (nq x k x d) may be NULL.
 * keys (nq x k, may be NULL) receives the stable-order key of each result,
 * (distance bits << 32 | encounter index), which is what a multi-GPU merge sorts on.
 * Limits (the reference has none): 1 <= k <= 1024 (k > 128 runs the exact query-major scan in
 * passes of 128 results), nprobe <= 1024 (after clamping to the number of lists), fewer than
 * 2^32 (query, probe) pairs and 2^32 scanned vectors per query. */
int spf_search_batch(spf_index* idx, const float* queries, uint64_t nq, uint32_t k,
                     uint32_t nprobe, float prune_factor, uint64_t* ids, float* dists,
                     uint32_t* counts, float* vectors, uint64_t* keys);
/* Bytes of posting-list vector payload the last spf_search_batch streamed
 * (sum over queries and probed lists of |L| * d * 4): the roofline numerator of the scan. */
uint64_t spf_index_last_scan_bytes(const spf_index* idx);

/* The same over a list-sharded index (north_star: "posting lists are sharded for query, with
 * per-GPU top-k merged at the end"; the per-query work is spann_index.rs:148-197): every rank of
 * `comm` packed its own lists with spf_index_pack(list_begin, list_end) and brings nq_local queries
 * of the batch (the same count on every rank).  Collective; on return ids / dists / counts hold the
 * global top-k of this rank's queries.  Query slices and probe tables are all-gathered over NCCL,
 * partial top-k are exchanged so that each rank merges its own queries on the device.
 * comm == NULL: identical to spf_search_batch. */
int spf_search_sharded(spf_index* idx, spf_comm* comm, const float* queries, uint64_t nq_local,
                       uint32_t k, uint32_t nprobe, float prune_factor, uint64_t* ids, float* dists,
                       uint32_t* counts);

/* Merge per-rank partial results of spf_search_batch (list-sharded index): `parts` rank-major
 * arrays of nq x k keys / ids / dists, counts per rank and query; writes the global top-k. */
int spf_topk_merge(uint32_t parts, uint64_t nq, uint32_t k, const uint64_t* keys,
                   const uint64_t* ids, const float* dists, const uint32_t* counts,
                   uint64_t* out_ids, float* out_dists, uint32_t* out_counts);

/* ---- tuning knobs (not part of the drop-in surface; used by the tests to reach rare paths) -- *
 * "cand_cap" candidate group records per point (default 128), "short_cap" resolve short-list
 * entries per point (power of two <= 64, default 64), "chunk_rows" points per internal assign
 * chunk (0 = automatic), "work_cap" exact-evaluation work-list entries per chunk (0 = automatic),
 * "scan_list_major" (0 query-major, 1 automatic, 2 list-major), "scan_tc" (tensor-core candidate
 * scan: 0 never, 1 automatic, 2 whenever supported), "scan_tc_bucket" (candidates per query, 0 =
 * automatic: 256, or 1024 for d > 256), "scan_tc_tau_probes" (0 = all), "scan_tc_cmax_mb" (budget for the one-pass variant, 0 = two GEMM passes), "force_exact", "tc_min_k", "tc_min_m",
 * "no_host_staging" (1: pageable host buffers are handed to cudaMemcpyAsync instead of the library's threaded pinned staging ring),
 * "kmpp_exact_sum" (1: sequential f32 sum by scan, 2: by the serial add chain — same bits; 0: tree sum), "cc_matrix_max_k",
 * "medoid_direct" (update_centroids medoid pass: 1 stage the member rows only and read the cluster mean in place, 0 stage both
 * rows of every pair), "sum_fast" (compute_mean producer warps: 1 suspended barrier waits + short copy loop, 0 polling waits),
 * "sum_hub" (clusters of at least this many members take the deep producer configuration, 0 = 8192), "finalize_lanes" (8 / 16 / 32).
 * Apart from "kmpp_exact_sum" = 0 every knob value returns the same bits; the knobs exist for A/B timing and to reach rare paths in the tests. */
int spf_ctx_set_param(spf_ctx* ctx, const char* name, int value);

/* Test hook: the strictly sequential f32 fold of hierarchical.rs:278 over n host values.  mode 1 is
 * the single-CTA scan kernel, mode 3 the thread-block-cluster scan (the k-means++ rounds use it for
 * more than 16 384 elements when "kmpp_exact_sum" = 1), mode 2 the serial add chain both must equal
 * bit for bit. */
int spf_seq_sum_f32(spf_ctx* ctx, const float* values, uint64_t n, int mode, float* out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* SPFRESH_B200_H */
