/*
 * spf_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the jairad26/spfresh CPU hot path, used only as the checker in
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * Nothing under spfresh_b200/ may call, link or import it.
 *
 * What it follows (paths relative to the reference checkout):
 *   src/distances/distance.rs:16-43        three metrics (ndarray-stats 0.6 DeviationExt)
 *   src/clustering/utils.rs:5-15           compute_mean (ndarray 0.16 select + mean_axis)
 *   src/clustering/hierarchical.rs:65-390  fit / init / assign / update / bisect
 *   src/spann/spann_index.rs:148-197       find_k_nearest_neighbor_spann
 *   src/spann/posting_lists.rs:64-129      bincode 1.3 posting-list files
 *
 * PINNING STATUS.  The Rust reference cannot be built here (no cargo/rustc, no vendored
 * crates), so there is no oracle/_ref.  The oracle is pinned against every known-answer
 * vector the reference's own tests hold for this path (tests/golden/reference_kats.json,
 * extracted from distance.rs:52-104, utils.rs:24-32, hierarchical.rs:444-486,
 * examples/build_index.rs:9-25).  Those KATs do NOT pin the f32 accumulation order inside
 * ndarray-stats / ndarray, kiddo's equal-distance order, the rand 0.9 streams, or the bincode
 * files: for those the restatement follows the crates' published algorithms and is
 * "parity unpinned" (see DESIGN.md §Oracle).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (sequential, un-fused f32 like rustc).
 * All random decisions of the reference (rand::SmallRng) are explicit inputs here.
 */
#ifndef SPF_ORACLE_H
#define SPF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_EUCLIDEAN = 0, ORC_MANHATTAN = 1, ORC_CHEBYSHEV = 2 };

/* distance.rs:16-43 — sq_l2_dist / l1_dist / linf_dist, sequential accumulation in F. */
float  orc_distance_f32(int metric, const float* a, const float* b, size_t d);
double orc_distance_f64(int metric, const double* a, const double* b, size_t d);

/* utils.rs:5-15 — gather rows in idx order, row-by-row f32 sum, true division by m. */
void orc_compute_mean_f32(const float* data, size_t d, const uint64_t* idx, size_t m, float* out);
void orc_compute_mean_f64(const double* data, size_t d, const uint64_t* idx, size_t m, double* out);

/* Cluster-major CSR: offsets[k+1], members[offsets[k]] (dataset row ids, input order). */
typedef struct {
  uint64_t  k;
  uint64_t  m;        /* number of points assigned (rows of best/dmin) */
  uint64_t* offsets;  /* k+1 */
  uint64_t* members;  /* offsets[k] */
  uint32_t* best;     /* m : nearest centroid slot per listed point */
  float*    dmin;     /* m : its distance */
} orc_assign_t;

/* hierarchical.rs:295-364 assign_points_to_clusters.  point_idx == NULL means 0..m-1.
 * threads <= 0 → all online cores (one task per point, as rayon does). */
int  orc_assign(const float* data, size_t n, size_t d, int metric,
                const uint64_t* point_idx, size_t m,
                const uint64_t* centroid_rows, size_t k,
                float boundary_factor, int threads, orc_assign_t* out);
void orc_assign_free(orc_assign_t* a);

/* hierarchical.rs:138-181 update_centroids: mean over members then leftmost-argmin medoid.
 * Empty cluster keeps old_rows[c].  means_out (k*d) optional. */
int orc_update_medoids(const float* data, size_t n, size_t d, int metric,
                       const uint64_t* offsets, const uint64_t* members, size_t k,
                       const uint64_t* old_rows, uint64_t* new_rows, float* means_out,
                       int threads);

/* hierarchical.rs:249-293 KMeans++.  first_row = the uniform first pick; u01[r] in [0,1) is
 * the uniform draw of round r (scaled by the f64 weight total like rand 0.9 WeightedIndex);
 * fallback_rows[r] is used when the weighted pick errors (all-zero / non-finite weights).
 * naive != 0 recomputes every centroid distance each round exactly like the reference
 * (O(n k^2 d)); naive == 0 keeps a running minimum (bit-identical: min is order-free).
 * fell_back (k-1, optional) records which rounds used the fallback. */
int orc_kmeanspp(const float* data, size_t n, size_t d, int metric, size_t k,
                 uint64_t first_row, const double* u01, const uint64_t* fallback_rows,
                 int naive, int threads, uint64_t* out_rows, uint8_t* fell_back);

/* hierarchical.rs:112-126 second bisect seed: argmax_{idx != c1} d(c1, idx), strict >,
 * identity (0, 0.0). */
uint64_t orc_farthest(const float* data, size_t d, int metric, uint64_t c1,
                      const uint64_t* members, size_t m);

/* Clusters as the reference keeps them (hierarchical.rs:26-30). */
typedef struct {
  uint64_t  centroid;
  uint64_t* points;
  uint64_t  len;
  uint64_t  depth;
} orc_cluster_t;

typedef struct {
  orc_cluster_t* c;
  size_t count, cap;
} orc_clusters_t;

/* pick(ctx, len) → index in [0,len): stands in for points.choose(&mut rng)
 * (hierarchical.rs:111). */
typedef uint64_t (*orc_pick_fn)(void* ctx, uint64_t len);

/* hierarchical.rs:65-71 fit(), given the already-initialised centroid rows (init is a
 * separate call so that its random draws stay explicit).  max_splits bounds the
 * subdivide loop (the reference can spin forever on duplicate-heavy clusters). */
int  orc_fit(const float* data, size_t n, size_t d, int metric,
             const uint64_t* init_rows, size_t k, uint64_t desired_cluster_size,
             orc_pick_fn pick, void* pick_ctx, uint64_t max_splits, int threads,
             orc_clusters_t* out);
void orc_clusters_free(orc_clusters_t* cl);

/* spann_index.rs:148-197 single query against in-memory posting lists given as CSR over the
 * dataset (list c = rows members[offsets[c]..offsets[c+1]), centroid = centroid_rows[c]).
 * nprobe == 0 → nprobe = k (the reference behaviour).  prune_factor 1.2 in the reference.
 * Returns number of results (0 == None), ids/dists filled up to k. */
size_t orc_search_one(const float* data, size_t d,
                      const uint64_t* offsets, const uint64_t* members,
                      const uint64_t* centroid_rows, size_t nlists,
                      const float* query, size_t k, size_t nprobe, float prune_factor,
                      uint64_t* ids, float* dists);

/* Batched driver over orc_search_one (threads over queries). counts[q] results each. */
int orc_search_batch(const float* data, size_t d,
                     const uint64_t* offsets, const uint64_t* members,
                     const uint64_t* centroid_rows, size_t nlists,
                     const float* queries, size_t nq, size_t k, size_t nprobe,
                     float prune_factor, int threads,
                     uint64_t* ids, float* dists, uint32_t* counts);

/* posting_lists.rs:64-113 — bincode 1.3 files: posting_list_{id}.bin = u64 n, then n x
 * (u64 point_id, u64 d, d x f32);  cluster_ids.bin = u64 m, m x u64. */
int orc_posting_list_write(const char* dir, uint64_t cluster_id, const float* data, size_t d,
                           const uint64_t* members, size_t len);
int orc_cluster_ids_write(const char* dir, const uint64_t* ids, size_t m);
/* Reads one list; *ids_out / *vec_out are malloc'd (caller frees with orc_free). */
int orc_posting_list_read(const char* dir, uint64_t cluster_id, uint64_t* len_out,
                          uint64_t* d_out, uint64_t** ids_out, float** vec_out);
void orc_free(void* p);

int orc_online_cpus(void);

/* EXTENSION (north_star: "the size-balancing penalty is applied in the same pass"; the reference has
 * no counterpart — its fit() is one assign + one update, hierarchical.rs:65-71 — so this part of
 * the oracle is a specification written here, PARITY UNPINNED): balanced assignment against k
 * explicit centroid vectors.  cost(x, j) = fl(d(x, c_j) + penalty[j]) with d the reference metric
 * (distance.rs:16-43, sequential f32) and one f32 add; best = argmin cost by the reference's fold
 * (strict <, identity (0, +inf): lowest slot on ties, :317-326); every point belongs to exactly its
 * best cluster (no boundary replication).  penalty == NULL means all zeros. */
int orc_assign_balanced(const float* data, size_t d, int metric, const uint64_t* point_idx, size_t m,
                        const float* centroids, const float* penalty, size_t k, int threads,
                        uint32_t* best, float* cost);

#ifdef __cplusplus
}
#endif
#endif
