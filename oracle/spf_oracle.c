/*
 * spf_oracle.c — CPU ORACLE (test infrastructure; see spf_oracle.h for scope and pinning).
 * Compile with -O2 -ffp-contract=off: every f32 sum below must stay sequential and un-fused,
 * like the rustc build of the reference.
 */
#define _GNU_SOURCE
#include "spf_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------
 * distances — src/distances/distance.rs:16-43 → ndarray-stats 0.6 DeviationExt:
 *   sq_l2_dist : acc += (a-b)*(a-b)        l1_dist : acc += |a-b|
 *   linf_dist  : if |a-b| > max { max = |a-b| }, starting from 0
 * element order, accumulator of type F, no FMA, no reassociation.
 * ---------------------------------------------------------------------------------------- */
#define DEFINE_DIST(NAME, T, ABS)                                                  \
  static inline T NAME##_sql2(const T* a, const T* b, size_t d) {                  \
    T acc = (T)0;                                                                  \
    for (size_t i = 0; i < d; ++i) { T df = a[i] - b[i]; T sq = df * df; acc = acc + sq; } \
    return acc;                                                                    \
  }                                                                                \
  static inline T NAME##_l1(const T* a, const T* b, size_t d) {                    \
    T acc = (T)0;                                                                  \
    for (size_t i = 0; i < d; ++i) { T df = a[i] - b[i]; acc = acc + ABS(df); }    \
    return acc;                                                                    \
  }                                                                                \
  static inline T NAME##_linf(const T* a, const T* b, size_t d) {                  \
    T mx = (T)0;                                                                   \
    for (size_t i = 0; i < d; ++i) { T df = ABS(a[i] - b[i]); if (df > mx) mx = df; } \
    return mx;                                                                     \
  }
DEFINE_DIST(f32, float, fabsf)
DEFINE_DIST(f64, double, fabs)

float orc_distance_f32(int metric, const float* a, const float* b, size_t d) {
  switch (metric) {
    case ORC_EUCLIDEAN: return f32_sql2(a, b, d);
    case ORC_MANHATTAN: return f32_l1(a, b, d);
    default:            return f32_linf(a, b, d);
  }
}
double orc_distance_f64(int metric, const double* a, const double* b, size_t d) {
  switch (metric) {
    case ORC_EUCLIDEAN: return f64_sql2(a, b, d);
    case ORC_MANHATTAN: return f64_l1(a, b, d);
    default:            return f64_linf(a, b, d);
  }
}

/* ------------------------------------------------------------------------------------------
 * compute_mean — src/clustering/utils.rs:5-15.  ndarray 0.16 select(Axis(0)) gathers the rows
 * in idx order into a C-order matrix; mean_axis(Axis(0)) = sum_axis / F::from_usize(m), and
 * sum_axis over the non-contiguous axis is `res = res + row` row by row from zeros.
 * ---------------------------------------------------------------------------------------- */
void orc_compute_mean_f32(const float* data, size_t d, const uint64_t* idx, size_t m, float* out) {
  for (size_t j = 0; j < d; ++j) out[j] = 0.0f;
  if (m == 0) return;
  for (size_t r = 0; r < m; ++r) {
    const float* row = data + (size_t)idx[r] * d;
    for (size_t j = 0; j < d; ++j) out[j] = out[j] + row[j];
  }
  float fm = (float)m;
  for (size_t j = 0; j < d; ++j) out[j] = out[j] / fm;
}
void orc_compute_mean_f64(const double* data, size_t d, const uint64_t* idx, size_t m, double* out) {
  for (size_t j = 0; j < d; ++j) out[j] = 0.0;
  if (m == 0) return;
  for (size_t r = 0; r < m; ++r) {
    const double* row = data + (size_t)idx[r] * d;
    for (size_t j = 0; j < d; ++j) out[j] = out[j] + row[j];
  }
  double fm = (double)m;
  for (size_t j = 0; j < d; ++j) out[j] = out[j] / fm;
}

/* ------------------------------------------------------------------------------------------
 * tiny parallel-for (stands in for rayon's par_iter: results are gathered by index, so the
 * outcome never depends on scheduling).
 * ---------------------------------------------------------------------------------------- */
int orc_online_cpus(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

typedef void (*range_fn)(void* ctx, size_t lo, size_t hi);
typedef struct {
  range_fn fn; void* ctx; size_t n, chunk; size_t next; pthread_mutex_t mu;
} pf_t;

static void* pf_worker(void* p) {
  pf_t* s = (pf_t*)p;
  for (;;) {
    pthread_mutex_lock(&s->mu);
    size_t lo = s->next; s->next += s->chunk;
    pthread_mutex_unlock(&s->mu);
    if (lo >= s->n) break;
    size_t hi = lo + s->chunk; if (hi > s->n) hi = s->n;
    s->fn(s->ctx, lo, hi);
  }
  return NULL;
}

static void parallel_for(size_t n, int threads, size_t chunk, range_fn fn, void* ctx) {
  if (threads <= 0) threads = orc_online_cpus();
  if (n == 0) return;
  if (chunk == 0) chunk = 1;
  if (threads == 1 || n <= chunk) { fn(ctx, 0, n); return; }
  pf_t s; s.fn = fn; s.ctx = ctx; s.n = n; s.chunk = chunk; s.next = 0;
  pthread_mutex_init(&s.mu, NULL);
  if (threads > 256) threads = 256;
  pthread_t th[256];
  int started = 0;
  for (int t = 0; t < threads - 1; ++t)
    if (pthread_create(&th[started], NULL, pf_worker, &s) == 0) ++started;
  pf_worker(&s);
  for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
  pthread_mutex_destroy(&s.mu);
}

/* ------------------------------------------------------------------------------------------
 * assign_points_to_clusters — src/clustering/hierarchical.rs:295-364
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint32_t* v; uint32_t len, cap; } u32vec;

typedef struct {
  const float* data; size_t d; int metric;
  const uint64_t* point_idx; const uint64_t* crow; size_t k; float factor;
  uint32_t* best; float* dmin; u32vec* lists;
} assign_ctx;

static void assign_range(void* p, size_t lo, size_t hi) {
  assign_ctx* c = (assign_ctx*)p;
  float* dist = (float*)malloc(sizeof(float) * (c->k ? c->k : 1));
  for (size_t i = lo; i < hi; ++i) {
    size_t row = c->point_idx ? (size_t)c->point_idx[i] : i;
    const float* x = c->data + row * c->d;
    /* :309-315 all distances, point first, centroid second */
    for (size_t j = 0; j < c->k; ++j)
      dist[j] = orc_distance_f32(c->metric, x, c->data + (size_t)c->crow[j] * c->d, c->d);
    /* :317-326 fold from (0, +inf) with strict < → lowest slot wins ties */
    size_t bj = 0; float bd = INFINITY;
    for (size_t j = 0; j < c->k; ++j) if (dist[j] < bd) { bd = dist[j]; bj = j; }
    /* :328 threshold in F */
    float thr = bd * c->factor;
    const float* c1 = c->data + (size_t)c->crow[bj] * c->d;
    u32vec* L = &c->lists[i];
    L->len = 0;
    /* :331-346 — members are emitted in slot order, best included at its slot */
    for (size_t j = 0; j < c->k; ++j) {
      int take = 0;
      if (j == bj) take = 1;
      else if (dist[j] < thr) {
        float cc = orc_distance_f32(c->metric, c1, c->data + (size_t)c->crow[j] * c->d, c->d);
        if (cc >= dist[j]) take = 1;
      }
      if (take) {
        if (L->len == L->cap) {
          L->cap = L->cap ? L->cap * 2 : 4;
          L->v = (uint32_t*)realloc(L->v, sizeof(uint32_t) * L->cap);
        }
        L->v[L->len++] = (uint32_t)j;
      }
    }
    c->best[i] = (uint32_t)bj;
    c->dmin[i] = bd;
  }
  free(dist);
}

int orc_assign(const float* data, size_t n, size_t d, int metric,
               const uint64_t* point_idx, size_t m,
               const uint64_t* centroid_rows, size_t k,
               float boundary_factor, int threads, orc_assign_t* out) {
  (void)n;
  memset(out, 0, sizeof(*out));
  out->k = k; out->m = m;
  out->offsets = (uint64_t*)calloc(k + 1, sizeof(uint64_t));
  out->best = (uint32_t*)calloc(m ? m : 1, sizeof(uint32_t));
  out->dmin = (float*)calloc(m ? m : 1, sizeof(float));
  u32vec* lists = (u32vec*)calloc(m ? m : 1, sizeof(u32vec));
  if (!out->offsets || !out->best || !out->dmin || !lists) return -1;
  if (k == 0) {  /* reference: fold yields (0, inf), then centroids[0] panics */
    free(lists); out->members = (uint64_t*)calloc(1, sizeof(uint64_t)); return m ? -2 : 0;
  }
  assign_ctx c = { data, d, metric, point_idx, centroid_rows, k, boundary_factor,
                   out->best, out->dmin, lists };
  parallel_for(m, threads, 64, assign_range, &c);
  /* :353-361 serial merge in input order */
  for (size_t i = 0; i < m; ++i)
    for (uint32_t t = 0; t < lists[i].len; ++t) out->offsets[lists[i].v[t] + 1]++;
  for (size_t j = 0; j < k; ++j) out->offsets[j + 1] += out->offsets[j];
  uint64_t total = out->offsets[k];
  out->members = (uint64_t*)malloc(sizeof(uint64_t) * (total ? total : 1));
  uint64_t* cur = (uint64_t*)malloc(sizeof(uint64_t) * (k + 1));
  memcpy(cur, out->offsets, sizeof(uint64_t) * (k + 1));
  for (size_t i = 0; i < m; ++i) {
    uint64_t row = point_idx ? point_idx[i] : (uint64_t)i;
    for (uint32_t t = 0; t < lists[i].len; ++t) out->members[cur[lists[i].v[t]]++] = row;
    free(lists[i].v);
  }
  free(cur); free(lists);
  return 0;
}

/* EXTENSION (parity unpinned, see spf_oracle.h): balanced assignment */
typedef struct {
  const float* data; size_t d; int metric; const uint64_t* point_idx;
  const float* centroids; const float* penalty; size_t k; uint32_t* best; float* cost;
} bal_ctx;

static void bal_range(void* p, size_t lo, size_t hi) {
  bal_ctx* c = (bal_ctx*)p;
  for (size_t i = lo; i < hi; ++i) {
    size_t row = c->point_idx ? (size_t)c->point_idx[i] : i;
    uint32_t bj = 0; float bd = INFINITY;
    for (size_t j = 0; j < c->k; ++j) {
      float dd = orc_distance_f32(c->metric, c->data + row * c->d, c->centroids + j * c->d, c->d);
      float cc = c->penalty ? dd + c->penalty[j] : dd;      /* one f32 add (-ffp-contract=off) */
      if (cc < bd) { bd = cc; bj = (uint32_t)j; }
    }
    c->best[i] = bj; c->cost[i] = bd;
  }
}

int orc_assign_balanced(const float* data, size_t d, int metric, const uint64_t* point_idx, size_t m,
                        const float* centroids, const float* penalty, size_t k, int threads,
                        uint32_t* best, float* cost) {
  if (k == 0) return -2;
  bal_ctx c = { data, d, metric, point_idx, centroids, penalty, k, best, cost };
  parallel_for(m, threads, 64, bal_range, &c);
  return 0;
}

void orc_assign_free(orc_assign_t* a) {
  if (!a) return;
  free(a->offsets); free(a->members); free(a->best); free(a->dmin);
  memset(a, 0, sizeof(*a));
}

/* ------------------------------------------------------------------------------------------
 * update_centroids — src/clustering/hierarchical.rs:138-181
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const float* data; size_t d; int metric;
  const uint64_t* offsets; const uint64_t* members; const uint64_t* old_rows;
  uint64_t* new_rows; float* means_out;
} upd_ctx;

static void upd_range(void* p, size_t lo, size_t hi) {
  upd_ctx* c = (upd_ctx*)p;
  float* mean = (float*)malloc(sizeof(float) * (c->d ? c->d : 1));
  for (size_t j = lo; j < hi; ++j) {
    const uint64_t* pts = c->members + c->offsets[j];
    size_t m = (size_t)(c->offsets[j + 1] - c->offsets[j]);
    if (m == 0) {                                   /* :146-149 */
      c->new_rows[j] = c->old_rows[j];
      if (c->means_out) for (size_t t = 0; t < c->d; ++t) c->means_out[j * c->d + t] = 0.0f;
      continue;
    }
    orc_compute_mean_f32(c->data, c->d, pts, m, mean);   /* :152 */
    /* :155-171 reduce from (0, +inf), strict <, left operand kept on ties → leftmost min */
    uint64_t bi = 0; float bd = INFINITY;
    for (size_t t = 0; t < m; ++t) {
      float dd = orc_distance_f32(c->metric, c->data + (size_t)pts[t] * c->d, mean, c->d);
      if (dd < bd) { bd = dd; bi = pts[t]; }
    }
    c->new_rows[j] = bi;
    if (c->means_out) memcpy(c->means_out + j * c->d, mean, sizeof(float) * c->d);
  }
  free(mean);
}

int orc_update_medoids(const float* data, size_t n, size_t d, int metric,
                       const uint64_t* offsets, const uint64_t* members, size_t k,
                       const uint64_t* old_rows, uint64_t* new_rows, float* means_out,
                       int threads) {
  (void)n;
  upd_ctx c = { data, d, metric, offsets, members, old_rows, new_rows, means_out };
  parallel_for(k, threads, 1, upd_range, &c);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * KMeans++ — src/clustering/hierarchical.rs:249-293
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const float* data; size_t d; int metric; const uint64_t* rows; size_t nc; int naive;
  float* mind;
} kpp_ctx;

static void kpp_range(void* p, size_t lo, size_t hi) {
  kpp_ctx* c = (kpp_ctx*)p;
  for (size_t i = lo; i < hi; ++i) {
    const float* x = c->data + i * c->d;
    if (c->naive) {                                 /* :265-274 min over all centroids */
      float best = orc_distance_f32(c->metric, x, c->data + (size_t)c->rows[0] * c->d, c->d);
      for (size_t j = 1; j < c->nc; ++j) {
        float dd = orc_distance_f32(c->metric, x, c->data + (size_t)c->rows[j] * c->d, c->d);
        if (dd < best) best = dd;
      }
      c->mind[i] = best;
    } else {                                        /* running min against the newest one */
      float dd = orc_distance_f32(c->metric, x, c->data + (size_t)c->rows[c->nc - 1] * c->d, c->d);
      if (c->nc == 1 || dd < c->mind[i]) c->mind[i] = dd;
    }
  }
}

int orc_kmeanspp(const float* data, size_t n, size_t d, int metric, size_t k,
                 uint64_t first_row, const double* u01, const uint64_t* fallback_rows,
                 int naive, int threads, uint64_t* out_rows, uint8_t* fell_back) {
  if (k == 0 || n == 0) return -1;
  float* mind = (float*)malloc(sizeof(float) * n);
  if (!mind) return -1;
  out_rows[0] = first_row;                          /* :253-256 */
  for (size_t r = 1; r < k; ++r) {                  /* :259 */
    kpp_ctx c = { data, d, metric, out_rows, r, naive, mind };
    parallel_for(n, threads, 256, kpp_range, &c);
    float sum = 0.0f;                               /* :278 sequential fold in F */
    for (size_t i = 0; i < n; ++i) sum = sum + mind[i];
    float denom = fmaxf(sum, 1e-10f);               /* :281 F::max(sum, F::from(1e-10)) */
    /* :285-286 rand 0.9 WeightedIndex::new over f64 weights, then
     * cumulative.partition_point(|w| w <= u),  u = u01 * total                      */
    int bad = 0; double total = 0.0;
    for (size_t i = 0; i < n; ++i) {
      float w = (mind[i] * mind[i]) / denom;        /* :281 */
      double wd = (double)w;
      if (!(wd >= 0.0)) { bad = 1; break; }
      total += wd;
    }
    if (!bad && (total == 0.0 || !isfinite(total))) bad = 1;
    uint64_t chosen;
    if (bad) {                                      /* :287-290 uniform fallback */
      chosen = fallback_rows ? fallback_rows[r - 1] : 0;
    } else {
      double u = u01[r - 1] * total;
      double cum = 0.0; size_t idx = n - 1;
      for (size_t i = 0; i + 1 < n; ++i) {
        float w = (mind[i] * mind[i]) / denom;
        cum += (double)w;
        if (!(cum <= u)) { idx = i; break; }
      }
      chosen = (uint64_t)idx;
    }
    if (fell_back) fell_back[r - 1] = (uint8_t)bad;
    out_rows[r] = chosen;
  }
  free(mind);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * create_subclusters / subdivide_clusters / fit — hierarchical.rs:65-135
 * ---------------------------------------------------------------------------------------- */
uint64_t orc_farthest(const float* data, size_t d, int metric, uint64_t c1,
                      const uint64_t* members, size_t m) {
  uint64_t mi = 0; float md = 0.0f;                 /* :115 identity (0, F::zero()) */
  for (size_t t = 0; t < m; ++t) {
    if (members[t] == c1) continue;                 /* :114 */
    float dd = orc_distance_f32(metric, data + (size_t)c1 * d, data + (size_t)members[t] * d, d);
    if (dd > md) { md = dd; mi = members[t]; }      /* :120 strict > */
  }
  return mi;
}

static void clusters_push(orc_clusters_t* cl, orc_cluster_t c) {
  if (cl->count == cl->cap) {
    cl->cap = cl->cap ? cl->cap * 2 : 8;
    cl->c = (orc_cluster_t*)realloc(cl->c, sizeof(orc_cluster_t) * cl->cap);
  }
  cl->c[cl->count++] = c;
}

static uint64_t* dup_u64(const uint64_t* src, size_t n) {
  uint64_t* p = (uint64_t*)malloc(sizeof(uint64_t) * (n ? n : 1));
  if (n) memcpy(p, src, sizeof(uint64_t) * n);
  return p;
}

int orc_fit(const float* data, size_t n, size_t d, int metric,
            const uint64_t* init_rows, size_t k, uint64_t desired_cluster_size,
            orc_pick_fn pick, void* pick_ctx, uint64_t max_splits, int threads,
            orc_clusters_t* out) {
  memset(out, 0, sizeof(*out));
  /* :67 assign_points over all rows */
  orc_assign_t a;
  int rc = orc_assign(data, n, d, metric, NULL, n, init_rows, k, 1.1f, threads, &a);
  if (rc) { orc_assign_free(&a); return rc; }
  /* :68 update_centroids */
  uint64_t* rows = (uint64_t*)malloc(sizeof(uint64_t) * (k ? k : 1));
  orc_update_medoids(data, n, d, metric, a.offsets, a.members, k, init_rows, rows, NULL, threads);
  for (size_t j = 0; j < k; ++j) {
    orc_cluster_t c;
    c.centroid = rows[j]; c.depth = 0;
    c.len = a.offsets[j + 1] - a.offsets[j];
    c.points = dup_u64(a.members + a.offsets[j], (size_t)c.len);
    clusters_push(out, c);
  }
  free(rows);
  orc_assign_free(&a);
  /* :74-105 subdivide_clusters */
  size_t i = 0; uint64_t splits = 0;
  while (i < out->count) {
    if (out->c[i].len > desired_cluster_size) {
      if (max_splits && splits >= max_splits) return 1;   /* guard: reference would spin */
      ++splits;
      orc_cluster_t cur = out->c[i];
      uint64_t depth = cur.depth + 1;
      /* :107-135 create_subclusters */
      uint64_t c1 = cur.points[pick(pick_ctx, cur.len)];
      uint64_t c2 = orc_farthest(data, d, metric, c1, cur.points, (size_t)cur.len);
      uint64_t cr[2] = { c1, c2 };
      orc_assign_t s;
      rc = orc_assign(data, n, d, metric, cur.points, (size_t)cur.len, cr, 2, 1.1f, threads, &s);
      if (rc) { orc_assign_free(&s); return rc; }
      orc_cluster_t s1, s2;
      s1.centroid = c1; s1.depth = depth; s1.len = s.offsets[1] - s.offsets[0];
      s1.points = dup_u64(s.members + s.offsets[0], (size_t)s1.len);
      s2.centroid = c2; s2.depth = depth; s2.len = s.offsets[2] - s.offsets[1];
      s2.points = dup_u64(s.members + s.offsets[1], (size_t)s2.len);
      orc_assign_free(&s);
      free(cur.points);
      out->c[i] = s1;                               /* :95 */
      clusters_push(out, s2);                       /* :98 */
    } else {
      ++i;
    }
  }
  return 0;
}

void orc_clusters_free(orc_clusters_t* cl) {
  if (!cl) return;
  for (size_t i = 0; i < cl->count; ++i) free(cl->c[i].points);
  free(cl->c);
  memset(cl, 0, sizeof(*cl));
}

/* ------------------------------------------------------------------------------------------
 * find_k_nearest_neighbor_spann — src/spann/spann_index.rs:148-197
 * ---------------------------------------------------------------------------------------- */
typedef struct { float d; uint64_t id; } cand_t;

static void stable_sort_cands(cand_t* a, cand_t* tmp, size_t n) {  /* merge sort, stable */
  if (n < 2) return;
  size_t h = n / 2;
  stable_sort_cands(a, tmp, h);
  stable_sort_cands(a + h, tmp, n - h);
  size_t i = 0, j = h, o = 0;
  while (i < h && j < n) {
    /* sort_by(partial_cmp().unwrap_or(Equal)): take right only when strictly less */
    if (a[j].d < a[i].d) tmp[o++] = a[j++]; else tmp[o++] = a[i++];
  }
  while (i < h) tmp[o++] = a[i++];
  while (j < n) tmp[o++] = a[j++];
  memcpy(a, tmp, sizeof(cand_t) * n);
}

size_t orc_search_one(const float* data, size_t d,
                      const uint64_t* offsets, const uint64_t* members,
                      const uint64_t* centroid_rows, size_t nlists,
                      const float* query, size_t k, size_t nprobe, float prune_factor,
                      uint64_t* ids, float* dists) {
  if (nprobe == 0) nprobe = k;                      /* :164 nearest_n(query, k) */
  if (nlists == 0 || k == 0) return 0;
  if (nprobe > nlists) nprobe = nlists;
  /* kiddo nearest_n::<SquaredEuclidean>: exact, ascending; equal distances ordered by
   * cluster id here (kiddo leaves that order unspecified — documented near-tie). */
  cand_t* cd = (cand_t*)malloc(sizeof(cand_t) * nlists);
  cand_t* tmp = (cand_t*)malloc(sizeof(cand_t) * nlists);
  for (size_t c = 0; c < nlists; ++c) {
    cd[c].d = f32_sql2(query, data + (size_t)centroid_rows[c] * d, d);
    cd[c].id = c;
  }
  stable_sort_cands(cd, tmp, nlists);
  free(tmp);
  float thr = prune_factor * (cd[0].d + FLT_EPSILON);   /* :165 */
  size_t cap = 64, cnt = 0;
  cand_t* all = (cand_t*)malloc(sizeof(cand_t) * cap);
  for (size_t p = 0; p < nprobe; ++p) {             /* :168 */
    size_t c = (size_t)cd[p].id;
    for (uint64_t t = offsets[c]; t < offsets[c + 1]; ++t) {
      float dist = f32_sql2(query, data + (size_t)members[t] * d, d);   /* :172 */
      if (dist <= thr) {                            /* :176 */
        if (cnt == cap) { cap *= 2; all = (cand_t*)realloc(all, sizeof(cand_t) * cap); }
        all[cnt].d = dist; all[cnt].id = members[t]; ++cnt;
      }
    }
  }
  free(cd);
  if (cnt == 0) { free(all); return 0; }            /* :183-186 → None */
  cand_t* t2 = (cand_t*)malloc(sizeof(cand_t) * cnt);
  stable_sort_cands(all, t2, cnt);                  /* :188-189 */
  free(t2);
  size_t outn = cnt < k ? cnt : k;                  /* :191-193 */
  for (size_t i = 0; i < outn; ++i) { ids[i] = all[i].id; dists[i] = all[i].d; }
  free(all);
  return outn;
}

typedef struct {
  const float* data; size_t d; const uint64_t* offsets; const uint64_t* members;
  const uint64_t* crow; size_t nlists; const float* q; size_t k, nprobe; float pf;
  uint64_t* ids; float* dists; uint32_t* counts;
} sb_ctx;

static void sb_range(void* p, size_t lo, size_t hi) {
  sb_ctx* c = (sb_ctx*)p;
  for (size_t q = lo; q < hi; ++q)
    c->counts[q] = (uint32_t)orc_search_one(c->data, c->d, c->offsets, c->members, c->crow,
                                            c->nlists, c->q + q * c->d, c->k, c->nprobe, c->pf,
                                            c->ids + q * c->k, c->dists + q * c->k);
}

int orc_search_batch(const float* data, size_t d,
                     const uint64_t* offsets, const uint64_t* members,
                     const uint64_t* centroid_rows, size_t nlists,
                     const float* queries, size_t nq, size_t k, size_t nprobe,
                     float prune_factor, int threads,
                     uint64_t* ids, float* dists, uint32_t* counts) {
  sb_ctx c = { data, d, offsets, members, centroid_rows, nlists, queries, k, nprobe,
               prune_factor, ids, dists, counts };
  parallel_for(nq, threads, 1, sb_range, &c);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * posting-list files — src/spann/posting_lists.rs:64-113 (bincode 1.3 defaults: little
 * endian, fixed-width ints, usize→u64, Vec = u64 length + items, struct fields in order)
 * ---------------------------------------------------------------------------------------- */
static void put_u64(FILE* f, uint64_t v) {
  unsigned char b[8];
  for (int i = 0; i < 8; ++i) b[i] = (unsigned char)(v >> (8 * i));
  fwrite(b, 1, 8, f);
}
static int get_u64(FILE* f, uint64_t* v) {
  unsigned char b[8];
  if (fread(b, 1, 8, f) != 8) return -1;
  uint64_t r = 0;
  for (int i = 0; i < 8; ++i) r |= (uint64_t)b[i] << (8 * i);
  *v = r; return 0;
}

int orc_posting_list_write(const char* dir, uint64_t cluster_id, const float* data, size_t d,
                           const uint64_t* members, size_t len) {
  char path[4096];
  snprintf(path, sizeof(path), "%s/posting_list_%llu.bin", dir, (unsigned long long)cluster_id);
  FILE* f = fopen(path, "wb");
  if (!f) return -1;
  put_u64(f, (uint64_t)len);
  for (size_t i = 0; i < len; ++i) {
    put_u64(f, members[i]);                         /* point_id: usize */
    put_u64(f, (uint64_t)d);                        /* vector: Vec<f32> length */
    const float* row = data + (size_t)members[i] * d;
    for (size_t j = 0; j < d; ++j) {
      uint32_t bits; memcpy(&bits, &row[j], 4);
      unsigned char b[4] = { (unsigned char)bits, (unsigned char)(bits >> 8),
                             (unsigned char)(bits >> 16), (unsigned char)(bits >> 24) };
      fwrite(b, 1, 4, f);
    }
  }
  fclose(f);
  return 0;
}

int orc_cluster_ids_write(const char* dir, const uint64_t* ids, size_t m) {
  char path[4096];
  snprintf(path, sizeof(path), "%s/cluster_ids.bin", dir);
  FILE* f = fopen(path, "wb");
  if (!f) return -1;
  put_u64(f, (uint64_t)m);
  for (size_t i = 0; i < m; ++i) put_u64(f, ids[i]);
  fclose(f);
  return 0;
}

int orc_posting_list_read(const char* dir, uint64_t cluster_id, uint64_t* len_out,
                          uint64_t* d_out, uint64_t** ids_out, float** vec_out) {
  char path[4096];
  snprintf(path, sizeof(path), "%s/posting_list_%llu.bin", dir, (unsigned long long)cluster_id);
  FILE* f = fopen(path, "rb");
  if (!f) return -1;
  uint64_t n = 0, d = 0;
  if (get_u64(f, &n)) { fclose(f); return -2; }
  uint64_t* ids = (uint64_t*)malloc(sizeof(uint64_t) * (n ? n : 1));
  float* vec = NULL;
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t di;
    if (get_u64(f, &ids[i]) || get_u64(f, &di)) { fclose(f); free(ids); free(vec); return -2; }
    if (i == 0) { d = di; size_t tot = (size_t)(n * d); vec = (float*)malloc(sizeof(float) * (tot ? tot : 1)); }
    else if (di != d) { fclose(f); free(ids); free(vec); return -3; }
    for (uint64_t j = 0; j < d; ++j) {
      unsigned char b[4];
      if (fread(b, 1, 4, f) != 4) { fclose(f); free(ids); free(vec); return -2; }
      uint32_t bits = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
      memcpy(&vec[i * d + j], &bits, 4);
    }
  }
  fclose(f);
  *len_out = n; *d_out = d; *ids_out = ids; *vec_out = vec;
  return 0;
}

void orc_free(void* p) { free(p); }
