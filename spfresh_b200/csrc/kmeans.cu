// kmeans.cu — spf_kmeans: the device-resident, row-sharded k-means iteration.
//
// One iteration = HierarchicalClustering::assign_points + update_centroids
// (src/clustering/hierarchical.rs:368-390, 138-181) with the rows sharded contiguously over the
// ranks and the k centroid vectors replicated.  Everything stays on the device and on the
// context's stream; the only inter-rank traffic is two all-gathers per iteration:
//   C1  per-cluster partial sums (k x ld f32) + member counts (k u32)   -> summed in RANK ORDER on
//       every rank (a fixed order keeps the means identical on all ranks and reproducible; it
//       differs from one sequential pass over all members only by f32 rounding), then
//       the true division of compute_mean (src/clustering/utils.rs:13-14)
//   C2  per cluster the best local member for the new mean: (distance, global row) + its vector
//       -> minimum with strict < in rank order (lowest rank wins ties = the leftmost member, the
//       shards being contiguous row ranges, :155-171); the winner's vector becomes the centroid
// An empty cluster keeps its centroid (:146-149); a cluster whose distances are all inf/NaN takes
// global row 0 (the fold identity (0, +inf), :163).
// From the second iteration on the assignment is seeded: d(x, c_new[best_old(x)]) is an exact
// distance to one of the new centroids, i.e. a certified upper bound of the new minimum distance,
// which lets the tcgen05 candidate pass emit ~3x fewer candidates (assign_tc.cu).
//
// EXTENSION (north_star: "the centroid update is a fused segmented-sum/count kernel, and the
// size-balancing penalty is applied in the same pass"; the reference's fit() is a single assign +
// update, hierarchical.rs:65-71, so there is no reference behaviour to match — the oracle's
// orc_assign_balanced is the specification, PARITY UNPINNED):
//   SPF_KMEANS_BALANCED  assignment by cost(x, j) = fl(d(x, c_j) + lambda * n_j), n_j the global size
//                        of cluster j after the previous iteration (0 in the first); every point
//                        belongs to exactly its best cluster.  On the tensor path the penalty rides
//                        in the K extension of the GEMM next to |c|^2, so the candidate pass ranks by
//                        cost at no extra work; exact costs add the penalty with one f32 add.
//   SPF_KMEANS_MEANS     Lloyd iterations: the new centroid is the mean itself instead of the member
//                        nearest to it (no medoid pass, no C2 exchange).
#include "comm.cuh"
#include "kernels.cuh"

using namespace spf;

struct spf_kmeans {
  spf_dataset* ds = nullptr;
  spf_comm* comm = nullptr;
  int metric = 0;
  uint64_t row0 = 0;
  uint32_t k = 0;
  float factor = 1.1f;
  int flags = 0;
  bool have_centroids = false;
  uint64_t iterations = 0;
  spf_assign_result* last = nullptr;
  float lambda = 0.0f;          // balanced mode: penalty[j] = lambda * (float)(members of cluster j after the last iteration)
  DevBuf<float> cvec, means, seed, penalty;
  DevBuf<uint64_t> crow, gcount;
  DevBuf<uint8_t> msg1, gath1, msg2, gath2;
  size_t msg1_bytes = 0, msg2_bytes = 0;
};

namespace spf {
namespace {

struct CandMsg {          // 16 bytes per cluster in the C2 message
  uint64_t row;           // global row of the best local member, ~0 when there is none
  uint32_t dist_bits;     // its distance to the mean (+inf when there is none)
  uint32_t pad;
};

__global__ void km_counts_kernel(const uint64_t* __restrict__ offsets, uint32_t k, uint32_t* __restrict__ counts) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < k) counts[c] = (uint32_t)(offsets[c + 1] - offsets[c]);
}

// means[c] = (sum over ranks, in rank order, of the partial sums) / (float)(sum of counts); one
// thread per element.  part r starts at gath + r * stride_bytes: k * ld floats, then k u32 counts.
__global__ void km_reduce_means_kernel(const uint8_t* __restrict__ gath, size_t stride_bytes, int world, uint32_t k,
                                       uint32_t ld, float* __restrict__ means, uint64_t* __restrict__ gcount) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (uint64_t)k * ld) return;
  const uint32_t c = (uint32_t)(e / ld);
  float tot = 0.0f;
  unsigned long long cnt = 0;
  for (int r = 0; r < world; ++r) {
    const float* sums = reinterpret_cast<const float*>(gath + (size_t)r * stride_bytes);
    const uint32_t* counts = reinterpret_cast<const uint32_t*>(sums + (size_t)k * ld);
    tot = __fadd_rn(tot, sums[e]);
    cnt += counts[c];
  }
  means[e] = cnt ? __fdiv_rn(tot, __ull2float_rn(cnt)) : 0.0f;
  if (e == (uint64_t)c * ld) gcount[c] = cnt;
}

// C2 message of this rank: block per cluster (block k = the extra slot holding local row 0, which
// rank 0 contributes for the identity case)
__global__ void km_pack_cand_kernel(const unsigned long long* __restrict__ keys, const uint64_t* __restrict__ offsets,
                                    const uint32_t* __restrict__ members, const float* __restrict__ X, uint32_t ld,
                                    uint32_t k, uint64_t row0, CandMsg* __restrict__ cand, float* __restrict__ vec) {
  const uint32_t c = blockIdx.x;
  uint64_t lrow = ~0ull;
  if (c == k) {
    lrow = 0;
  } else {
    const unsigned long long key = keys[c];
    if (key != ~0ull) lrow = members[offsets[c] + (key & 0xffffffffull)];
    if (threadIdx.x == 0) {
      CandMsg m;
      m.row = lrow == ~0ull ? ~0ull : lrow + row0;
      m.dist_bits = key == ~0ull ? 0x7f800000u : (uint32_t)(key >> 32);
      m.pad = 0;
      cand[c] = m;
    }
  }
  for (uint32_t i = threadIdx.x; i < ld; i += blockDim.x)
    vec[(size_t)c * ld + i] = lrow == ~0ull ? 0.0f : X[(size_t)lrow * ld + i];
}

// new centroid of every cluster from the gathered candidates; part r of `gath`: k CandMsg, then
// (k + 1) x ld floats
__global__ void km_select_kernel(const uint8_t* __restrict__ gath, size_t stride_bytes, int world, uint32_t k, uint32_t ld,
                                 const uint64_t* __restrict__ gcount, uint64_t* __restrict__ crow, float* __restrict__ cvec) {
  const uint32_t c = blockIdx.x;
  if (gcount[c] == 0) return;                                   // empty cluster keeps its centroid (:146-149)
  float best_d = __int_as_float(0x7f800000);
  int best_r = -1;
  uint64_t best_row = 0;
  for (int r = 0; r < world; ++r) {
    const CandMsg m = reinterpret_cast<const CandMsg*>(gath + (size_t)r * stride_bytes)[c];
    const float dv = __uint_as_float(m.dist_bits);
    if (dv < best_d) { best_d = dv; best_r = r; best_row = m.row; }   // strict <: the lowest rank wins ties
  }
  const float* src;
  if (best_r < 0) {                                             // identity (0, +inf): global row 0, held by rank 0
    best_row = 0;
    src = reinterpret_cast<const float*>(gath + (size_t)k * sizeof(CandMsg)) + (size_t)k * ld;
  } else {
    src = reinterpret_cast<const float*>(gath + (size_t)best_r * stride_bytes + (size_t)k * sizeof(CandMsg)) + (size_t)c * ld;
  }
  if (threadIdx.x == 0) crow[c] = best_row;
  for (uint32_t i = threadIdx.x; i < ld; i += blockDim.x) cvec[(size_t)c * ld + i] = src[i];
}

__global__ void km_take_means_kernel(const float* __restrict__ means, uint32_t ld, const uint64_t* __restrict__ gcount,
                                     uint64_t* __restrict__ crow, float* __restrict__ cvec) {
  const uint32_t c = blockIdx.x;
  if (gcount[c] == 0) return;                                   // empty cluster keeps its centroid
  if (threadIdx.x == 0) crow[c] = ~0ull;                        // the centroid is no dataset row any more
  for (uint32_t i = threadIdx.x; i < ld; i += blockDim.x) cvec[(size_t)c * ld + i] = means[(size_t)c * ld + i];
}

int km_update(spf_kmeans* s) {
  spf_dataset* ds = s->ds;
  spf_ctx* c = ds->ctx;
  cudaStream_t st = c->stream;
  const uint32_t k = s->k, ld = ds->ld;
  const spf_assign_result* r = s->last;
  const int world = s->comm ? s->comm->world : 1;
  DevBuf<uint64_t> d_rows;
  DevBuf<unsigned long long> keys;
  SPF_TRY(d_rows.alloc_cached(c, "member_rows", r->total));
  SPF_TRY(assign_members_as_rows(r, d_rows.p));
  float* sums = reinterpret_cast<float*>(s->msg1.p);
  uint32_t* counts = reinterpret_cast<uint32_t*>(sums + (size_t)k * ld);
  {
    KernelTimer t(c, "kmeans_sums");
    SPF_TRY(launch_cluster_sums(c, ds->x, ld, r->offsets, d_rows.p, k, sums, 0));
    km_counts_kernel<<<(k + 255) / 256, 256, 0, st>>>(r->offsets, k, counts);
    SPF_TRY(check_launch(c, "km_counts_kernel"));
  }
  {
    KernelTimer t(c, "kmeans_exchange");
    SPF_TRY(comm_allgather(c, s->comm, s->msg1.p, s->gath1.p, s->msg1_bytes));
  }
  {
    KernelTimer t(c, "kmeans_means");
    km_reduce_means_kernel<<<(unsigned)ceil_div((uint64_t)k * ld, 256), 256, 0, st>>>(s->gath1.p, s->msg1_bytes, world, k, ld,
                                                                                      s->means.p, s->gcount.p);
    SPF_TRY(check_launch(c, "km_reduce_means_kernel"));
  }
  if (s->flags & SPF_KMEANS_MEANS) {           // Lloyd: centroid = mean of the non-empty clusters
    km_take_means_kernel<<<k, 128, 0, st>>>(s->means.p, ld, s->gcount.p, s->crow.p, s->cvec.p);
    return check_launch(c, "km_take_means_kernel");
  }
  SPF_TRY(keys.alloc(st, k));
  CandMsg* cand = reinterpret_cast<CandMsg*>(s->msg2.p);
  float* vec = reinterpret_cast<float*>(s->msg2.p + (size_t)k * sizeof(CandMsg));
  {
    KernelTimer t(c, "kmeans_medoid");
    SPF_TRY(launch_medoid_keys(c, s->metric, ds->x, ld, d_rows.p, r->total, r->offsets, k, s->means.p, keys.p));
    km_pack_cand_kernel<<<k + 1, 128, 0, st>>>(keys.p, r->offsets, r->members, ds->x, ld, k, s->row0, cand, vec);
    SPF_TRY(check_launch(c, "km_pack_cand_kernel"));
  }
  {
    KernelTimer t(c, "kmeans_exchange");
    SPF_TRY(comm_allgather(c, s->comm, s->msg2.p, s->gath2.p, s->msg2_bytes));
  }
  km_select_kernel<<<k, 128, 0, st>>>(s->gath2.p, s->msg2_bytes, world, k, ld, s->gcount.p, s->crow.p, s->cvec.p);
  return check_launch(c, "km_select_kernel");
}

}  // namespace
}  // namespace spf

extern "C" {

int spf_kmeans_create(spf_dataset* ds, spf_comm* comm, int metric, uint64_t row0, uint32_t k, float boundary_factor,
                      int flags, spf_kmeans** out) {
  return spf::guarded([&]() -> int {
  if (!ds || !out) return fail(SPF_E_INVALID, "spf_kmeans_create: NULL argument");
  *out = nullptr;
  if (metric < 0 || metric > 2) return fail(SPF_E_INVALID, "unknown metric %d", metric);
  if (k == 0) return fail(SPF_E_INVALID, "k must be > 0");
  if (comm && comm->ctx != ds->ctx) return fail(SPF_E_INVALID, "communicator and dataset belong to different contexts");
  if ((!comm || comm->rank == 0) && row0 != 0)
    return fail(SPF_E_INVALID, "rank 0 must hold global row 0 (contiguous row sharding in rank order)");
  spf_ctx* c = ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  spf_kmeans* s = new (std::nothrow) spf_kmeans();
  if (!s) return fail(SPF_E_OOM, "out of host memory");
  s->ds = ds; s->comm = comm; s->metric = metric; s->row0 = row0; s->k = k; s->factor = boundary_factor; s->flags = flags;
  const uint32_t ld = ds->ld;
  const int world = comm ? comm->world : 1;
  s->msg1_bytes = ((size_t)k * ld + k) * 4;
  s->msg2_bytes = (size_t)k * sizeof(CandMsg) + ((size_t)k + 1) * ld * 4;
  int rc = s->cvec.alloc(st, (size_t)k * ld);
  if (rc >= 0) rc = s->means.alloc(st, (size_t)k * ld);
  if (rc >= 0) rc = s->crow.alloc(st, k);
  if (rc >= 0) rc = s->gcount.alloc(st, k);
  if (rc >= 0) rc = s->penalty.alloc(st, k);
  if (rc >= 0) rc = s->msg1.alloc(st, s->msg1_bytes);
  if (rc >= 0) rc = s->gath1.alloc(st, s->msg1_bytes * world);
  if (rc >= 0) rc = s->msg2.alloc(st, s->msg2_bytes);
  if (rc >= 0) rc = s->gath2.alloc(st, s->msg2_bytes * world);
  if (rc < 0) { delete s; return rc; }
  *out = s;
  return SPF_OK;
  });
}

int spf_kmeans_set_centroids(spf_kmeans* s, const uint64_t* global_rows, const float* vectors) {
  return spf::guarded([&]() -> int {
  if (!s || !global_rows || !vectors) return fail(SPF_E_INVALID, "spf_kmeans_set_centroids: NULL argument");
  spf_ctx* c = s->ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const uint32_t ld = s->ds->ld, d = s->ds->d;
  if (ld != d) SPF_CUDA(cudaMemsetAsync(s->cvec.p, 0, (size_t)s->k * ld * sizeof(float), st));
  SPF_CUDA(cudaMemcpy2DAsync(s->cvec.p, (size_t)ld * 4, vectors, (size_t)d * 4, (size_t)d * 4, s->k, cudaMemcpyHostToDevice, st));
  SPF_CUDA(cudaMemcpyAsync(s->crow.p, global_rows, (size_t)s->k * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  s->have_centroids = true;
  s->iterations = 0;                                    // the next assignment is unseeded
  return SPF_OK;
  });
}

int spf_kmeans_step(spf_kmeans* s) {
  return spf::guarded([&]() -> int {
  if (!s) return fail(SPF_E_INVALID, "spf_kmeans_step: NULL argument");
  if (!s->have_centroids) return fail(SPF_E_STATE, "spf_kmeans_step: no centroids set");
  spf_dataset* ds = s->ds;
  spf_ctx* c = ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  c->kernel_ms.clear();
  const float* seed = nullptr;
  const float* penalty = nullptr;
  if (s->flags & SPF_KMEANS_BALANCED) {
    if (s->iterations == 0) SPF_CUDA(cudaMemsetAsync(s->penalty.p, 0, (size_t)s->k * sizeof(float), st));
    else SPF_TRY(launch_scale_u64_f32(c, s->gcount.p, s->lambda, s->k, s->penalty.p));
    penalty = s->penalty.p;
  }
  if (s->last && s->iterations > 0 && !(s->flags & (SPF_KMEANS_UNSEEDED | SPF_KMEANS_BALANCED))) {
    // exact distance of every point to the NEW centroid of the slot it was nearest to
    KernelTimer t(c, "kmeans_seed");
    SPF_TRY(s->seed.alloc(st, ds->n));
    SPF_TRY(launch_pair_dist(c, s->metric, ds->x, ds->ld, nullptr, s->cvec.p, ds->ld, s->last->best, UINT64_MAX, ds->ld,
                             ds->n, s->seed.p));
    seed = s->seed.p;
  }
  spf_assign_result* res = nullptr;
  SPF_TRY(assign_device_centroids(ds, s->metric, s->cvec.p, s->k, s->factor, 0, seed, penalty, &res));
  if (s->last) spf_assign_free(s->last);
  s->last = res;
  SPF_TRY(km_update(s));
  ++s->iterations;
  return SPF_OK;
  });
}

int spf_kmeans_fetch(spf_kmeans* s, uint64_t* rows, float* vectors, float* means, uint64_t* counts) {
  return spf::guarded([&]() -> int {
  if (!s) return fail(SPF_E_INVALID, "spf_kmeans_fetch: NULL argument");
  spf_ctx* c = s->ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const uint32_t ld = s->ds->ld, d = s->ds->d;
  if (rows) SPF_CUDA(cudaMemcpyAsync(rows, s->crow.p, (size_t)s->k * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  if (counts) SPF_CUDA(cudaMemcpyAsync(counts, s->gcount.p, (size_t)s->k * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  if (vectors)
    SPF_CUDA(cudaMemcpy2DAsync(vectors, (size_t)d * 4, s->cvec.p, (size_t)ld * 4, (size_t)d * 4, s->k, cudaMemcpyDeviceToHost, st));
  if (means)
    SPF_CUDA(cudaMemcpy2DAsync(means, (size_t)d * 4, s->means.p, (size_t)ld * 4, (size_t)d * 4, s->k, cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  return SPF_OK;
  });
}

int spf_kmeans_set_balance(spf_kmeans* s, float lambda) {
  if (!s) return fail(SPF_E_INVALID, "spf_kmeans_set_balance: NULL argument");
  if (!(lambda >= 0.0f)) return fail(SPF_E_INVALID, "lambda must be a non-negative number");
  if (!(s->flags & SPF_KMEANS_BALANCED)) return fail(SPF_E_STATE, "the session was not created with SPF_KMEANS_BALANCED");
  s->lambda = lambda;
  return SPF_OK;
}

const spf_assign_result* spf_kmeans_assignment(const spf_kmeans* s) { return s ? s->last : nullptr; }

void spf_kmeans_free(spf_kmeans* s) {
  if (!s) return;
  cudaSetDevice(s->ds->ctx->device);
  if (s->last) spf_assign_free(s->last);
  delete s;
}

}  // extern "C"
