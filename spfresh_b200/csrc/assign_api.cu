// assign_api.cu — spf_assign: the batched replacement of
// HierarchicalClustering::assign_points_to_clusters (src/clustering/hierarchical.rs:295-364).
#include <string.h>
#include <atomic>
#include <functional>
#include <thread>

#include <vector>

#include "kernels.cuh"

using namespace spf;


namespace spf {

namespace {

__global__ void positions_to_rows_kernel(const uint32_t* __restrict__ pos, const uint64_t* __restrict__ pidx,
                                         uint64_t total, uint64_t* __restrict__ out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < total) out[t] = pidx ? pidx[pos[t]] : (uint64_t)pos[t];
}

}  // namespace

int assign_members_as_rows(const spf_assign_result* r, uint64_t* d_out) {
  if (r->total == 0) return SPF_OK;
  positions_to_rows_kernel<<<(unsigned)ceil_div(r->total, 256), 256, 0, r->ctx->stream>>>(
      r->members, r->point_idx, r->total, d_out);
  return check_launch(r->ctx, "positions_to_rows_kernel");
}

// Rounded copy, squared norms and rounding residuals of all dataset rows (tensor path only):
// allocated by dataset_prep_alloc, filled once per dataset by dataset_prep (or chunk by chunk by
// the host-streamed assign).
int dataset_prep_alloc(spf_dataset* ds) {
  if (ds->xtf) return SPF_OK;
  cudaStream_t st = ds->ctx->stream;
  float *tf = nullptr, *nrm = nullptr, *res = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&tf, (size_t)ds->n * ds->ld * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMallocAsync((void**)&nrm, (size_t)ds->n * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMallocAsync((void**)&res, (size_t)ds->n * sizeof(float), st);
  if (e != cudaSuccess) {
    if (tf) cudaFreeAsync(tf, st);
    if (nrm) cudaFreeAsync(nrm, st);
    if (res) cudaFreeAsync(res, st);
    return fail(SPF_E_OOM, "allocation of the rounded dataset copy failed: %s", cudaGetErrorString(e));
  }
  ds->xtf = tf;
  ds->xnorm = nrm;
  ds->xres = res;
  return SPF_OK;
}

int dataset_prep(spf_dataset* ds) {
  if (ds->prepped) return SPF_OK;
  SPF_TRY(dataset_prep_alloc(ds));
  SPF_TRY(launch_row_prep(ds->ctx, ds->x, ds->ld, nullptr, ds->n, ds->xtf, ds->xnorm, ds->xres));
  ds->prepped = true;
  return SPF_OK;
}

// ---- pageable host buffers -------------------------------------------------------------------------
// cudaMemcpyAsync from ordinary heap memory is staged by the driver on one thread.  The reference's
// callers hold ndarray views (heap memory), so the host-facing entry points stage such buffers
// themselves: worker threads copy 4 MB blocks into a ring of pinned slots (rows already in the padded
// device pitch), the calling thread hands every filled slot to the copy engine in block order and
// launches a chunk's kernels as soon as its last block is queued.
bool host_pointer_is_pageable(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return at.type == cudaMemoryTypeUnregistered;
}

int stage_ring(spf_ctx* c) {
  if (c->stage.base) return SPF_OK;
  SPF_CUDA(cudaHostAlloc((void**)&c->stage.base, spf_ctx::HostStage::SLOTS * spf_ctx::HostStage::SLOT_BYTES, cudaHostAllocDefault));
  for (cudaEvent_t& e : c->stage.ev) SPF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return SPF_OK;
}

int stage_threads() {
  const unsigned hc = std::thread::hardware_concurrency();
  unsigned t = hc ? hc / 2 : 4;                            // the copies are memory bound: a few cores saturate the DRAM channels
  if (t < 2) t = 2;
  if (t > 8) t = 8;
  return (int)t;
}

struct StageBlock { uint64_t r0; uint32_t rows; uint32_t chunk; bool last_in_chunk; };

// rows [0, n) of a pageable row-major host matrix -> dst (pitch ld floats, zero padded) on `copy`;
// after_chunk(ci) runs on the calling thread right after chunk ci's last block is queued.
int staged_upload(spf_ctx* c, const float* rows, uint64_t n, uint32_t d, uint64_t row_stride, float* dst, uint32_t ld,
                  uint64_t chunk_rows, cudaStream_t copy, const std::function<int(uint64_t)>& after_chunk) {
  constexpr int SLOTS = spf_ctx::HostStage::SLOTS;
  constexpr size_t SLOT_BYTES = spf_ctx::HostStage::SLOT_BYTES;
  SPF_TRY(stage_ring(c));
  const size_t row_bytes = (size_t)ld * sizeof(float);
  if (row_bytes > SLOT_BYTES) return fail(SPF_E_INVALID, "row too long for the staging ring");
  const uint32_t brows = (uint32_t)(SLOT_BYTES / row_bytes);
  std::vector<StageBlock> blocks;
  for (uint64_t c0 = 0, ci = 0; c0 < n; c0 += chunk_rows, ++ci) {
    const uint64_t c1 = c0 + chunk_rows < n ? c0 + chunk_rows : n;
    for (uint64_t r = c0; r < c1; r += brows) {
      const uint32_t nr = (uint32_t)(c1 - r < brows ? c1 - r : brows);
      blocks.push_back({r, nr, (uint32_t)ci, r + nr == c1});
    }
  }
  const size_t nb = blocks.size();
  std::vector<std::atomic<int>> filled(nb);
  for (auto& f : filled) f.store(0, std::memory_order_relaxed);
  std::atomic<size_t> next{0}, submitted{0};
  std::atomic<int> err{0};
  const int device = c->device;
  auto worker = [&]() {
    cudaSetDevice(device);
    for (;;) {
      const size_t b = next.fetch_add(1);
      if (b >= nb || err.load()) return;
      const int slot = (int)(b % SLOTS);
      if (b >= (size_t)SLOTS) {                            // the slot's previous block must have left it
        while (submitted.load(std::memory_order_acquire) < b - SLOTS + 1) {
          if (err.load()) return;
          std::this_thread::yield();
        }
        if (cudaEventSynchronize(c->stage.ev[slot]) != cudaSuccess) { err.store(1); return; }
      }
      const StageBlock& blk = blocks[b];
      uint8_t* out = c->stage.base + (size_t)slot * SLOT_BYTES;
      const float* src = rows + blk.r0 * row_stride;
      if (ld == d && row_stride == d) {
        memcpy(out, src, (size_t)blk.rows * row_bytes);
      } else {
        for (uint32_t r = 0; r < blk.rows; ++r) {
          float* o = reinterpret_cast<float*>(out + (size_t)r * row_bytes);
          memcpy(o, src + (size_t)r * row_stride, (size_t)d * sizeof(float));
          for (uint32_t j = d; j < ld; ++j) o[j] = 0.0f;
        }
      }
      filled[b].store(1, std::memory_order_release);
    }
  };
  std::vector<std::thread> pool;
  const int nt = stage_threads();
  try {
    for (int t = 0; t < nt; ++t) pool.emplace_back(worker);
  } catch (...) {                                          // thread limit reached: carry on with the ones that started
    if (pool.empty()) return fail(SPF_E_OOM, "staged upload: cannot start a worker thread");
  }
  int rc = SPF_OK;
  for (size_t b = 0; b < nb && rc == SPF_OK; ++b) {
    while (!filled[b].load(std::memory_order_acquire)) {
      if (err.load()) { rc = fail(SPF_E_CUDA, "staged upload: a worker failed"); break; }
      std::this_thread::yield();
    }
    if (rc != SPF_OK) break;
    const int slot = (int)(b % SLOTS);
    const StageBlock& blk = blocks[b];
    cudaError_t e = cudaMemcpyAsync(dst + blk.r0 * ld, c->stage.base + (size_t)slot * SLOT_BYTES, (size_t)blk.rows * row_bytes,
                                    cudaMemcpyHostToDevice, copy);
    if (e == cudaSuccess) e = cudaEventRecord(c->stage.ev[slot], copy);
    if (e != cudaSuccess) { rc = fail(SPF_E_CUDA, "staged upload: %s", cudaGetErrorString(e)); break; }
    submitted.store(b + 1, std::memory_order_release);
    if (blk.last_in_chunk) rc = after_chunk(blk.chunk);
  }
  if (rc != SPF_OK) err.store(1);
  for (auto& t : pool) t.join();
  return rc;
}

// device -> pageable host through the same ring: the copy engine fills slots, workers copy them out
int staged_download(spf_ctx* c, void* host, const void* dev, size_t bytes, cudaStream_t st) {
  constexpr int SLOTS = spf_ctx::HostStage::SLOTS;
  constexpr size_t SLOT_BYTES = spf_ctx::HostStage::SLOT_BYTES;
  if (bytes == 0) return SPF_OK;
  SPF_TRY(stage_ring(c));
  const size_t nb = (bytes + SLOT_BYTES - 1) / SLOT_BYTES;
  std::atomic<size_t> queued{0}, next{0};
  std::vector<std::atomic<int>> drained(nb);
  for (auto& f : drained) f.store(0, std::memory_order_relaxed);
  std::atomic<int> err{0};
  const int device = c->device;
  auto worker = [&]() {
    cudaSetDevice(device);
    for (;;) {
      const size_t b = next.fetch_add(1);
      if (b >= nb || err.load()) return;
      while (queued.load(std::memory_order_acquire) < b + 1) {
        if (err.load()) return;
        std::this_thread::yield();
      }
      const int slot = (int)(b % SLOTS);
      if (cudaEventSynchronize(c->stage.ev[slot]) != cudaSuccess) { err.store(1); return; }
      const size_t off = b * SLOT_BYTES, len = bytes - off < SLOT_BYTES ? bytes - off : SLOT_BYTES;
      memcpy((uint8_t*)host + off, c->stage.base + (size_t)slot * SLOT_BYTES, len);
      drained[b].store(1, std::memory_order_release);
    }
  };
  std::vector<std::thread> pool;
  const int nt = stage_threads();
  try {
    for (int t = 0; t < nt; ++t) pool.emplace_back(worker);
  } catch (...) {
    if (pool.empty()) return fail(SPF_E_OOM, "staged download: cannot start a worker thread");
  }
  int rc = SPF_OK;
  for (size_t b = 0; b < nb; ++b) {
    const int slot = (int)(b % SLOTS);
    if (b >= (size_t)SLOTS)
      while (!drained[b - SLOTS].load(std::memory_order_acquire)) {
        if (err.load()) break;
        std::this_thread::yield();
      }
    if (err.load()) { rc = fail(SPF_E_CUDA, "staged download: a worker failed"); break; }
    const size_t off = b * SLOT_BYTES, len = bytes - off < SLOT_BYTES ? bytes - off : SLOT_BYTES;
    cudaError_t e = cudaMemcpyAsync(c->stage.base + (size_t)slot * SLOT_BYTES, (const uint8_t*)dev + off, len, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaEventRecord(c->stage.ev[slot], st);
    if (e != cudaSuccess) { rc = fail(SPF_E_CUDA, "staged download: %s", cudaGetErrorString(e)); break; }
    queued.store(b + 1, std::memory_order_release);
  }
  if (rc != SPF_OK) err.store(1);
  for (auto& t : pool) t.join();
  if (rc == SPF_OK && err.load()) rc = fail(SPF_E_CUDA, "staged download: a worker failed");
  return rc;
}

}  // namespace spf

namespace spf {

namespace {

// One assign call: centroid-side operands prepared once, the point list processed in chunks
// (bounded scratch; the host-streamed entry point overlaps uploads with the chunks' kernels).
struct AssignCall {
  spf_ctx* c = nullptr;
  int metric = 0;
  uint32_t k = 0, ld = 0;
  float factor = 1.0f;
  bool want_members = true, use_tc = false;
  const float* seed = nullptr;   // optional device array (m): upper bounds of the minimum distances
  const float* penalty = nullptr;   // optional device array (k): balanced assignment, cost = fl(d + penalty[j])
  uint32_t eld = 0;              // row length entering the certified error bound (ld, + slack with penalties)
  uint64_t m = 0, chunk_rows = 0;
  DevBuf<float> Cg, ctf, cnorm, cres, cstat, cext, cc_own;
  const float* cc = nullptr;     // k x k exact centroid-centroid distances (own buffer or the context's cache)
  DevBuf<CandRec> cand_rec;
  DevBuf<RowInfo> cand_info;
  CandBuf cand;
  DevBuf<uint32_t> best, nmem;
  DevBuf<float> dmin;
  ResolveState* rs = nullptr;
  ~AssignCall() { if (rs) resolve_free(rs); }
};

// Candidate group records per point.  Long rows have concentrated distances (the 1.1 boundary band
// covers far more centroids until the running minimum has tightened) and a wider TF32 bound, so the
// tensor path keeps four times as many records per point there unless the knob was set explicitly.
int effective_cand_cap(const spf_ctx* c, bool use_tc, uint32_t ld) {
  if (c->params.cand_cap == 128 && use_tc && ld > 256) return 512;
  // tensor path, short rows: 96 records per segment.  With 64, about 0.1 % of the points of the
  // 1M x 128 N(0,1) benchmark overflow a segment and go through the dense fallback (0.1 ms per step);
  // the records are written sparsely, so the larger scratch costs address space, not traffic.
  if (c->params.cand_cap == 128 && use_tc) return 192;
  return c->params.cand_cap;
}

uint64_t pick_chunk_rows(const spf_ctx* c, uint64_t m, bool streamed, int cand_cap) {
  // multiples of one full wave of the tensor kernel (one 128-point row block per SM)
  uint64_t rows = (uint64_t)c->sm_count * 128 * (streamed ? 4 : 64);
  // candidate scratch of a chunk stays below ~8 GB
  while (rows > (uint64_t)c->sm_count * 128 && rows * (uint64_t)cand_cap * sizeof(CandRec) > (8ull << 30)) rows /= 2;
  if (c->params.chunk_rows > 0) rows = (uint64_t)c->params.chunk_rows;
  return rows < m ? rows : m;
}

// Everything that depends only on the centroids (Cg already holds the k gathered rows).
int assign_setup(AssignCall& a) {
  spf_ctx* c = a.c;
  cudaStream_t st = c->stream;
  a.cand.cap = effective_cand_cap(c, a.use_tc, a.ld);
  SPF_TRY(a.cand_rec.alloc_cached(c, "cand_rec", (size_t)a.chunk_rows * a.cand.cap));
  SPF_TRY(a.cand_info.alloc_cached(c, "cand_info", a.chunk_rows));
  a.cand.rec = a.cand_rec.p;
  a.cand.info = a.cand_info.p;
  if (a.use_tc) {
    const uint32_t kpad = round_up(a.k, 256);
    SPF_TRY(a.ctf.alloc(st, (size_t)a.k * a.ld));
    SPF_TRY(a.cnorm.alloc(st, kpad));
    SPF_TRY(a.cres.alloc(st, a.k));
    SPF_TRY(a.cstat.alloc(st, 2));
    SPF_TRY(launch_row_prep(c, a.Cg.p, a.ld, nullptr, a.k, a.ctf.p, a.cnorm.p, a.cres.p));
    // balanced assignment: the penalty rides with |c|^2 in the K extension, s = x.c - (|c|^2 + p)/2, so
    // the accumulator yields the approximate COST |x|^2 - 2 s.  The two extra roundings (|c|^2 + p,
    // d + p) are covered by 8 more units of row length in the error bound.
    a.eld = a.ld + (a.penalty ? 8u : 0u);
    if (a.penalty) SPF_TRY(launch_add_f32(c, a.cnorm.p, a.penalty, a.k));
    SPF_TRY(launch_max2_f32(c, a.cnorm.p, a.cres.p, a.k, a.cstat.p));
    SPF_TRY(a.cext.alloc(st, (size_t)kpad * 8));
    SPF_TRY(launch_centroid_ext(c, a.cnorm.p, a.k, kpad, a.cext.p));
  }
  // exact centroid-centroid distances for the boundary rule `d(c_best, c_j) >= d_j` (:337-342)
  if (a.want_members && a.k > 1 && (int)a.k <= c->params.cc_matrix_max_k) {
    KernelTimer t(c, "cc_matrix");
    if (c->params.cc_cache != 0 && a.k >= 64 && a.k <= 8192) {
      // The matrix depends only on the centroid vectors: keep it in the context and recompute it
      // only when they changed.  The comparison runs on the device (no host round trip): the
      // matrix kernel itself looks at the flag and returns at once when the cache is current.
      spf_ctx::CcCache& cc = c->cc_cache;
      const size_t nC = (size_t)a.k * a.ld;
      if (cc.k != a.k || cc.ld != a.ld || !cc.cc) {
        SPF_CUDA(cudaStreamSynchronize(st));
        if (cc.cc) cudaFree(cc.cc);
        if (cc.C) cudaFree(cc.C);
        cc.cc = cc.C = nullptr;
        cc.valid = false;
        cc.k = a.k; cc.ld = a.ld;
        if (!cc.same) SPF_CUDA(cudaMalloc((void**)&cc.same, sizeof(int)));
        cudaError_t e = cudaMalloc((void**)&cc.cc, (size_t)a.k * a.k * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc((void**)&cc.C, nC * sizeof(float));
        if (e != cudaSuccess) {
          if (cc.cc) cudaFree(cc.cc);
          cc.cc = nullptr; cc.k = 0;
          return fail(SPF_E_OOM, "allocation of the centroid matrix failed: %s", cudaGetErrorString(e));
        }
      }
      const int init = (cc.valid && cc.metric == a.metric) ? 1 : 0;
      SPF_CUDA(cudaMemsetAsync(cc.same, init, sizeof(int), st));
      if (init) SPF_TRY(launch_rows_equal(c, a.Cg.p, cc.C, nC, cc.same));
      SPF_TRY(launch_assign_exact(c, a.metric, a.Cg.p, a.k, a.Cg.p, a.k, a.ld, 1.0f, nullptr, cc.cc, cc.same));
      SPF_CUDA(cudaMemcpyAsync(cc.C, a.Cg.p, nC * sizeof(float), cudaMemcpyDeviceToDevice, st));
      cc.metric = a.metric;
      cc.valid = true;
      a.cc = cc.cc;
    } else {
      SPF_TRY(a.cc_own.alloc(st, (size_t)a.k * a.k));
      SPF_TRY(launch_assign_exact(c, a.metric, a.Cg.p, a.k, a.Cg.p, a.k, a.ld, 1.0f, nullptr, a.cc_own.p));
      a.cc = a.cc_own.p;
    }
  }
  SPF_TRY(a.best.alloc(st, a.m));
  SPF_TRY(a.dmin.alloc(st, a.m));
  SPF_TRY(a.nmem.alloc_cached(c, "nmem", a.m));
  SPF_TRY(resolve_begin(c, a.m, a.chunk_rows, a.use_tc, a.want_members, &a.rs));
  return SPF_OK;
}

ResolveArgs resolve_args(const AssignCall& a, const float* P, uint64_t m, const float* xnorm, const float* xres,
                         uint64_t r0) {
  ResolveArgs r;
  r.metric = a.metric; r.P = P; r.m = m; r.C = a.Cg.p; r.k = a.k; r.ld = a.ld; r.factor = a.factor;
  r.cand = a.cand; r.nseg = a.use_tc ? 2 : 1;
  r.seed = (a.use_tc && a.seed) ? a.seed + r0 : nullptr;
  r.penalty = a.penalty;
  r.eld = a.eld;
  r.xnorm = a.use_tc ? xnorm : nullptr; r.xres = a.use_tc ? xres : nullptr;
  r.d_cstat = a.use_tc ? a.cstat.p : nullptr;
  r.cc = a.cc; r.want_members = a.want_members;
  r.best = a.best.p + r0; r.dmin = a.dmin.p + r0; r.nmem = a.nmem.p + r0;
  return r;
}

// Candidate kernel + resolve for the points [r0, r0 + mc) of the list; all pointers chunk-relative.
int assign_rows(AssignCall& a, const float* P, const float* Ptf, const float* xnorm, const float* xres,
                uint64_t r0, uint64_t mc) {
  spf_ctx* c = a.c;
  if (a.use_tc) {
    KernelTimer t(c, "assign_tc");
    SPF_TRY(launch_assign_tc(c, Ptf, mc, a.ctf.p, a.k, a.ld, xnorm, xres, a.cext.p, a.cstat.p,
                             a.seed ? a.seed + r0 : nullptr, a.factor, a.cand, a.eld));
  } else {
    KernelTimer t(c, "assign_exact");
    SPF_TRY(launch_assign_exact(c, a.metric, P, mc, a.Cg.p, a.k, a.ld, a.factor, &a.cand, nullptr, nullptr, a.penalty,
                                (a.seed && !a.penalty) ? a.seed + r0 : nullptr));
  }
  return resolve_chunk(c, a.rs, resolve_args(a, P, mc, xnorm, xres, r0), r0);
}

// Overflow rows, CSR, and the result object (takes ownership of the output buffers).
int assign_finish(AssignCall& a, const float* P_all, const float* xnorm_all, const float* xres_all,
                  DevBuf<uint64_t>* d_pidx, spf_assign_result** out) {
  spf_ctx* c = a.c;
  cudaStream_t st = c->stream;
  spf_assign_result* r = new (std::nothrow) spf_assign_result();
  if (!r) return fail(SPF_E_OOM, "out of host memory");
  r->ctx = c;
  r->m = a.m;
  r->k = a.k;
  CsrOut csr;
  ResolveArgs ra = resolve_args(a, P_all, a.m, xnorm_all, xres_all, 0);
  int rc = resolve_finish(c, a.rs, ra, a.want_members ? &csr : nullptr);
  if (rc >= 0) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(SPF_E_CUDA, "assign failed on the device: %s", cudaGetErrorString(e));
  }
  if (rc < 0) {
    if (csr.offsets) cudaFreeAsync(csr.offsets, st);
    if (csr.members) cudaFreeAsync(csr.members, st);
    delete r;
    return rc;
  }
  r->best = a.best.take();
  r->dmin = a.dmin.take();
  r->has_csr = a.want_members;
  r->total = csr.total;
  r->offsets = csr.offsets;
  r->members = csr.members;
  r->point_idx = (d_pidx && d_pidx->p) ? d_pidx->take() : nullptr;
  *out = r;
  return SPF_OK;
}

}  // namespace

}  // namespace spf

// Shared body of spf_assign (centroids = dataset rows) and spf_assign_vectors (centroids = explicit
// k x d host vectors, the sharded build where a centroid may live on another rank).
// centroid_dev: the vectors already on the device (k x ld, padded); d_seed: optional device array of
// m per-point upper bounds of the minimum distance.  The caller holds the context lock.
static int assign_resident(spf_dataset* ds, int metric, const uint64_t* point_idx, uint64_t m,
                           const uint64_t* centroid_rows, const float* centroid_vecs, const float* centroid_dev,
                           uint32_t k, float boundary_factor, int flags, const float* d_seed, const float* d_penalty,
                           spf_assign_result** out) {
  if (!ds || !out || (!centroid_rows && !centroid_vecs && !centroid_dev))
    return fail(SPF_E_INVALID, "spf_assign: NULL argument");
  *out = nullptr;
  if (metric < 0 || metric > 2) return fail(SPF_E_INVALID, "unknown metric %d", metric);
  if (k == 0) return fail(SPF_E_INVALID, "k must be > 0 (the reference indexes centroids[0])");
  if (k > (REC_G_MASK << 2)) return fail(SPF_E_INVALID, "k must be < 2^30");
  if (!point_idx && m != ds->n) return fail(SPF_E_INVALID, "point_idx == NULL requires m == n");
  if (m == 0) return fail(SPF_E_INVALID, "m must be > 0");
  if (m >= (1ull << 32)) return fail(SPF_E_INVALID, "m must be < 2^32");
  spf_ctx* c = ds->ctx;
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const uint32_t ld = ds->ld;

  DevBuf<uint64_t> d_crow, d_pidx;
  DevBuf<int> d_flag;
  SPF_TRY(d_flag.alloc(st, 1));
  SPF_CUDA(cudaMemsetAsync(d_flag.p, 0, sizeof(int), st));
  if (centroid_rows) {
    SPF_TRY(d_crow.alloc(st, k));
    SPF_CUDA(cudaMemcpyAsync(d_crow.p, centroid_rows, (size_t)k * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    SPF_TRY(launch_check_rows(c, d_crow.p, k, ds->n, d_flag.p));
  }
  if (point_idx) {
    SPF_TRY(d_pidx.alloc(st, m));
    SPF_CUDA(cudaMemcpyAsync(d_pidx.p, point_idx, (size_t)m * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    SPF_TRY(launch_check_rows(c, d_pidx.p, m, ds->n, d_flag.p));
  }
  int h_flag = 0;
  SPF_CUDA(cudaMemcpyAsync(&h_flag, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  if (h_flag) return fail(SPF_E_INVALID, "a point or centroid row index is >= n (%llu)", (unsigned long long)ds->n);

  AssignCall a;
  a.c = c; a.metric = metric; a.k = k; a.ld = ld; a.m = m;
  a.want_members = !(flags & SPF_ASSIGN_NO_CSR);
  a.factor = a.want_members ? boundary_factor : 1.0f;
  a.use_tc = metric == SPF_METRIC_EUCLIDEAN && !(flags & SPF_ASSIGN_FORCE_EXACT) && !c->params.force_exact &&
             assign_tc_supported(c, m, k, ld);
  a.chunk_rows = pick_chunk_rows(c, m, false, effective_cand_cap(c, a.use_tc, ld));
  a.seed = d_seed;
  a.penalty = d_penalty;
  if (d_penalty) a.factor = 1.0f;             // balanced assignment: exactly one cluster per point

  // dense operands: centroids always materialised; points gathered only for a subset
  SPF_TRY(a.Cg.alloc(st, (size_t)k * ld));
  if (centroid_rows) {
    SPF_TRY(launch_gather_rows(c, ds->x, ld, d_crow.p, k, a.Cg.p));
  } else if (centroid_dev) {
    SPF_CUDA(cudaMemcpyAsync(a.Cg.p, centroid_dev, (size_t)k * ld * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else {
    if (ld != ds->d) SPF_CUDA(cudaMemsetAsync(a.Cg.p, 0, (size_t)k * ld * sizeof(float), st));
    SPF_CUDA(cudaMemcpy2DAsync(a.Cg.p, (size_t)ld * sizeof(float), centroid_vecs, (size_t)ds->d * sizeof(float),
                               (size_t)ds->d * sizeof(float), k, cudaMemcpyHostToDevice, st));
    SPF_CUDA(cudaStreamSynchronize(st));
  }
  DevBuf<float> Pg, ptf_sub, xnorm_sub, xres_sub;
  const float* P = ds->x;
  const float *Ptf = nullptr, *xnorm = nullptr, *xres = nullptr;
  if (point_idx) {
    SPF_TRY(Pg.alloc(st, (size_t)m * ld));
    SPF_TRY(launch_gather_rows(c, ds->x, ld, d_pidx.p, m, Pg.p));
    P = Pg.p;
  }
  if (a.use_tc) {
    if (point_idx) {
      SPF_TRY(ptf_sub.alloc(st, (size_t)m * ld));
      SPF_TRY(xnorm_sub.alloc(st, m));
      SPF_TRY(xres_sub.alloc(st, m));
      SPF_TRY(launch_row_prep(c, P, ld, nullptr, m, ptf_sub.p, xnorm_sub.p, xres_sub.p));
      Ptf = ptf_sub.p; xnorm = xnorm_sub.p; xres = xres_sub.p;
    } else {
      SPF_TRY(dataset_prep(ds));
      Ptf = ds->xtf; xnorm = ds->xnorm; xres = ds->xres;
    }
  }
  // Manhattan / Chebyshev on long rows: distances concentrate, so until the kernel meets a centroid of
  // the point's own neighbourhood nearly every centroid lies inside the 1.1 band of the running minimum
  // (1M x 960 clustered rows: 60 % of the points overflowed their 512 candidate records and were
  // recomputed by the dense fallback).  The tensor cores seed the running minimum: one squared-L2
  // assign (TF32 GEMM, nearest centroid only) names a near centroid per point, its exact distance under
  // the requested metric is an upper bound of the minimum, and the direct-form kernel starts with it.
  DevBuf<float> seed_own;
  if (!a.use_tc && !a.seed && !a.penalty && metric != SPF_METRIC_EUCLIDEAN && a.want_members && c->params.exact_seed != 0 &&
      (c->params.exact_seed > 1 || (ld >= 256 && (uint64_t)m * k >= (1ull << 28))) && k >= 256 &&
      assign_tc_supported(c, m, k, ld) && !c->params.force_exact) {
    KernelTimer t(c, "exact_seed");
    spf_assign_result* pre = nullptr;
    int rc = assign_resident(ds, SPF_METRIC_EUCLIDEAN, point_idx, m, nullptr, nullptr, a.Cg.p, k, 1.0f, SPF_ASSIGN_NO_CSR,
                             nullptr, nullptr, &pre);
    if (rc >= 0) rc = seed_own.alloc(st, m);
    if (rc >= 0)
      rc = launch_pair_dist(c, metric, P, ld, nullptr, a.Cg.p, ld, pre->best, UINT64_MAX, ld, m, seed_own.p);
    if (pre) spf_assign_free(pre);
    if (rc >= 0) a.seed = seed_own.p;
    else cudaGetLastError();                  // e.g. no room for the TF32 copy of the rows: run unseeded
  }
  SPF_TRY(assign_setup(a));
  for (uint64_t r0 = 0; r0 < m; r0 += a.chunk_rows) {
    const uint64_t mc = (m - r0) < a.chunk_rows ? (m - r0) : a.chunk_rows;
    SPF_TRY(assign_rows(a, P + r0 * ld, Ptf ? Ptf + r0 * ld : nullptr, xnorm ? xnorm + r0 : nullptr,
                        xres ? xres + r0 : nullptr, r0, mc));
  }
  return assign_finish(a, P, xnorm, xres, &d_pidx, out);
}

int spf::assign_device_centroids(spf_dataset* ds, int metric, const float* d_centroids, uint32_t k,
                                 float boundary_factor, int flags, const float* d_seed, const float* d_penalty,
                                 spf_assign_result** out) {
  return assign_resident(ds, metric, nullptr, ds ? ds->n : 0, nullptr, nullptr, d_centroids, k, boundary_factor, flags,
                         d_seed, d_penalty, out);
}

extern "C" {

int spf_assign(spf_dataset* ds, int metric, const uint64_t* point_idx, uint64_t m,
               const uint64_t* centroid_rows, uint32_t k, float boundary_factor, int flags,
               spf_assign_result** out) {
  return spf::guarded([&]() -> int {
  if (!ds || !centroid_rows) return fail(SPF_E_INVALID, "spf_assign: NULL argument");
  std::lock_guard<std::mutex> lk(ds->ctx->mu);
  ds->ctx->kernel_ms.clear();
  return assign_resident(ds, metric, point_idx, m, centroid_rows, nullptr, nullptr, k, boundary_factor, flags, nullptr, nullptr, out);
  });
}

int spf_assign_vectors(spf_dataset* ds, int metric, const uint64_t* point_idx, uint64_t m,
                       const float* centroids, uint32_t k, float boundary_factor, int flags,
                       spf_assign_result** out) {
  return spf::guarded([&]() -> int {
  if (!ds || !centroids) return fail(SPF_E_INVALID, "spf_assign_vectors: NULL argument");
  std::lock_guard<std::mutex> lk(ds->ctx->mu);
  ds->ctx->kernel_ms.clear();
  return assign_resident(ds, metric, point_idx, m, nullptr, centroids, nullptr, k, boundary_factor, flags, nullptr, nullptr, out);
  });
}

int spf_assign_balanced(spf_dataset* ds, int metric, const uint64_t* point_idx, uint64_t m, const float* centroids,
                        const float* penalty, uint32_t k, int flags, spf_assign_result** out) {
  return spf::guarded([&]() -> int {
    if (!ds || !centroids || !out) return fail(SPF_E_INVALID, "spf_assign_balanced: NULL argument");
    if (k == 0) return fail(SPF_E_INVALID, "k must be > 0");
    spf_ctx* c = ds->ctx;
    std::lock_guard<std::mutex> lk(c->mu);
    c->kernel_ms.clear();
    SPF_CUDA(cudaSetDevice(c->device));
    DevBuf<float> d_pen;
    SPF_TRY(d_pen.alloc(c->stream, k));
    if (penalty) {
      for (uint32_t j = 0; j < k; ++j)
        if (!(penalty[j] >= 0.0f)) return fail(SPF_E_INVALID, "penalty[%u] must be a non-negative number", j);
      SPF_CUDA(cudaMemcpyAsync(d_pen.p, penalty, (size_t)k * sizeof(float), cudaMemcpyHostToDevice, c->stream));
      SPF_CUDA(cudaStreamSynchronize(c->stream));
    } else {
      SPF_CUDA(cudaMemsetAsync(d_pen.p, 0, (size_t)k * sizeof(float), c->stream));
    }
    return assign_resident(ds, metric, point_idx, m, nullptr, centroids, nullptr, k, 1.0f, flags & ~SPF_ASSIGN_NO_CSR, nullptr,
                           d_pen.p, out);
  });
}

int spf_assign_host(spf_ctx* c, const float* rows, uint64_t n, uint32_t d, uint64_t row_stride, int metric,
                    const uint64_t* centroid_rows, uint32_t k, float boundary_factor, int flags,
                    spf_dataset** ds_out, spf_assign_result** out) {
  return spf::guarded([&]() -> int {
  if (!c || !rows || !out || !centroid_rows) return fail(SPF_E_INVALID, "spf_assign_host: NULL argument");
  *out = nullptr;
  if (ds_out) *ds_out = nullptr;
  if (metric < 0 || metric > 2) return fail(SPF_E_INVALID, "unknown metric %d", metric);
  if (k == 0) return fail(SPF_E_INVALID, "k must be > 0 (the reference indexes centroids[0])");
  if (k > (REC_G_MASK << 2)) return fail(SPF_E_INVALID, "k must be < 2^30");
  if (row_stride < d) return fail(SPF_E_INVALID, "row_stride (%llu) < d (%u)", (unsigned long long)row_stride, d);
  for (uint32_t j = 0; j < k; ++j)
    if (centroid_rows[j] >= n)
      return fail(SPF_E_INVALID, "a point or centroid row index is >= n (%llu)", (unsigned long long)n);
  spf_dataset* ds = nullptr;
  SPF_TRY(dataset_alloc(c, n, d, &ds));
  struct Guard {   // frees the dataset unless it is handed to the caller
    spf_dataset* d;
    ~Guard() { if (d) spf_dataset_free(d); }
  } guard{ds};
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  c->kernel_ms.clear();
  const uint32_t ld = ds->ld;
  if (!c->copy_stream) SPF_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));

  AssignCall a;
  a.c = c; a.metric = metric; a.k = k; a.ld = ld; a.m = n;
  a.want_members = !(flags & SPF_ASSIGN_NO_CSR);
  a.factor = a.want_members ? boundary_factor : 1.0f;
  a.use_tc = metric == SPF_METRIC_EUCLIDEAN && !(flags & SPF_ASSIGN_FORCE_EXACT) && !c->params.force_exact &&
             assign_tc_supported(c, n, k, ld);
  a.chunk_rows = pick_chunk_rows(c, n, true, effective_cand_cap(c, a.use_tc, ld));
  if (a.use_tc) SPF_TRY(dataset_prep_alloc(ds));

  // the k centroid vectors come straight from the host rows (small gather + one copy)
  {
    std::vector<float> cg((size_t)k * ld, 0.0f);
    for (uint32_t j = 0; j < k; ++j)
      memcpy(&cg[(size_t)j * ld], rows + (size_t)centroid_rows[j] * row_stride, (size_t)d * sizeof(float));
    SPF_TRY(a.Cg.alloc(st, (size_t)k * ld));
    SPF_CUDA(cudaMemcpyAsync(a.Cg.p, cg.data(), cg.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    SPF_CUDA(cudaStreamSynchronize(st));   // cg goes out of scope
  }
  SPF_TRY(assign_setup(a));

  // upload chunk by chunk on the copy stream; the compute stream follows one event behind
  const uint64_t nchunks = ceil_div(n, a.chunk_rows);
  std::vector<cudaEvent_t> evs(nchunks, nullptr);
  struct EvGuard {
    std::vector<cudaEvent_t>& v;
    ~EvGuard() { for (cudaEvent_t e : v) if (e) cudaEventDestroy(e); }
  } evguard{evs};
  // the copy stream must not overwrite memory the pool may still be recycling on the main stream
  {
    cudaEvent_t e0;
    SPF_CUDA(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
    cudaEventRecord(e0, st);
    cudaStreamWaitEvent(c->copy_stream, e0, 0);
    cudaEventDestroy(e0);
  }
  auto compute_chunk = [&](uint64_t ci) -> int {          // the chunk's kernels follow its upload by one event
    const uint64_t r0 = ci * a.chunk_rows;
    const uint64_t mc = (n - r0) < a.chunk_rows ? (n - r0) : a.chunk_rows;
    SPF_CUDA(cudaEventCreateWithFlags(&evs[ci], cudaEventDisableTiming));
    SPF_CUDA(cudaEventRecord(evs[ci], c->copy_stream));
    SPF_CUDA(cudaStreamWaitEvent(st, evs[ci], 0));
    if (a.use_tc)
      SPF_TRY(launch_row_prep(c, ds->x + r0 * ld, ld, nullptr, mc, ds->xtf + r0 * ld, ds->xnorm + r0, ds->xres + r0));
    return assign_rows(a, ds->x + r0 * ld, a.use_tc ? ds->xtf + r0 * ld : nullptr,
                       a.use_tc ? ds->xnorm + r0 : nullptr, a.use_tc ? ds->xres + r0 : nullptr, r0, mc);
  };
  if (host_pointer_is_pageable(rows) && !c->params.no_host_staging) {
    // ordinary heap memory (what the reference's callers hold): threaded staging through pinned slots
    SPF_TRY(staged_upload(c, rows, n, d, row_stride, ds->x, ld, a.chunk_rows, c->copy_stream, compute_chunk));
  } else {
    for (uint64_t ci = 0; ci < nchunks; ++ci) {
      const uint64_t r0 = ci * a.chunk_rows;
      const uint64_t mc = (n - r0) < a.chunk_rows ? (n - r0) : a.chunk_rows;
      float* dst = ds->x + r0 * ld;
      const float* src = rows + r0 * row_stride;
      if (ld == d && row_stride == d) {
        SPF_CUDA(cudaMemcpyAsync(dst, src, (size_t)mc * d * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
      } else {
        if (ld != d) SPF_CUDA(cudaMemsetAsync(dst, 0, (size_t)mc * ld * sizeof(float), c->copy_stream));
        SPF_CUDA(cudaMemcpy2DAsync(dst, (size_t)ld * sizeof(float), src, (size_t)row_stride * sizeof(float),
                                   (size_t)d * sizeof(float), (size_t)mc, cudaMemcpyHostToDevice, c->copy_stream));
      }
      SPF_TRY(compute_chunk(ci));
    }
  }
  SPF_CUDA(cudaStreamSynchronize(c->copy_stream));   // `rows` is not read after the call returns
  SPF_TRY(assign_finish(a, ds->x, ds->xnorm, ds->xres, nullptr, out));
  if (ds_out) {
    *ds_out = ds;
    guard.d = nullptr;
  }
  return SPF_OK;
  });
}

uint64_t spf_assign_points(const spf_assign_result* r) { return r ? r->m : 0; }
uint32_t spf_assign_clusters(const spf_assign_result* r) { return r ? r->k : 0; }
uint64_t spf_assign_total(const spf_assign_result* r) { return r ? r->total : 0; }

int spf_assign_fetch(const spf_assign_result* r, uint32_t* best, float* dmin, uint64_t* offsets,
                     uint64_t* members) {
  return spf::guarded([&]() -> int {
  if (!r) return fail(SPF_E_INVALID, "result is NULL");
  spf_ctx* c = r->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  // pinned destinations: plain asynchronous copies; ordinary heap memory: through the staging ring
  auto d2h = [&](void* host, const void* dev, size_t bytes) -> int {
    if (bytes >= (8u << 20) && host_pointer_is_pageable(host) && !c->params.no_host_staging) {
      SPF_CUDA(cudaStreamSynchronize(st));                 // the ring's events are shared: drain what is queued first
      return staged_download(c, host, dev, bytes, st);
    }
    SPF_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st));
    return SPF_OK;
  };
  if (best) SPF_TRY(d2h(best, r->best, r->m * sizeof(uint32_t)));
  if (dmin) SPF_TRY(d2h(dmin, r->dmin, r->m * sizeof(float)));
  if ((offsets || members) && !r->has_csr)
    return fail(SPF_E_STATE, "the result was computed with SPF_ASSIGN_NO_CSR");
  if (offsets)
    SPF_CUDA(cudaMemcpyAsync(offsets, r->offsets, ((size_t)r->k + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  DevBuf<uint64_t> rows;
  if (members && r->total) {
    SPF_TRY(rows.alloc(st, r->total));
    SPF_TRY(assign_members_as_rows(r, rows.p));
    SPF_TRY(d2h(members, rows.p, r->total * sizeof(uint64_t)));
  }
  SPF_CUDA(cudaStreamSynchronize(st));
  return SPF_OK;
  });
}

void spf_assign_free(spf_assign_result* r) {
  if (!r) return;
  cudaSetDevice(r->ctx->device);
  cudaStream_t st = r->ctx->stream;
  if (r->best) cudaFreeAsync(r->best, st);
  if (r->dmin) cudaFreeAsync(r->dmin, st);
  if (r->offsets) cudaFreeAsync(r->offsets, st);
  if (r->members) cudaFreeAsync(r->members, st);
  if (r->point_idx) cudaFreeAsync(r->point_idx, st);
  delete r;
}

}  // extern "C"
