// assign_api.cu — spf_assign: the batched replacement of
// HierarchicalClustering::assign_points_to_clusters (src/clustering/hierarchical.rs:295-364).
#include "kernels.cuh"

using namespace spf;


namespace spf {

namespace {

__global__ void positions_to_rows_kernel(const uint32_t* __restrict__ pos, const uint64_t* __restrict__ pidx,
                                         uint64_t total, uint64_t* __restrict__ out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < total) out[t] = pidx ? pidx[pos[t]] : (uint64_t)pos[t];
}

__global__ void pad_inf_kernel(float* p, uint32_t from, uint32_t to) {
  const uint32_t t = from + blockIdx.x * blockDim.x + threadIdx.x;
  if (t < to) p[t] = __int_as_float(0x7f800000);
}

}  // namespace

int assign_members_as_rows(const spf_assign_result* r, uint64_t* d_out) {
  if (r->total == 0) return SPF_OK;
  positions_to_rows_kernel<<<(unsigned)ceil_div(r->total, 256), 256, 0, r->ctx->stream>>>(
      r->members, r->point_idx, r->total, d_out);
  return check_launch(r->ctx, "positions_to_rows_kernel");
}

// Rounded copy, squared norms and rounding residuals of all dataset rows, computed once per
// dataset (tensor path only).
int dataset_prep(spf_dataset* ds) {
  if (ds->xtf) return SPF_OK;
  spf_ctx* c = ds->ctx;
  float *tf = nullptr, *nrm = nullptr, *res = nullptr;
  cudaError_t e = cudaMalloc((void**)&tf, (size_t)ds->n * ds->ld * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&nrm, (size_t)ds->n * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&res, (size_t)ds->n * sizeof(float));
  int rc = e == cudaSuccess ? launch_row_prep(c, ds->x, ds->ld, nullptr, ds->n, tf, nrm, res)
                            : fail(SPF_E_OOM, "cudaMalloc for the rounded dataset copy failed: %s", cudaGetErrorString(e));
  if (rc < 0) {
    cudaFree(tf); cudaFree(nrm); cudaFree(res);
    return rc;
  }
  ds->xtf = tf;
  ds->xnorm = nrm;
  ds->xres = res;
  return SPF_OK;
}

}  // namespace spf

extern "C" {

int spf_assign(spf_dataset* ds, int metric, const uint64_t* point_idx, uint64_t m,
               const uint64_t* centroid_rows, uint32_t k, float boundary_factor, int flags,
               spf_assign_result** out) {
  if (!ds || !out || !centroid_rows) return fail(SPF_E_INVALID, "spf_assign: NULL argument");
  *out = nullptr;
  if (metric < 0 || metric > 2) return fail(SPF_E_INVALID, "unknown metric %d", metric);
  if (k == 0) return fail(SPF_E_INVALID, "k must be > 0 (the reference indexes centroids[0])");
  if (k > (REC_G_MASK << 2)) return fail(SPF_E_INVALID, "k must be < 2^30");
  if (!point_idx && m != ds->n) return fail(SPF_E_INVALID, "point_idx == NULL requires m == n");
  if (m == 0) return fail(SPF_E_INVALID, "m must be > 0");
  if (m >= (1ull << 32)) return fail(SPF_E_INVALID, "m must be < 2^32");
  spf_ctx* c = ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const uint32_t ld = ds->ld;
  const bool want_members = !(flags & SPF_ASSIGN_NO_CSR);
  const float factor = want_members ? boundary_factor : 1.0f;

  DevBuf<uint64_t> d_crow, d_pidx;
  DevBuf<int> d_flag;
  SPF_TRY(d_crow.alloc(st, k));
  SPF_TRY(d_flag.alloc(st, 1));
  SPF_CUDA(cudaMemsetAsync(d_flag.p, 0, sizeof(int), st));
  SPF_CUDA(cudaMemcpyAsync(d_crow.p, centroid_rows, (size_t)k * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  SPF_TRY(launch_check_rows(c, d_crow.p, k, ds->n, d_flag.p));
  if (point_idx) {
    SPF_TRY(d_pidx.alloc(st, m));
    SPF_CUDA(cudaMemcpyAsync(d_pidx.p, point_idx, (size_t)m * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    SPF_TRY(launch_check_rows(c, d_pidx.p, m, ds->n, d_flag.p));
  }
  int h_flag = 0;
  SPF_CUDA(cudaMemcpyAsync(&h_flag, d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  if (h_flag) return fail(SPF_E_INVALID, "a point or centroid row index is >= n (%llu)", (unsigned long long)ds->n);

  // dense operands: centroids always gathered; points gathered only for a subset
  DevBuf<float> Cg, Pg;
  SPF_TRY(Cg.alloc(st, (size_t)k * ld));
  SPF_TRY(launch_gather_rows(c, ds->x, ld, d_crow.p, k, Cg.p));
  const float* P = ds->x;
  if (point_idx) {
    SPF_TRY(Pg.alloc(st, (size_t)m * ld));
    SPF_TRY(launch_gather_rows(c, ds->x, ld, d_pidx.p, m, Pg.p));
    P = Pg.p;
  }

  CandBuf cand;
  cand.cap = c->params.cand_cap;
  DevBuf<CandRec> cand_rec;
  DevBuf<RowInfo> cand_info;
  SPF_TRY(cand_rec.alloc(st, (size_t)m * cand.cap));
  SPF_TRY(cand_info.alloc(st, m));
  cand.rec = cand_rec.p;
  cand.info = cand_info.p;

  const bool use_tc = metric == SPF_METRIC_EUCLIDEAN && !(flags & SPF_ASSIGN_FORCE_EXACT) &&
                      !c->params.force_exact && assign_tc_supported(c, m, k, ld);
  DevBuf<float> ptf_sub, xnorm_sub, xres_sub, ctf, cnorm, cres, cstat;
  const float* xnorm = nullptr;
  const float* xres = nullptr;
  if (use_tc) {
    const float* Ptf = nullptr;
    if (point_idx) {   // gather + round in one pass (P itself is only needed by resolve)
      SPF_TRY(ptf_sub.alloc(st, (size_t)m * ld));
      SPF_TRY(xnorm_sub.alloc(st, m));
      SPF_TRY(xres_sub.alloc(st, m));
      SPF_TRY(launch_row_prep(c, P, ld, nullptr, m, ptf_sub.p, xnorm_sub.p, xres_sub.p));
      Ptf = ptf_sub.p; xnorm = xnorm_sub.p; xres = xres_sub.p;
    } else {
      SPF_TRY(dataset_prep(ds));
      Ptf = ds->xtf; xnorm = ds->xnorm; xres = ds->xres;
    }
    const uint32_t kpad = round_up(k, 256);
    SPF_TRY(ctf.alloc(st, (size_t)k * ld));
    SPF_TRY(cnorm.alloc(st, kpad));
    SPF_TRY(cres.alloc(st, k));
    SPF_TRY(cstat.alloc(st, 2));
    SPF_TRY(launch_row_prep(c, Cg.p, ld, nullptr, k, ctf.p, cnorm.p, cres.p));
    SPF_TRY(launch_max2_f32(c, cnorm.p, cres.p, k, cstat.p));
    if (kpad > k) {
      pad_inf_kernel<<<(kpad - k + 255) / 256, 256, 0, st>>>(cnorm.p, k, kpad);
      SPF_TRY(check_launch(c, "pad_inf_kernel"));
    }
    KernelTimer t(c, "assign_tc");
    SPF_TRY(launch_assign_tc(c, Ptf, m, ctf.p, k, ld, xnorm, xres, cnorm.p, cstat.p, factor, cand));
  } else {
    KernelTimer t(c, "assign_exact");
    SPF_TRY(launch_assign_exact(c, metric, P, m, Cg.p, k, ld, factor, &cand, nullptr));
  }

  // exact centroid-centroid distances for the boundary rule `d(c_best, c_j) >= d_j` (:337-342)
  DevBuf<float> cc;
  if (want_members && k > 1 && (int)k <= c->params.cc_matrix_max_k) {
    SPF_TRY(cc.alloc(st, (size_t)k * k));
    KernelTimer t(c, "cc_matrix");
    SPF_TRY(launch_assign_exact(c, metric, Cg.p, k, Cg.p, k, ld, 1.0f, nullptr, cc.p));
  }

  spf_assign_result* r = new (std::nothrow) spf_assign_result();
  if (!r) return fail(SPF_E_OOM, "out of host memory");
  r->ctx = c;
  r->m = m;
  r->k = k;
  DevBuf<uint32_t> best, nmem;
  DevBuf<float> dmin;
  int rc = best.alloc(st, m);
  if (rc >= 0) rc = dmin.alloc(st, m);
  if (rc >= 0) rc = nmem.alloc(st, m);
  CsrOut csr;
  if (rc >= 0) {
    ResolveArgs a;
    a.metric = metric; a.P = P; a.m = m; a.C = Cg.p; a.k = k; a.ld = ld; a.factor = factor;
    a.cand = cand; a.nseg = use_tc ? 2 : 1;
    a.xnorm = use_tc ? xnorm : nullptr; a.xres = use_tc ? xres : nullptr;
    a.d_cstat = use_tc ? cstat.p : nullptr;
    a.cc = cc.p; a.want_members = want_members;
    a.best = best.p; a.dmin = dmin.p; a.nmem = nmem.p;
    rc = run_resolve(c, a, want_members ? &csr : nullptr);
  }
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
  r->best = best.take();
  r->dmin = dmin.take();
  r->has_csr = want_members;
  r->total = csr.total;
  r->offsets = csr.offsets;
  r->members = csr.members;
  r->point_idx = point_idx ? d_pidx.take() : nullptr;
  *out = r;
  return SPF_OK;
}

uint64_t spf_assign_points(const spf_assign_result* r) { return r ? r->m : 0; }
uint32_t spf_assign_clusters(const spf_assign_result* r) { return r ? r->k : 0; }
uint64_t spf_assign_total(const spf_assign_result* r) { return r ? r->total : 0; }

int spf_assign_fetch(const spf_assign_result* r, uint32_t* best, float* dmin, uint64_t* offsets,
                     uint64_t* members) {
  if (!r) return fail(SPF_E_INVALID, "result is NULL");
  spf_ctx* c = r->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  if (best) SPF_CUDA(cudaMemcpyAsync(best, r->best, r->m * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  if (dmin) SPF_CUDA(cudaMemcpyAsync(dmin, r->dmin, r->m * sizeof(float), cudaMemcpyDeviceToHost, st));
  if ((offsets || members) && !r->has_csr)
    return fail(SPF_E_STATE, "the result was computed with SPF_ASSIGN_NO_CSR");
  if (offsets)
    SPF_CUDA(cudaMemcpyAsync(offsets, r->offsets, ((size_t)r->k + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  DevBuf<uint64_t> rows;
  if (members && r->total) {
    SPF_TRY(rows.alloc(st, r->total));
    SPF_TRY(assign_members_as_rows(r, rows.p));
    SPF_CUDA(cudaMemcpyAsync(members, rows.p, r->total * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  }
  SPF_CUDA(cudaStreamSynchronize(st));
  return SPF_OK;
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
