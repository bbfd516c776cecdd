// kernels.cuh — internal launch interface between the translation units of libspfresh_b200.
#pragma once
#include "common.cuh"

namespace spf {

// Candidate record written by the assign kernels: x = centroid slot (bit 31 set once the
// distance has been recomputed exactly), y = distance bits.
static constexpr uint32_t CAND_EXACT_BIT = 0x80000000u;
static constexpr uint32_t CAND_MEMBER_BIT = 0x40000000u;
static constexpr uint32_t CAND_SLOT_MASK = 0x3fffffffu;
static constexpr uint32_t NMEM_OVERFLOW_BIT = 0x80000000u;

// Certified bound on |d_tf32 - d_ref| for the tensor path, where d_tf32 = |x|^2 - 2 x.c + |c|^2
// with the dot product taken on TF32 operands and d_ref is the reference's sequential f32 sum.
//  * TF32 operands are off by < 2^-10 relative each, so every product by < 2^-9 (1 + 2^-11);
//    Cauchy-Schwarz bounds the dot-product error by 2^-9 |x||c| and the distance error by twice
//    that; 10 % head-room covers the fp32 accumulation inside the tensor core.
//  * d_ref itself is within (ld + 2) 2^-24 D of the real distance D <= 2 (|x|^2 + |c|^2); the
//    fp32 norms and epilogue adds contribute a few 2^-24 (|x|^2 + |c|^2) more.
// Used identically by the GEMM epilogue (candidate slack) and by resolve (decision bands).
__host__ __device__ inline float tc_err_bound(float xn, float cnmax, uint32_t ld) {
  return 1.1f * 0.00390625f * sqrtf(xn * cnmax) + (float)(ld + 16) * 1.1920929e-7f * (xn + cnmax);
}

// ---- support.cu --------------------------------------------------------------------------
int launch_gather_rows(spf_ctx* c, const float* src, uint32_t ld, const uint64_t* d_idx, uint64_t m,
                       float* dst);
int launch_row_sqnorm(spf_ctx* c, const float* rows, uint32_t ld, uint64_t m, float* out);
int launch_fill_f32(spf_ctx* c, float* p, uint64_t n, float v);
int launch_fill_u64(spf_ctx* c, uint64_t* p, uint64_t n, uint64_t v);
int launch_max_f32(spf_ctx* c, const float* p, uint64_t n, float* out1);   // out1[0] = max
// dist[i] = metric(A row ai, B row bi): ai = idxA ? idxA[i] : i ;
// bi = idxB32 ? idxB32[i] : (fixedB == UINT64_MAX ? i : fixedB)
int launch_pair_dist(spf_ctx* c, int metric, const float* A, uint32_t ldA, const uint64_t* idxA,
                     const float* B, uint32_t ldB, const uint32_t* idxB32, uint64_t fixedB,
                     uint32_t ld, uint64_t count, float* out);
int launch_check_rows(spf_ctx* c, const uint64_t* d_idx, uint64_t m, uint64_t n, int* d_flag);

// ---- assign_exact.cu ----------------------------------------------------------------------
// CUDA-core direct-form kernel: every distance of the m x k problem, exact.  Emits boundary
// candidates (cand != NULL) and/or the dense m x k matrix (dense != NULL).
int launch_assign_exact(spf_ctx* c, int metric, const float* P, uint64_t m, const float* C, uint32_t k,
                        uint32_t ld, float factor, uint2* cand, uint32_t* cand_cnt, int cap, float* dense);

// ---- assign_tc.cu -------------------------------------------------------------------------
// tcgen05 (TF32) candidate GEMM for squared-Euclidean: approximate distances with a certified
// error bound, candidates only.  cnorm_pad has round_up(k,256) entries (+inf padding).
bool assign_tc_supported(const spf_ctx* c, uint64_t m, uint32_t k, uint32_t ld);
int launch_assign_tc(spf_ctx* c, const float* P, uint64_t m, const float* C, uint32_t k, uint32_t ld,
                     const float* xnorm, const float* cnorm_pad, const float* d_cnmax, float factor,
                     uint2* cand, uint32_t* cand_cnt, int cap);

// ---- resolve.cu ---------------------------------------------------------------------------
struct ResolveArgs {
  int metric;
  const float* P; uint64_t m; const float* C; uint32_t k; uint32_t ld;
  float factor;
  uint2* cand; uint32_t* cand_cnt; int cap;
  int nseg;                // segments per row buffer: 1 (exact kernel) or 2 (tensor kernel)
  const float* xnorm;      // NULL on the exact path (error bound 0)
  const float* d_cnmax;    // device scalar, tensor path only
  const float* cc;         // k x k exact centroid-centroid distances or NULL (computed on demand)
  bool want_members;
  // outputs
  uint32_t* best; float* dmin; uint32_t* nmem;   // m each
};
// Cluster-major CSR built from the per-row member lists (stable in input order).
struct CsrOut {
  uint64_t total = 0;
  uint64_t* offsets = nullptr;   // device, k+1
  uint32_t* members = nullptr;   // device, total: positions 0..m-1 into the assign's point list
};
// Resolves best/dmin/members per row; csr == NULL skips the CSR build.
int run_resolve(spf_ctx* c, const ResolveArgs& a, CsrOut* csr);

// ---- assign_api.cu ------------------------------------------------------------------------
int dataset_norms(spf_dataset* ds);
int assign_members_as_rows(const spf_assign_result* r, uint64_t* d_out);   // positions → dataset rows

}  // namespace spf
