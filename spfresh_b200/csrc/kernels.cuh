// kernels.cuh — internal launch interface between the translation units of libspfresh_b200.
#pragma once
#include <functional>

#include "common.cuh"

struct spf_comm;

namespace spf {

// Candidate records written by the assign kernels.  A record (32 bytes, one HBM sector) covers the
// four consecutive centroid slots 4*g .. 4*g+3 of one point ("group record"): `t` holds one value
// per slot, `g` holds the group index (bits 0-27) and one "exact" flag per slot (bits 28-31).
//   exact flag set   : the value is the exact direct-form distance d(x, c)
//   exact flag clear : the value is s = x.c - |c|^2/2 from the TF32 GEMM; d ~ |x|^2 - 2 s within E
// Slots >= k carry no candidate (they are skipped by index).  A point owns `cap` records, split
// into `nseg` segments (one per writer thread of the producing kernel).  After resolve the front
// of the point's record area is reused for its member list (uint32 centroid slots).
struct __align__(16) CandRec {
  float4 t;
  uint32_t g;
  uint32_t pad[3];   // written as zeros by the tensor kernel (one 256-bit store per record: whole sector, no DRAM read-fill)
};
static constexpr uint32_t REC_G_MASK = 0x0fffffffu;
static constexpr int REC_EXACT_SHIFT = 28;
static constexpr uint32_t REC_ALL_EXACT = 0xf0000000u;
static constexpr uint32_t NMEM_OVERFLOW_BIT = 0x80000000u;
// Per point: x = records in segment 0 (a value > segment capacity marks an overflow), y = bits of
// the best value the producer saw in segment 0's columns (tensor kernel: largest s; exact kernel:
// smallest distance), z / w = the same for segment 1.
typedef uint4 RowInfo;

// Certified bound on |d_tf32 - d_ref| for the tensor path.  d_tf32 = |x|^2 - 2 x'.c' + |c|^2 where
// x' = rn_tf32(x), c' = rn_tf32(c) are the operands the GEMM reads (rounded copies made by
// row_prep_kernel, so the hardware's own treatment of the low mantissa bits never matters) and
// d_ref is the reference's sequential f32 sum.
//  * operands: x.c - x'.c' = dx.c + x'.dc with dx = x - x', dc = c - c'; by Cauchy-Schwarz
//    |.| <= |dx| |c| + (|x| + |dx|) |dc|.  |dx| is known per point (xres), |c| and |dc| are
//    bounded by their maxima over the centroids.  The distance error is twice that.
//  * products of two TF32 values are exact in fp32; the fp32 accumulation inside the tensor core
//    (16 MMA steps of 8 products, alignment truncation included) stays below
//    (ld + 16) 2^-23 |x||c| for the dot product, i.e. (ld + 16) 2^-23 (|x|^2 + |c|^2) for the distance.
//  * d_ref itself is within (ld + 2) 2^-24 D of the real distance, D <= 2 (|x|^2 + |c|^2); the
//    fp32 norms and the epilogue adds contribute a few 2^-24 (|x|^2 + |c|^2) more.
// Used identically by the GEMM epilogue (candidate slack) and by resolve (decision bands).
__host__ __device__ inline float tc_err_bound(float xn, float xres, float cnmax, float dcmax, uint32_t ld) {
  const float sx = sqrtf(xn), sc = sqrtf(cnmax);
  const float op = 2.0f * (xres * sc + (sx + xres) * dcmax);
  return 1.01f * op + 3.0f * (float)(ld + 16) * 1.1920929e-7f * (xn + cnmax);
}

// Seeded candidate pass: `seed` is an upper bound of the point's minimum reference distance, so some
// centroid has d_tf32 <= seed + E and the final maximum of s = x.c - |c|^2/2 is at least this value.
// Evaluated with the same operations by the GEMM epilogue (initial bound) and by resolve (which
// checks that the observed maximum reaches it; otherwise the seed was not a valid bound and the
// point goes to the dense fallback).
__device__ __forceinline__ float tc_seed_bound(float xn, float seed, float E, float cnmax) {
  const float slop = __fadd_rn(__fmul_rn(1e-6f, __fadd_rn(xn, cnmax)), 1e-30f);
  return __fsub_rn(__fmul_rn(0.5f, __fsub_rn(__fsub_rn(xn, seed), E)), slop);
}

// Candidate scratch of one assign call (device).
struct CandBuf {
  CandRec* rec = nullptr;    // m * cap records
  RowInfo* info = nullptr;   // m
  int cap = 0;               // records per point (all segments together)
};

// Rounded operands + norms for the tensor path: for every row r of `rows`
//   tf[r]    = rn_tf32(rows[r])              (cvt.rna.tf32.f32, element-wise)
//   norm[r]  = |rows[r]|^2                   (fp32, any order)
//   res[r]   = |rows[r] - tf[r]|             (fp32)
// idx == NULL: rows 0..m-1 of src; otherwise row idx[r] of src (gather fused in).
int launch_row_prep(spf_ctx* c, const float* src, uint32_t ld, const uint64_t* d_idx, uint64_t m,
                    float* tf, float* norm, float* res);
// out2[0] = max(a[0..n)), out2[1] = max(b[0..n))   (values >= 0)
int launch_max2_f32(spf_ctx* c, const float* a, const float* b, uint64_t n, float* out2);

// ---- support.cu --------------------------------------------------------------------------
int launch_gather_rows(spf_ctx* c, const float* src, uint32_t ld, const uint64_t* d_idx, uint64_t m,
                       float* dst);
int launch_row_sqnorm(spf_ctx* c, const float* rows, uint32_t ld, uint64_t m, float* out);
int launch_fill_f32(spf_ctx* c, float* p, uint64_t n, float v);
int launch_add_f32(spf_ctx* c, float* dst, const float* src, uint64_t n);          // dst[i] = fl(dst[i] + src[i])
int launch_scale_u64_f32(spf_ctx* c, const uint64_t* src, float scale, uint64_t n, float* dst);   // dst[i] = fl(scale * (float)src[i])
int launch_fill_u64(spf_ctx* c, uint64_t* p, uint64_t n, uint64_t v);
int launch_max_f32(spf_ctx* c, const float* p, uint64_t n, float* out1);   // out1[0] = max
// dist[i] = metric(A row ai, B row bi): ai = idxA ? idxA[i] : i ;
// bi = idxB32 ? idxB32[i] : (fixedB == UINT64_MAX ? i : fixedB)
int launch_pair_dist(spf_ctx* c, int metric, const float* A, uint32_t ldA, const uint64_t* idxA,
                     const float* B, uint32_t ldB, const uint32_t* idxB32, uint64_t fixedB,
                     uint32_t ld, uint64_t count, float* out);
// *d_flag = 0 when a[0..n) and b[0..n) differ in any bit (left untouched otherwise)
int launch_rows_equal(spf_ctx* c, const float* a, const float* b, uint64_t n, int* d_flag);
int launch_check_rows(spf_ctx* c, const uint64_t* d_idx, uint64_t m, uint64_t n, int* d_flag);

// ---- assign_exact.cu ----------------------------------------------------------------------
// CUDA-core direct-form kernel: every distance of the m x k problem, exact.  Emits boundary
// candidates (cand != NULL, one segment, all records exact) and/or the dense m x k matrix.
// d_skip (optional device flag): when it is non-zero at run time the kernel does nothing.
// penalty (optional, k floats, candidate mode only): candidates are formed on fl(d + penalty[j]).
// seed (optional, m floats, candidate mode only): per point the exact distance to SOME centroid (an
// upper bound of its minimum) — the boundary threshold starts tight instead of at +inf.
int launch_assign_exact(spf_ctx* c, int metric, const float* P, uint64_t m, const float* C, uint32_t k,
                        uint32_t ld, float factor, const CandBuf* cand, float* dense, const int* d_skip = nullptr,
                        const float* penalty = nullptr, const float* seed = nullptr);

// ---- assign_tc.cu -------------------------------------------------------------------------
// tcgen05 (TF32) candidate GEMM for squared-Euclidean: approximate distances with a certified
// error bound, candidates only (two segments per point).  Ptf / Ctf are the rounded operands;
// cext_pad holds round_up(k,256) K-extension rows of 8 floats (launch_centroid_ext);
// cstat = {max |c|^2, max |c - c'|}.  Record values are s = x.c - |c|^2/2 (d ~ |x|^2 - 2 s).
// seed (optional, m floats): per point an upper bound of its minimum distance.
bool assign_tc_supported(const spf_ctx* c, uint64_t m, uint32_t k, uint32_t ld);
int launch_assign_tc(spf_ctx* c, const float* Ptf, uint64_t m, const float* Ctf, uint32_t k, uint32_t ld,
                     const float* xnorm, const float* xres, const float* cext_pad, const float* d_cstat,
                     const float* seed, float factor, const CandBuf& cand, uint32_t eld = 0);
// K-extension rows for the tensor kernel: row j < k = {h, m, 0, 0, l, 0, 0, 0} with h + m + l =
// -|c_j|^2 / 2 split into three TF32 values (residual < 2^-33 |c_j|^2); rows k .. kpad-1 = {-inf, 0, ...}.
int launch_centroid_ext(spf_ctx* c, const float* cnorm, uint32_t k, uint32_t kpad, float* cext);

// ---- resolve.cu ---------------------------------------------------------------------------
struct ResolveArgs {
  int metric;
  const float* P; uint64_t m; const float* C; uint32_t k; uint32_t ld;   // exact (unrounded) rows
  float factor;
  CandBuf cand;
  int nseg;                // segments per point: 1 (exact kernel) or 2 (tensor kernel)
  const float* seed;       // tensor path, optional: the seeds the candidate kernel used (validated here)
  const float* penalty = nullptr;   // balanced assignment (extension): k per-centroid penalties added to every exact distance
  uint32_t eld = 0;        // row length entering the certified error bound (0: ld)
  const float* xnorm;      // NULL on the exact path (error bound 0)
  const float* xres;       // tensor path only
  const float* d_cstat;    // device {max |c|^2, max |c - c'|}, tensor path only
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
// The point list of one assign is resolved in chunks (bounded scratch; lets the host-streamed
// path overlap uploads with compute):
//   resolve_begin   allocates the per-call state for m_total points, chunks of <= chunk_rows
//   resolve_chunk   classify / exact_eval / finalize for one chunk; `a` holds chunk-relative
//                   pointers (P, xnorm, xres, best, dmin, nmem, cand) and a.m = rows of the chunk,
//                   r0 = first row of the chunk in the point list
//   resolve_finish  dense fallback for the overflow rows of all chunks, then the CSR (csr == NULL
//                   skips it); `a` describes the whole list (candidate fields unused)
struct ResolveState;
int resolve_begin(spf_ctx* c, uint64_t m_total, uint64_t chunk_rows, bool approx, bool want_members,
                  ResolveState** out);
int resolve_chunk(spf_ctx* c, ResolveState* s, const ResolveArgs& a, uint64_t r0);
int resolve_finish(spf_ctx* c, ResolveState* s, const ResolveArgs& a, CsrOut* csr);
void resolve_free(ResolveState* s);

// ---- search.cu / scan_tc.cu -----------------------------------------------------------------
// Arguments shared by the posting-list scan kernels (see search.cu).
struct ScanArgs {
  const float* vecs; const uint64_t* slot_ids; const uint64_t* grp_off; const uint32_t* lens;
  uint32_t ld; uint32_t d;
  const float* Q; const uint32_t* probe; const float* thr; const uint32_t* seqbase;
  uint32_t nprobe; uint32_t K;
  uint64_t* out_ids; float* out_dists; uint32_t* out_counts; unsigned long long* out_keys;
  unsigned long long* out_slots;   // nq x K slot index of each result (vector gather)
  unsigned long long* bytes;
  const uint8_t* only;             // query-major kernel: when set, only queries with only[q] != 0 run
  // query-major kernel, K > 128 in passes of <= 128 results: this pass writes columns
  // [out_col, out_col + K) of rows out_stride wide and takes only keys above the previous pass's last
  // one (keys are unique per query).  out_stride = 0: a single pass, rows K wide.
  uint32_t out_stride = 0, out_col = 0;
};

// TF32 side structures of an index for the tensor-core candidate scan (scan_tc.cu), made lazily
// from the slot layout: row-major rounded copies of the slot vectors (row = slot), their
// K-extension rows {h, m, 0, 0, l, 0, 0, 0} (h + m + l = -|v|^2/2; {-inf, 0, ...} for pad slots)
// and vstat = {max |v|^2, max |v - v'|}.
struct ScanTcSide {
  float* vtf = nullptr;
  float* vext = nullptr;
  float* vstat = nullptr;
  uint64_t nslots = 0;
  bool ready = false;
};
int scan_tc_prepare(spf_ctx* c, const float* vecs, const uint64_t* slot_ids, uint64_t nslots, uint32_t ld,
                    ScanTcSide* side);
void scan_tc_release(ScanTcSide* side);
bool scan_tc_supported(const spf_ctx* c, uint32_t ld, uint64_t nslots, uint32_t K, uint64_t npairs);
// Tensor-core candidate scan of all (query, probe) pairs: TF32 GEMM of the probing queries of a list
// against the list's vectors, certified candidate filter, exact re-evaluation and top-k per query.
// Fills the out_* arrays of `s` for every query with qflag[q] == 0; the flagged ones (no certified
// bound, or more candidates than the bucket holds) must be re-run by the exact query-major kernel.
struct ScanTcCall {
  ScanArgs s;
  const ScanTcSide* side;
  const uint32_t* pair_sorted;     // (q * nprobe + p) grouped by probed list
  const uint32_t* list_off;        // nlists + 1 offsets into pair_sorted
  uint32_t nlists;
  uint64_t nq;
  uint8_t* qflag;                  // nq, out
  bool is_probe = false;           // the centroid probe (names of the timers only)
  // list-sharded search: the certified bound on every query's K-th distance is min-reduced over the
  // ranks before the candidates are flagged (a rank that holds none of a query's near lists would
  // otherwise refine against its own, much looser bound).  Every rank runs exactly one reduction
  // per scan, whatever path it takes (search.cu: search_scan).
  const spf_comm* comm = nullptr;
  bool* bound_exchanged = nullptr;
};
int scan_tc_run(spf_ctx* c, const ScanTcCall& call);
// Dense tensor-core centroid probe for 32 < nprobe <= 1024, nlists <= 4096: TF32 s of every query
// against every centroid, nprobe-th largest s per query, exact evaluation of everything within the
// certified bound, sort.  Writes probe / thr / seqbase; sets *d_redo when the exact probe must run.
int probe_tc_dense(spf_ctx* c, const ScanTcSide& side, const float* centroids, const float* Q, uint64_t nq, uint32_t ld,
                   uint32_t nlists, uint32_t nprobe, float prune_factor, const uint32_t* lens, uint32_t* probe,
                   float* thr, uint32_t* seqbase, int* d_redo);

// ---- ops.cu (shared with kmeans.cu) ---------------------------------------------------------
int launch_cluster_sums(spf_ctx* c, const float* X, uint32_t ld, const uint64_t* d_offsets, const uint64_t* d_rows,
                        uint32_t k, float* out, int divide);
int launch_medoid_keys(spf_ctx* c, int metric, const float* X, uint32_t ld, const uint64_t* d_rows, uint64_t total,
                       const uint64_t* d_offsets, uint32_t k, const float* means, unsigned long long* keys);

// ---- assign_api.cu ------------------------------------------------------------------------
// spf_assign_vectors with the k centroid vectors already on the device (k x ld, rows zero padded to
// ld); the caller holds the context lock.
int assign_device_centroids(spf_dataset* ds, int metric, const float* d_centroids, uint32_t k, float boundary_factor,
                            int flags, const float* d_seed, const float* d_penalty, spf_assign_result** out);
int dataset_alloc(spf_ctx* c, uint64_t n, uint32_t d, spf_dataset** out);   // api.cu: device buffer only
int dataset_prep_alloc(spf_dataset* ds);
int dataset_prep(spf_dataset* ds);   // rounded copy + norms of all rows, once per dataset
int assign_members_as_rows(const spf_assign_result* r, uint64_t* d_out);   // positions → dataset rows

// pageable host buffers through the context's pinned staging ring (assign_api.cu)
bool host_pointer_is_pageable(const void* p);
int staged_upload(spf_ctx* c, const float* rows, uint64_t n, uint32_t d, uint64_t row_stride, float* dst, uint32_t ld,
                  uint64_t chunk_rows, cudaStream_t copy, const std::function<int(uint64_t)>& after_chunk);
int staged_download(spf_ctx* c, void* host, const void* dev, size_t bytes, cudaStream_t st);

}  // namespace spf
