// assign_exact.cu — CUDA-core direct-form distance kernel (all three metrics).
//
// Replaces the n x k loop of assign_points_to_clusters, src/clustering/hierarchical.rs:302-326
// (reference), where every distance is one DistanceMetric::compute call
// (src/distances/distance.rs:16-43).  Each (point, centroid) accumulator runs over the
// dimensions in index order with un-fused f32 ops, so every distance is bit-identical to the
// reference.  It is the production path for Manhattan / Chebyshev (FP32-pipe bound: 2 lane
// instructions per element-op) and the exact fallback / validator for squared-Euclidean.
//
// Tiling: CTA = 64 points x 64 centroids, 256 threads, 4x4 register micro-tile per thread,
// 16-dimension stages double-buffered through shared memory (stored dimension-major so a
// thread fetches its 4 points / 4 centroids of one dimension with two LDS.128).  A CTA keeps
// its 64 points and walks all centroid tiles, so the running row minimum and the candidate
// counters live in shared memory and no global atomics are needed.
#include "kernels.cuh"

namespace spf {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;
constexpr int NTHREADS = 256;

template <int METRIC>
__global__ void __launch_bounds__(NTHREADS)
assign_exact_kernel(const float* __restrict__ P, uint32_t m, const float* __restrict__ C, uint32_t k,
                    uint32_t ld, float factor, CandRec* __restrict__ cand, RowInfo* __restrict__ info,
                    int cap, float* __restrict__ dense, int symmetric, const int* __restrict__ skip,
                    const float* __restrict__ penalty) {
  if (skip != nullptr && *skip != 0) return;   // the caller already holds this result (cached centroid matrix)
  __shared__ __align__(16) float Xs[2][BK][BM + PAD];
  __shared__ __align__(16) float Cs[2][BK][BN + PAD];
  __shared__ unsigned rowmin[BM];
  __shared__ unsigned rowcnt[BM];

  const int tid = threadIdx.x;
  const int tx = tid & 15;   // centroid direction
  const int ty = tid >> 4;   // point direction
  const uint32_t row0 = blockIdx.x * BM;

  if (tid < BM) {
    rowmin[tid] = 0x7f800000u;   // +inf
    rowcnt[tid] = 0;
  }
  __syncthreads();

  // global -> smem staging map: one float4 (4 dims) of one row per thread
  const int lrow = tid >> 2;        // 0..63
  const int lkq = (tid & 3) * 4;    // 0,4,8,12
  const uint32_t nkb = (ld + BK - 1) / BK;
  const bool prow_ok = (row0 + lrow) < m;
  const float* prow = P + (size_t)(row0 + lrow) * ld;

  // dense-only launches may split the centroid tiles over blockIdx.y (candidate mode keeps a
  // CTA-wide running minimum per row and therefore uses gridDim.y == 1)
  for (uint32_t c0 = blockIdx.y * BN; c0 < k; c0 += gridDim.y * BN) {
    // symmetric mode (P == C, e.g. the k x k centroid matrix): d(a,b) == d(b,a) bit for bit for all
    // three metrics, so tiles strictly below the diagonal are skipped and mirrored from above
    if (symmetric && c0 + BN <= row0) continue;
    const bool crow_ok = (c0 + lrow) < k;
    const float* crow = C + (size_t)(c0 + lrow) * ld;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    float4 xr, cr;
    auto load_global = [&](uint32_t kb) {
      uint32_t col = kb * BK + lkq;
      xr = (prow_ok && col < ld) ? __ldg(reinterpret_cast<const float4*>(prow + col))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
      cr = (crow_ok && col < ld) ? __ldg(reinterpret_cast<const float4*>(crow + col))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto store_smem = [&](int buf) {
      Xs[buf][lkq + 0][lrow] = xr.x; Xs[buf][lkq + 1][lrow] = xr.y;
      Xs[buf][lkq + 2][lrow] = xr.z; Xs[buf][lkq + 3][lrow] = xr.w;
      Cs[buf][lkq + 0][lrow] = cr.x; Cs[buf][lkq + 1][lrow] = cr.y;
      Cs[buf][lkq + 2][lrow] = cr.z; Cs[buf][lkq + 3][lrow] = cr.w;
    };

    int buf = 0;
    load_global(0);
    store_smem(0);
    __syncthreads();
    for (uint32_t kb = 0; kb < nkb; ++kb) {
      if (kb + 1 < nkb) load_global(kb + 1);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 xa = *reinterpret_cast<const float4*>(&Xs[buf][kk][ty * 4]);
        const float4 cb = *reinterpret_cast<const float4*>(&Cs[buf][kk][tx * 4]);
        const float xv[4] = {xa.x, xa.y, xa.z, xa.w};
        const float cv[4] = {cb.x, cb.y, cb.z, cb.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = dist_step<METRIC>(acc[i][j], xv[i], cv[j]);
      }
      if (kb + 1 < nkb) store_smem(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }

    // ---- epilogue for this 64 x 64 tile -----------------------------------------------------
    if (dense != nullptr) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t r = row0 + ty * 4 + i;
        if (r < m) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t cj = c0 + tx * 4 + j;
            if (cj < k) {
              dense[(size_t)r * k + cj] = acc[i][j];
              if (symmetric && cj >= row0 + BM) dense[(size_t)cj * k + r] = acc[i][j];   // mirror (off-diagonal tiles)
            }
          }
        }
      }
    }
    if (cand != nullptr) {
      if (penalty != nullptr) {             // balanced assignment: candidates are formed on the costs
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float pj = (c0 + tx * 4 + j < k) ? penalty[c0 + tx * 4 + j] : 0.0f;
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i][j] = __fadd_rn(acc[i][j], pj);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float tmin = __int_as_float(0x7f800000);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c0 + tx * 4 + j < k) tmin = fminf(tmin, acc[i][j]);   // fminf skips NaN
        atomicMin(&rowmin[ty * 4 + i], __float_as_uint(tmin));       // distances are >= +0
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t r = row0 + ty * 4 + i;
        if (r >= m) continue;
        const float rm = __uint_as_float(rowmin[ty * 4 + i]);
        const float thr = __fmul_rn(rm, factor);
        bool hit = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float dv = acc[i][j];
          hit = hit || (c0 + tx * 4 + j < k && (dv < thr || dv == rm));
        }
        if (hit) {   // one group record for the thread's 4 consecutive centroids (slots >= k are
                     // ignored by index in resolve)
          const unsigned slot = atomicAdd(&rowcnt[ty * 4 + i], 1u);
          if (slot < (unsigned)cap) {
            CandRec* w = cand + (size_t)r * cap + slot;
            w->t = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            w->g = ((c0 >> 2) + tx) | REC_ALL_EXACT;
          }
        }
      }
    }
  }
  if (cand != nullptr) {
    __syncthreads();
    if (tid < BM && row0 + tid < m) info[row0 + tid] = make_uint4(rowcnt[tid], rowmin[tid], 0u, 0x7f800000u);
  }
}

}  // namespace

int launch_assign_exact(spf_ctx* c, int metric, const float* P, uint64_t m, const float* C, uint32_t k,
                        uint32_t ld, float factor, const CandBuf* cb, float* dense, const int* d_skip, const float* penalty) {
  // a dense P x P request is the symmetric centroid matrix: half the tiles
  const int symmetric = (cb == nullptr && dense != nullptr && P == C && m == k) ? 1 : 0;
  if (m == 0 || k == 0) return SPF_OK;
  CandRec* cand = cb ? cb->rec : nullptr;
  RowInfo* info = cb ? cb->info : nullptr;
  const int cap = cb ? cb->cap : 0;
  dim3 grid((unsigned)ceil_div(m, BM)), block(NTHREADS);
  if (cand == nullptr) {   // dense matrix only: fill the machine even when m is small
    const uint64_t ctiles = ceil_div(k, BN);
    uint64_t want = ceil_div((uint64_t)c->sm_count * 4, grid.x);
    grid.y = (unsigned)(want < 1 ? 1 : (want > ctiles ? ctiles : want));
  }
  switch (metric) {
    case SPF_METRIC_EUCLIDEAN:
      assign_exact_kernel<SPF_METRIC_EUCLIDEAN><<<grid, block, 0, c->stream>>>(
          P, (uint32_t)m, C, k, ld, factor, cand, info, cap, dense, symmetric, d_skip, cb ? penalty : nullptr);
      break;
    case SPF_METRIC_MANHATTAN:
      assign_exact_kernel<SPF_METRIC_MANHATTAN><<<grid, block, 0, c->stream>>>(
          P, (uint32_t)m, C, k, ld, factor, cand, info, cap, dense, symmetric, d_skip, cb ? penalty : nullptr);
      break;
    case SPF_METRIC_CHEBYSHEV:
      assign_exact_kernel<SPF_METRIC_CHEBYSHEV><<<grid, block, 0, c->stream>>>(
          P, (uint32_t)m, C, k, ld, factor, cand, info, cap, dense, symmetric, d_skip, cb ? penalty : nullptr);
      break;
    default:
      return fail(SPF_E_INVALID, "unknown metric %d", metric);
  }
  return check_launch(c, "assign_exact_kernel");
}

}  // namespace spf
