// assign_exact.cu — CUDA-core direct-form distance kernel (all three metrics).
//
// Replaces the n x k loop of assign_points_to_clusters, src/clustering/hierarchical.rs:302-326
// (reference), where every distance is one DistanceMetric::compute call
// (src/distances/distance.rs:16-43).  Each (point, centroid) accumulator runs over the
// dimensions in index order with un-fused f32 ops, so every distance is bit-identical to the
// reference.  It is the production path for Manhattan / Chebyshev (FP32-pipe bound: 2 lane
// instructions per element-op) and the exact fallback / validator for squared-Euclidean.
//
// Two kernels:
//  * assign_exact_tma_kernel (Chebyshev by default; any metric through the "exact_tma" knob):
//    CTA = 128 points x 128 centroids, 8 compute warps + 1 TMA producer warp, 8 x 8 register tile
//    per thread.  Point and centroid tiles of 16 dimensions (64-byte rows, SWIZZLE_64B) arrive by
//    cp.async.bulk.tensor into a 4-stage mbarrier ring; a thread reads its 8 points / 8 centroids
//    of four dimensions with 16 conflict-free LDS.128 and issues 512 metric instructions on them
//    (3 % load overhead).  The centroid rows are read through a permuted copy (tile row
//    tx + 16 j <-> slot 8 tx + j), which makes the swizzled loads conflict-free and gives every
//    thread 8 consecutive slots = two group records.  The row minimum is folded with warp
//    shuffles over the 16 lanes that share a point and kept in registers across the centroid
//    tiles; record slots are handed out by ballot.  No shared-memory atomics, no __syncthreads.
//  * assign_exact_kernel (Manhattan and squared-L2 by default, and tiny problems): CTA = 64 x 64,
//    4 x 4 register tile, __ldg staging, shared-memory atomics for the row minimum.
#include "tc_ptx.cuh"

namespace spf {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;
constexpr int NTHREADS = 256;

template <int METRIC>
__global__ void __launch_bounds__(NTHREADS)
assign_exact_kernel(const float* __restrict__ P, uint32_t m, const float* __restrict__ C, uint32_t k,
                    uint32_t ld, float factor, CandRec* __restrict__ cand, RowInfo* __restrict__ info,
                    int cap, float* __restrict__ dense, int symmetric, const int* __restrict__ skip,
                    const float* __restrict__ penalty, const float* __restrict__ seed) {
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
    // seed (optional): the distance of the point to SOME centroid, i.e. an upper bound of its minimum —
    // the boundary threshold is tight from the first tile on instead of after the nearest one was met
    rowmin[tid] = (seed != nullptr && row0 + tid < m) ? __float_as_uint(fminf(seed[row0 + tid], __int_as_float(0x7f800000)))
                                                      : 0x7f800000u;   // +inf
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

// ---------------------------------------------------------------------------------------------
// TMA-staged, register-tiled kernel
// ---------------------------------------------------------------------------------------------
constexpr int XT_M = 128, XT_N = 128, XT_K = 16;      // points x centroids x dimensions per stage
constexpr int xt_stages(int minb) { return minb == 1 ? 8 : 4; }   // one CTA per SM: room for a deeper ring
constexpr int XT_TILE_BYTES = XT_M * XT_K * 4;        // 8 KB per operand and stage
constexpr int XT_THREADS = 256;                       // 8 compute warps; thread 0 also issues the TMA loads
// MINB (resident CTAs per SM the kernel is compiled for) also selects the shape: 1 or 2 -> 128 points x 128
// centroids, 256 threads; 3 -> 64 points x 128 centroids, 128 threads (12 warps per SM at <= 168 registers)
constexpr int xt_rows(int minb) { return minb == 3 ? 64 : 128; }
constexpr int xt_smem(int minb) { return xt_stages(minb) * (xt_rows(minb) * XT_K * 4 + XT_TILE_BYTES) + 256 + 1024; }

// slot (within a 128-centroid tile) held by row r of the permuted centroid copy
__host__ __device__ inline uint32_t xt_slot_of_row(uint32_t r) { return 8u * (r & 15u) + (r >> 4); }

__global__ void xt_permute_rows_kernel(const float4* __restrict__ C, uint32_t ld4, uint32_t k, uint32_t kpad,
                                       float4* __restrict__ out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint64_t)kpad * ld4) return;
  const uint32_t row = (uint32_t)(t / ld4), col = (uint32_t)(t - (uint64_t)row * ld4);
  const uint32_t slot = (row & ~127u) + xt_slot_of_row(row & 127u);
  out[t] = slot < k ? C[(size_t)slot * ld4 + col] : make_float4(0.f, 0.f, 0.f, 0.f);
}

// PACKED (Manhattan / Chebyshev): the differences of two consecutive dimensions come from one
// sub.rn.f32x2 (FADD2: both lanes IEEE round-to-nearest like the scalar FADD), the accumulator chain
// stays scalar and in dimension order: 6 issue slots per 4 element-ops instead of 8.
template <int METRIC, bool PACKED, int MINB>
__global__ void __launch_bounds__(2 * xt_rows(MINB), MINB)
assign_exact_tma_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_c, uint32_t m,
                        uint32_t k, uint32_t ld, float factor, CandRec* __restrict__ cand, RowInfo* __restrict__ info,
                        int cap, float* __restrict__ dense, int symmetric, const int* __restrict__ skip,
                        const float* __restrict__ penalty, const float* __restrict__ seed) {
  if (skip != nullptr && *skip != 0) return;
  // (declared with its alignment instead of aligned by pointer arithmetic: the compiler must keep
  // seeing a shared-memory address, or the operand loads become generic LD instead of LDS)
  extern __shared__ __align__(1024) unsigned char xt_raw[];
  constexpr int XT_STAGES = xt_stages(MINB);
  constexpr int M = xt_rows(MINB);                 // points per CTA; a thread owns rows ty + (M / 8) i
  constexpr int XB = M * XT_K * 4;                 // bytes of a point tile
  constexpr uint32_t RSTEP = M / 8;
  unsigned char* smem = xt_raw;
  unsigned char* xs = smem;                                        // [XT_STAGES][128 rows][64 B]
  unsigned char* cs = smem + XT_STAGES * XB;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + XT_STAGES * (XB + XT_TILE_BYTES));
  uint64_t* empty = full + XT_STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t row0 = blockIdx.x * M;
  const uint32_t ntile = (k + XT_N - 1) / XT_N;
  const uint32_t nkb = (ld + XT_K - 1) / XT_K;
  if (tid == 0) {
    for (int i = 0; i < XT_STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 2 * M / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // The centroid tiles this CTA visits: t = blockIdx.y, + gridDim.y, ...; symmetric mode skips the
  // tiles strictly below the diagonal (mirrored from above).  Thread 0 is the TMA producer: it keeps
  // XT_STAGES - 1 stages in flight ahead of the compute loop (cursor pt / pkb), refilling a stage as
  // soon as all 8 warps have released it.
  auto next_tile = [&](uint32_t t) {                    // first visited tile >= t
    while (t < ntile && symmetric && t < blockIdx.x) t += gridDim.y;
    return t;
  };
  uint32_t pt = next_tile(blockIdx.y), pkb = 0, ps = 0, pph = 0;
  auto produce = [&]() {                                // thread 0 only: one stage, if any is left
    if (pt >= ntile) return;
    tc::mbar_wait(&empty[ps], pph ^ 1);
    tc::mbar_expect_tx(&full[ps], XB + XT_TILE_BYTES);
    tc::tma_load_2d(xs + ps * XB, &map_x, &full[ps], (int)(pkb * XT_K), (int)row0);
    tc::tma_load_2d(cs + ps * XT_TILE_BYTES, &map_c, &full[ps], (int)(pkb * XT_K), (int)(pt * XT_N));
    if (++ps == XT_STAGES) { ps = 0; pph ^= 1; }
    if (++pkb == nkb) { pkb = 0; pt = next_tile(pt + gridDim.y); }
  };
  if (tid == 0)
    for (int i = 0; i < XT_STAGES - 1; ++i) produce();
  // ------------------------------ compute warps ---------------------------------------------------
  const uint32_t tx = tid & 15, ty = tid >> 4;                     // centroid / point direction
  const uint32_t half_mask = (lane & 16) ? 0xffff0000u : 0x0000ffffu;   // the 16 lanes that share this thread's points
  // SWIZZLE_64B: the 16-byte chunk c of row r sits at chunk c ^ ((r >> 1) & 3); rows tx + 16 j (ty + 16 i)
  // all share (tx >> 1) & 3 ((ty >> 1) & 3)
  const uint32_t cbase = tx * 64u, cswz = (tx >> 1) & 3u;
  const uint32_t xbase = ty * 64u, xswz = (ty >> 1) & 3u;
  const float INF = __int_as_float(0x7f800000);
  float runmin[8];
  uint32_t cnt[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t r = row0 + ty + RSTEP * i;
    runmin[i] = (seed != nullptr && r < m) ? fminf(seed[r], INF) : INF;   // upper bound of the minimum (NaN -> +inf)
    cnt[i] = 0;
  }
  uint32_t s = 0, ph = 0;
  for (uint32_t t = blockIdx.y; t < ntile; t += gridDim.y) {
    if (symmetric && t < blockIdx.x) continue;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    for (uint32_t kb = 0; kb < nkb; ++kb) {
      if (tid == 0) produce();
      tc::mbar_wait(&full[s], ph);
      const unsigned char* xb = xs + s * XB;
      const unsigned char* cb = cs + s * XT_TILE_BYTES;
      // the four 16-byte chunks of the stage are a real loop (not unrolled): 512 metric steps per
      // iteration keep the body inside the instruction cache (fully unrolled, 17 % of the stall
      // samples were instruction fetches)
#pragma unroll 1
      for (uint32_t c = 0; c < 4; ++c) {
        const unsigned char* cp = cb + cbase + ((c ^ cswz) << 4);
        const unsigned char* xp = xb + xbase + ((c ^ xswz) << 4);
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {                 // 4 centroids at a time: 16 operand registers
          if constexpr (PACKED && METRIC != SPF_METRIC_EUCLIDEAN) {
            ulonglong2 cv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) cv[j] = *reinterpret_cast<const ulonglong2*>(cp + (jh * 4 + j) * 1024);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const ulonglong2 xv = *reinterpret_cast<const ulonglong2*>(xp + i * (RSTEP * 64));
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                unsigned long long d01, d23;
                asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d01) : "l"(xv.x), "l"(cv[j].x));
                asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d23) : "l"(xv.y), "l"(cv[j].y));
                float d0, d1, d2, d3;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d01));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(d2), "=f"(d3) : "l"(d23));
                float a = acc[i][jh * 4 + j];
                if constexpr (METRIC == SPF_METRIC_MANHATTAN) {
                  a = __fadd_rn(a, fabsf(d0)); a = __fadd_rn(a, fabsf(d1));
                  a = __fadd_rn(a, fabsf(d2)); a = __fadd_rn(a, fabsf(d3));
                } else {
                  a = fmaxf(a, fabsf(d0)); a = fmaxf(a, fabsf(d1));
                  a = fmaxf(a, fabsf(d2)); a = fmaxf(a, fabsf(d3));
                }
                acc[i][jh * 4 + j] = a;
              }
            }
          } else {
          float4 cv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) cv[j] = *reinterpret_cast<const float4*>(cp + (jh * 4 + j) * 1024);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 xv = *reinterpret_cast<const float4*>(xp + i * (RSTEP * 64));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float a = acc[i][jh * 4 + j];
              a = dist_step<METRIC>(a, xv.x, cv[j].x);
              a = dist_step<METRIC>(a, xv.y, cv[j].y);
              a = dist_step<METRIC>(a, xv.z, cv[j].z);
              a = dist_step<METRIC>(a, xv.w, cv[j].w);
              acc[i][jh * 4 + j] = a;
            }
          }
          }
        }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&empty[s]);
      if (++s == XT_STAGES) { s = 0; ph ^= 1; }
    }
    // ---- epilogue of this 128 x 128 tile: thread (tx, ty) holds points row0 + ty + 16 i and the
    // consecutive centroid slots c0 .. c0 + 7 ------------------------------------------------------
    const uint32_t c0 = t * XT_N + 8u * tx;
    if (dense != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t r = row0 + ty + RSTEP * i;
        if (r < m) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t cj = c0 + j;
            if (cj < k) {
              dense[(size_t)r * k + cj] = acc[i][j];
              if (symmetric && t > blockIdx.x) dense[(size_t)cj * k + r] = acc[i][j];   // mirror (off-diagonal tiles)
            }
          }
        }
      }
    }
    if (cand != nullptr) {
      if (penalty != nullptr) {             // balanced assignment: candidates are formed on the costs
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float pj = (c0 + j < k) ? penalty[c0 + j] : 0.0f;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i][j] = __fadd_rn(acc[i][j], pj);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float tmin = INF;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < k) tmin = fminf(tmin, acc[i][j]);             // fminf skips NaN
        // minimum over the 16 lanes that hold the other centroids of this point
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
        runmin[i] = fminf(runmin[i], tmin);
        const float rm = runmin[i];
        const float thr = __fmul_rn(rm, factor);
        const uint32_t r = row0 + ty + RSTEP * i;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          bool hit = false;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float dv = acc[i][g * 4 + q];
            hit = hit || (c0 + g * 4 + q < k && (dv < thr || dv == rm));
          }
          hit = hit && r < m;
          // record slots by ballot over the lanes of this point (one group record for the thread's 4
          // consecutive centroids; slots >= k are ignored by index in resolve)
          const unsigned b = __ballot_sync(0xffffffffu, hit) & half_mask;
          const uint32_t pos = cnt[i] + (uint32_t)__popc(b & ((1u << lane) - 1u));
          cnt[i] += (uint32_t)__popc(b);
          if (hit && pos < (uint32_t)cap) {
            CandRec* w = cand + (size_t)r * cap + pos;
            w->t = make_float4(acc[i][g * 4 + 0], acc[i][g * 4 + 1], acc[i][g * 4 + 2], acc[i][g * 4 + 3]);
            w->g = ((c0 + g * 4) >> 2) | REC_ALL_EXACT;
          }
        }
      }
    }
  }
  if (cand != nullptr && tx == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t r = row0 + ty + RSTEP * i;
      if (r < m) info[r] = make_uint4(cnt[i], __float_as_uint(runmin[i]), 0u, 0x7f800000u);
    }
  }
}

// 2-D map over a row-major `rows x ld` f32 matrix: boxes of 128 rows x 16 floats, SWIZZLE_64B
int xt_make_map(spf_ctx* c, CUtensorMap* map, const float* base, uint64_t rows, uint32_t ld, uint32_t box_rows) {
  cuuint64_t gdim[2] = {ld, rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {XT_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<tc::EncodeTiledFn>(c->tma_encode)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SPF_E_CUDA, "cuTensorMapEncodeTiled (exact kernel) failed with code %d", (int)r);
  return SPF_OK;
}

template <int METRIC, bool PACKED, int MINB>
int launch_xt_shape(spf_ctx* c, const float* P, uint64_t m, const float* Cperm, uint32_t k, uint32_t ld, float factor,
                    CandRec* cand, RowInfo* info, int cap, float* dense, int symmetric, const int* d_skip, const float* penalty,
                    const float* seed) {
  constexpr int M = xt_rows(MINB);
  CUtensorMap map_x, map_c;
  SPF_TRY(xt_make_map(c, &map_x, P, m, ld, M));
  SPF_TRY(xt_make_map(c, &map_c, Cperm, (uint64_t)round_up(k, XT_N), ld, XT_N));
  dim3 grid((unsigned)ceil_div(m, M));
  if (cand == nullptr) {   // dense matrix only: fill the machine even when m is small
    const uint64_t ctiles = ceil_div(k, XT_N);
    uint64_t want = ceil_div((uint64_t)c->sm_count * 2, grid.x);
    grid.y = (unsigned)(want < 1 ? 1 : (want > ctiles ? ctiles : want));
  }
  SPF_CUDA(cudaFuncSetAttribute(assign_exact_tma_kernel<METRIC, PACKED, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                xt_smem(MINB)));
  assign_exact_tma_kernel<METRIC, PACKED, MINB><<<grid, 2 * M, xt_smem(MINB), c->stream>>>(
      map_x, map_c, (uint32_t)m, k, ld, factor, cand, info, cap, dense, symmetric, d_skip, penalty, seed);
  return check_launch(c, "assign_exact_tma_kernel");
}

template <int METRIC, bool PACKED>
int launch_xt(spf_ctx* c, const float* P, uint64_t m, const float* Cperm, uint32_t k, uint32_t ld, float factor,
              CandRec* cand, RowInfo* info, int cap, float* dense, int symmetric, const int* d_skip, const float* penalty,
              const float* seed) {
  // the 64-row shape has no symmetric (tile index == row block index) mode: candidate launches only
  if (PACKED && cand != nullptr && ((c->params.exact_three_cta >> METRIC) & 1) != 0)
    return launch_xt_shape<METRIC, PACKED, 3>(c, P, m, Cperm, k, ld, factor, cand, info, cap, dense, symmetric, d_skip, penalty, seed);
  if (PACKED && ((c->params.exact_one_cta >> METRIC) & 1) != 0)
    return launch_xt_shape<METRIC, PACKED, 1>(c, P, m, Cperm, k, ld, factor, cand, info, cap, dense, symmetric, d_skip, penalty, seed);
  return launch_xt_shape<METRIC, PACKED, 2>(c, P, m, Cperm, k, ld, factor, cand, info, cap, dense, symmetric, d_skip, penalty, seed);
}

}  // namespace

int launch_assign_exact(spf_ctx* c, int metric, const float* P, uint64_t m, const float* C, uint32_t k,
                        uint32_t ld, float factor, const CandBuf* cb, float* dense, const int* d_skip, const float* penalty,
                        const float* seed) {
  // a dense P x P request is the symmetric centroid matrix: half the tiles
  const int symmetric = (cb == nullptr && dense != nullptr && P == C && m == k) ? 1 : 0;
  if (m == 0 || k == 0) return SPF_OK;
  CandRec* cand = cb ? cb->rec : nullptr;
  RowInfo* info = cb ? cb->info : nullptr;
  const int cap = cb ? cb->cap : 0;
  if (metric < 0 || metric > 2) return fail(SPF_E_INVALID, "unknown metric %d", metric);
  // per-metric choice (bit `metric` of the knobs), from measurements at d = 128 and d = 960 (fractions of
  // 2 lane instructions per element-op at one instruction per clock and SM sub-partition):
  //   Chebyshev  4 x 4 kernel 0.58-0.61 | 8 x 8 TMA kernel 0.74-0.84 | + packed FADD2 differences 0.89-0.92
  //   Manhattan  4 x 4 kernel 0.69-0.73 | 8 x 8 TMA 0.60-0.68 (dispatch stalls: 0.77 per issued instruction,
  //              128 registers) | packed 0.73 | packed, compiled for one CTA per SM (236 registers, no
  //              spills) 0.80-0.81
  //   squared-L2 stays on the 4 x 4 kernel (0.76 against 0.68); its packed form is not used because ptxas
  //              contracts mul.f32x2 + add.f32x2 into FFMA2, which changes the rounding
  const bool use_tma = c->tma_encode != nullptr && ((c->params.exact_tma >> metric) & 1) != 0 &&
                       (uint64_t)m * k >= (uint64_t)c->params.exact_tma_min_pairs && (reinterpret_cast<uintptr_t>(P) & 15) == 0;
  if (use_tma) {
    // permuted copy of the centroid rows (tile row tx + 16 j <-> slot 8 tx + j), zero rows beyond k
    const uint32_t kpad = round_up(k, XT_N), ld4 = ld / 4;
    DevBuf<float> cperm;
    SPF_TRY(cperm.alloc(c->stream, (size_t)kpad * ld));
    xt_permute_rows_kernel<<<(unsigned)ceil_div((uint64_t)kpad * ld4, 256), 256, 0, c->stream>>>(
        reinterpret_cast<const float4*>(C), ld4, k, kpad, reinterpret_cast<float4*>(cperm.p));
    SPF_TRY(check_launch(c, "xt_permute_rows_kernel"));
    const bool packed = ((c->params.exact_packed >> metric) & 1) != 0;
    const float* pen = cb ? penalty : nullptr;
    const float* sd = cb ? seed : nullptr;
    switch (metric) {
      case SPF_METRIC_EUCLIDEAN:
        return launch_xt<SPF_METRIC_EUCLIDEAN, false>(c, P, m, cperm.p, k, ld, factor, cand, info, cap, dense, symmetric, d_skip, pen, sd);
      case SPF_METRIC_MANHATTAN:
        if (packed)
          return launch_xt<SPF_METRIC_MANHATTAN, true>(c, P, m, cperm.p, k, ld, factor, cand, info, cap, dense, symmetric, d_skip, pen, sd);
        return launch_xt<SPF_METRIC_MANHATTAN, false>(c, P, m, cperm.p, k, ld, factor, cand, info, cap, dense, symmetric, d_skip, pen, sd);
      default:
        if (packed)
          return launch_xt<SPF_METRIC_CHEBYSHEV, true>(c, P, m, cperm.p, k, ld, factor, cand, info, cap, dense, symmetric, d_skip, pen, sd);
        return launch_xt<SPF_METRIC_CHEBYSHEV, false>(c, P, m, cperm.p, k, ld, factor, cand, info, cap, dense, symmetric, d_skip, pen, sd);
    }
  }
  dim3 grid((unsigned)ceil_div(m, BM)), block(NTHREADS);
  if (cand == nullptr) {   // dense matrix only: fill the machine even when m is small
    const uint64_t ctiles = ceil_div(k, BN);
    uint64_t want = ceil_div((uint64_t)c->sm_count * 4, grid.x);
    grid.y = (unsigned)(want < 1 ? 1 : (want > ctiles ? ctiles : want));
  }
  switch (metric) {
    case SPF_METRIC_EUCLIDEAN:
      assign_exact_kernel<SPF_METRIC_EUCLIDEAN><<<grid, block, 0, c->stream>>>(
          P, (uint32_t)m, C, k, ld, factor, cand, info, cap, dense, symmetric, d_skip, cb ? penalty : nullptr, cb ? seed : nullptr);
      break;
    case SPF_METRIC_MANHATTAN:
      assign_exact_kernel<SPF_METRIC_MANHATTAN><<<grid, block, 0, c->stream>>>(
          P, (uint32_t)m, C, k, ld, factor, cand, info, cap, dense, symmetric, d_skip, cb ? penalty : nullptr, cb ? seed : nullptr);
      break;
    default:
      assign_exact_kernel<SPF_METRIC_CHEBYSHEV><<<grid, block, 0, c->stream>>>(
          P, (uint32_t)m, C, k, ld, factor, cand, info, cap, dense, symmetric, d_skip, cb ? penalty : nullptr, cb ? seed : nullptr);
      break;
  }
  return check_launch(c, "assign_exact_kernel");
}

}  // namespace spf
