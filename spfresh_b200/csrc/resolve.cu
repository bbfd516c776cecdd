// resolve.cu — turns per-point candidate lists into the reference's assignment result.
//
// Implements, per listed point, src/clustering/hierarchical.rs:317-346 (reference):
//   (best, dmin) = fold from (0, +inf) with strict `<`            → lowest slot wins ties
//   thr = dmin * BOUNDARY_THRESHOLD (in f32)
//   member j != best  iff  d_j < thr  and  d(c_best, c_j) >= d_j
// then the serial merge :353-361 into per-cluster lists in input order (cluster-major CSR).
//
// Candidates come from the exact kernel (distances already bit-exact) or from the tcgen05 TF32
// GEMM (approximate, with a certified error bound E per point).  Every comparison above is
// decided on exact values: an approximate candidate is only accepted or rejected without
// recomputation when its whole interval [d-E, d+E] lies on one side of the test; otherwise the
// lane recomputes the direct-form distance (same rounding sequence as the reference).
#include <cub/cub.cuh>

#include <algorithm>

#include "kernels.cuh"
#include "pairdist.cuh"

namespace spf {

namespace {

struct ResolveDev {
  const float* P; uint32_t m; const float* C; uint32_t k; uint32_t ld;
  float factor;
  uint2* cand; const uint32_t* cand_cnt; int cap; int nseg;
  const float* xnorm; const float* cnmax; const float* cc;
  uint32_t* best; float* dmin; uint32_t* nmem;
  uint32_t* ovf_rows; uint32_t* ovf_count;
  int want_members;
};

__device__ __forceinline__ bool lex_less(float d1, uint32_t j1, float d2, uint32_t j2) {
  return d1 < d2 || (d1 == d2 && j1 < j2);
}

constexpr int RS_WARPS = 8;           // warps per CTA of the resolve kernel
constexpr int RS_CHUNK = 128;         // dimensions staged per step (one float4 per lane)
constexpr int RS_PAIRS = 4;           // exact distances evaluated per cooperative round
constexpr int RS_TB_STRIDE = RS_CHUNK + 4;

struct ResolveSmem {
  float xs[RS_CHUNK];
  float tb[RS_PAIRS][RS_TB_STRIDE];
};

// Exact distances d(x, C[j]) for the lanes with `want` set (their centroid slot in `j`), four at
// a time: the warp stages 128 dimensions of x and of the four centroid rows in shared memory
// with coalesced float4 loads, then lanes 0..3 each walk one pair in dimension order, so every
// value is the reference's sequential f32 sum (src/distances/distance.rs:16-43).  Returns the
// distance to the owning lane (unchanged `dv` for lanes without `want`).
template <int METRIC>
__device__ __forceinline__ float coop_exact(bool want, uint32_t j, float dv, const float* __restrict__ x,
                                            const float* __restrict__ C, uint32_t ld, ResolveSmem& sm, int lane,
                                            const float4* x0 = nullptr) {
  unsigned pending = __ballot_sync(0xffffffffu, want);
  while (pending) {
    int owner[RS_PAIRS];
    uint32_t jj[RS_PAIRS];
    int np = 0;
#pragma unroll
    for (int p = 0; p < RS_PAIRS; ++p) {
      owner[p] = -1;
      jj[p] = 0;
      if (pending) {
        owner[p] = __ffs(pending) - 1;
        pending &= pending - 1;
        np = p + 1;
      }
      jj[p] = __shfl_sync(0xffffffffu, j, owner[p] < 0 ? 0 : owner[p]);
    }
    float acc = 0.0f;
    for (uint32_t c0 = 0; c0 < ld; c0 += RS_CHUNK) {
      const uint32_t col = c0 + lane * 4;
      const bool ok = col < ld;                      // ld is a multiple of 4
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 xv = (x0 != nullptr && c0 == 0) ? *x0 : (ok ? __ldg(reinterpret_cast<const float4*>(x + col)) : z);
      float4 cv[RS_PAIRS];
#pragma unroll
      for (int p = 0; p < RS_PAIRS; ++p)
        cv[p] = (ok && p < np) ? __ldg(reinterpret_cast<const float4*>(C + (size_t)jj[p] * ld + col)) : z;
      __syncwarp();
      *reinterpret_cast<float4*>(&sm.xs[lane * 4]) = xv;
#pragma unroll
      for (int p = 0; p < RS_PAIRS; ++p) *reinterpret_cast<float4*>(&sm.tb[p][lane * 4]) = cv[p];
      __syncwarp();
      if (lane < np) {
        const int nn = (ld - c0) < (uint32_t)RS_CHUNK ? (int)(ld - c0) : RS_CHUNK;
#pragma unroll 8
        for (int i = 0; i < nn; ++i) acc = dist_step<METRIC>(acc, sm.xs[i], sm.tb[lane][i]);
      }
    }
#pragma unroll
    for (int p = 0; p < RS_PAIRS; ++p) {
      const float r = __shfl_sync(0xffffffffu, acc, p);
      if (lane == owner[p]) dv = r;
    }
  }
  return dv;
}

// One warp per listed point; the whole of hierarchical.rs:317-346 for that point in one pass
// over its candidate slots: approximate minimum → exact distances for everything within 2E of
// it → exact (dmin, best) → boundary tests, recomputing exactly only where the approximate
// value cannot certify the outcome → members compacted to the front of the row's buffer.
template <int METRIC>
__global__ void __launch_bounds__(RS_WARPS * 32, 3) resolve_kernel(ResolveDev a) {
  __shared__ ResolveSmem smem[RS_WARPS];
  const int lane = threadIdx.x & 31;
  ResolveSmem& sm = smem[threadIdx.x >> 5];
  const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
  const float INF = __int_as_float(0x7f800000);
  const uint32_t segcap = (uint32_t)a.cap / (uint32_t)a.nseg;
  // Everything a row needs first (its counts, the first 32 slots of each segment, 128 dimensions
  // of the point) is fetched one row ahead, so the dependent HBM round trips of row r+1 overlap
  // the work on row r.  Slots past the live count are read speculatively and ignored.
  struct Pre { uint32_t cnt0, cnt1; uint2 e0, e1; float4 xv; };
  auto prefetch = [&](uint32_t r) {
    Pre p;
    p.cnt0 = p.cnt1 = 0;
    p.e0 = p.e1 = make_uint2(0u, 0u);
    p.xv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < a.m) {
      p.cnt0 = a.cand_cnt[(size_t)r * a.nseg];
      if (a.nseg > 1) p.cnt1 = a.cand_cnt[(size_t)r * a.nseg + 1];
      const uint2* cr = a.cand + (size_t)r * a.cap;
      if ((uint32_t)lane < segcap) {
        p.e0 = cr[lane];
        if (a.nseg > 1) p.e1 = cr[segcap + lane];
      }
      if ((uint32_t)lane * 4 < a.ld) p.xv = __ldg(reinterpret_cast<const float4*>(a.P + (size_t)r * a.ld) + lane);
    }
    return p;
  };
  uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  Pre nx = prefetch(r);
  for (; r < a.m; r += warps_total) {
    const Pre cur = nx;
    nx = prefetch(r + warps_total);
    // the row's buffer holds nseg segments of segcap slots; cnt_s > segcap marks an overflow
    const uint32_t cnt0 = cur.cnt0, cnt1 = cur.cnt1;
    if (cnt0 > segcap || cnt1 > segcap) {   // overflowed: the brute-force kernels own this row
      if (lane == 0) {
        const uint32_t pos = atomicAdd(a.ovf_count, 1u);
        a.ovf_rows[pos] = r;
        a.nmem[r] = NMEM_OVERFLOW_BIT;
      }
      continue;
    }
    const uint32_t total = cnt0 + cnt1;
    auto slot_of = [&](uint32_t u) { return u < cnt0 ? u : segcap + (u - cnt0); };
    uint2* cr = a.cand + (size_t)r * a.cap;
    const float* x = a.P + (size_t)r * a.ld;
    const float E = a.xnorm ? tc_err_bound(a.xnorm[r], a.cnmax[0], a.ld) : 0.0f;

    // 1. approximate minimum over all candidates
    float ma = INF;
    if ((uint32_t)lane < cnt0) ma = __uint_as_float(cur.e0.y);
    if ((uint32_t)lane < cnt1) ma = fminf(ma, __uint_as_float(cur.e1.y));
    for (uint32_t s2 = 32 + lane; s2 < cnt0; s2 += 32) ma = fminf(ma, __uint_as_float(cr[s2].y));
    for (uint32_t s2 = 32 + lane; s2 < cnt1; s2 += 32) ma = fminf(ma, __uint_as_float(cr[segcap + s2].y));
    ma = warp_min(ma);
    const float min_band = ma + 2.0f * E;

    // 2. exact distances for everything that could be the true minimum → exact (dmin, best)
    float bd = INF;
    uint32_t bj = 0xffffffffu;
    for (uint32_t u0 = 0; u0 < total; u0 += 32) {
      const uint32_t u = u0 + lane;
      const bool valid = u < total;
      uint2 e = make_uint2(0u, 0u);
      if (valid) e = cr[slot_of(u)];
      float dv = __uint_as_float(e.y);
      const bool exact = (e.x & CAND_EXACT_BIT) != 0;
      const bool want = valid && !exact && dv <= min_band;
      const uint32_t j = e.x & CAND_SLOT_MASK;
      dv = coop_exact<METRIC>(want, j, dv, x, a.C, a.ld, sm, lane, &cur.xv);
      if (want) cr[slot_of(u)] = make_uint2(e.x | CAND_EXACT_BIT, __float_as_uint(dv));
      if (valid && (exact || want) && lex_less(dv, j, bd, bj)) { bd = dv; bj = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, o);
      const uint32_t oj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (lex_less(od, oj, bd, bj)) { bd = od; bj = oj; }
    }
    if (!(bd < INF)) { bd = INF; bj = 0; }   // fold identity (0, +inf): nothing was < inf
    if (lane == 0) {
      a.best[r] = bj;
      a.dmin[r] = bd;
    }
    if (!a.want_members) {
      if (lane == 0) a.nmem[r] = 1;
      continue;
    }
    __syncwarp();

    // 3. boundary membership, decided on exact values
    const float thr = __fmul_rn(bd, a.factor);
    const float* cb = a.C + (size_t)bj * a.ld;
    bool best_listed = false;
    for (uint32_t u0 = 0; u0 < total; u0 += 32) {
      const uint32_t u = u0 + lane;
      const bool valid = u < total;
      uint2 e = make_uint2(0u, 0u);
      if (valid) e = cr[slot_of(u)];
      const uint32_t j = e.x & CAND_SLOT_MASK;
      float dv = __uint_as_float(e.y);
      bool member = false, ambiguous = false, need_cc = false;
      float cc = 0.f, lo = 0.f, hi = 0.f;
      if (valid) {
        if (j == bj) {
          member = true;
          best_listed = true;
        } else {
          const bool exact = (e.x & CAND_EXACT_BIT) != 0;
          lo = exact ? dv : dv - E;
          hi = exact ? dv : dv + E;
          need_cc = lo < thr;                         // otherwise certainly d >= thr → not a member
        }
      }
      if (a.cc) {
        if (need_cc) cc = a.cc[(size_t)bj * a.k + j];
      } else {
        cc = coop_exact<METRIC>(need_cc, j, 0.f, cb, a.C, a.ld, sm, lane);   // no k x k matrix: on demand
      }
      if (need_cc && cc >= lo) {                      // otherwise certainly cc < d → not a member
        if (hi < thr && cc >= hi) member = true;      // certain on both tests (exact values end here)
        else ambiguous = true;
      }
      dv = coop_exact<METRIC>(ambiguous, j, dv, x, a.C, a.ld, sm, lane, &cur.xv);
      if (ambiguous) member = (dv < thr) && (cc >= dv);
      if (valid) cr[slot_of(u)] = make_uint2((e.x & ~CAND_MEMBER_BIT) | (member ? CAND_MEMBER_BIT : 0u), e.y);
    }
    best_listed = __any_sync(0xffffffffu, best_listed);
    __syncwarp();

    // 4. compact the member slots to the front of the row's buffer
    uint32_t out = 0;
    if (!best_listed) {         // only when every distance was inf/NaN: members = {slot 0}
      if (lane == 0) cr[0] = make_uint2(0u, __float_as_uint(bd));
      out = 1;
      __syncwarp();
    } else {
      for (uint32_t u0 = 0; u0 < total; u0 += 32) {
        const uint32_t u = u0 + lane;
        uint2 e = make_uint2(0u, 0u);
        bool mem = false;
        if (u < total) {
          e = cr[slot_of(u)];
          mem = (e.x & CAND_MEMBER_BIT) != 0;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, mem);
        __syncwarp();
        if (mem) cr[out + __popc(bal & ((1u << lane) - 1u))] = make_uint2(e.x & CAND_SLOT_MASK, e.y);
        out += __popc(bal);
        __syncwarp();
      }
    }
    if (lane == 0) a.nmem[r] = out;
  }
}

// Brute force for rows whose candidate buffer overflowed: one CTA per row, every distance
// recomputed exactly.  PHASE 0 writes best/dmin/nmem; PHASE 1 writes the member slots.
template <int METRIC, int PHASE>
__global__ void __launch_bounds__(256)
resolve_overflow_kernel(ResolveDev a, const uint64_t* __restrict__ row_off, uint32_t* __restrict__ keys,
                        uint32_t* __restrict__ vals) {
  __shared__ float s_bd[8];
  __shared__ uint32_t s_bj[8];
  __shared__ unsigned s_cnt;
  const float INF = __int_as_float(0x7f800000);
  const uint32_t novf = *a.ovf_count;
  for (uint32_t o = blockIdx.x; o < novf; o += gridDim.x) {
    const uint32_t r = a.ovf_rows[o];
    const float* x = a.P + (size_t)r * a.ld;
    float bd = INF;
    uint32_t bj = 0xffffffffu;
    if (PHASE == 0) {
      for (uint32_t j = threadIdx.x; j < a.k; j += blockDim.x) {
        const float dv = thread_dist<METRIC>(x, a.C + (size_t)j * a.ld, a.ld);
        if (lex_less(dv, j, bd, bj)) { bd = dv; bj = j; }
      }
#pragma unroll
      for (int of = 16; of > 0; of >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, of);
        const uint32_t oj = __shfl_xor_sync(0xffffffffu, bj, of);
        if (lex_less(od, oj, bd, bj)) { bd = od; bj = oj; }
      }
      if ((threadIdx.x & 31) == 0) { s_bd[threadIdx.x >> 5] = bd; s_bj[threadIdx.x >> 5] = bj; }
      if (threadIdx.x == 0) s_cnt = 0;
      __syncthreads();
      bd = s_bd[0]; bj = s_bj[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (lex_less(s_bd[w], s_bj[w], bd, bj)) { bd = s_bd[w]; bj = s_bj[w]; }
      if (!(bd < INF)) { bd = INF; bj = 0; }
      if (threadIdx.x == 0) { a.best[r] = bj; a.dmin[r] = bd; }
    } else {
      bd = a.dmin[r];
      bj = a.best[r];
      if (threadIdx.x == 0) s_cnt = 0;
      __syncthreads();
    }
    const float thr = __fmul_rn(bd, a.factor);
    const float* cb = a.C + (size_t)bj * a.ld;
    for (uint32_t j = threadIdx.x; j < a.k; j += blockDim.x) {
      bool member = (j == bj);
      if (!member && a.want_members) {
        const float dv = thread_dist<METRIC>(x, a.C + (size_t)j * a.ld, a.ld);
        if (dv < thr) {
          const float cc = a.cc ? a.cc[(size_t)bj * a.k + j]
                                : thread_dist<METRIC>(cb, a.C + (size_t)j * a.ld, a.ld);
          member = cc >= dv;
        }
      }
      if (member) {
        const unsigned pos = atomicAdd(&s_cnt, 1u);
        if (PHASE == 1) {
          keys[row_off[r] + pos] = j;
          vals[row_off[r] + pos] = r;
        }
      }
    }
    __syncthreads();
    if (PHASE == 0 && threadIdx.x == 0) a.nmem[r] = s_cnt | NMEM_OVERFLOW_BIT;
    __syncthreads();
  }
}

struct CountOp {
  __host__ __device__ uint64_t operator()(uint32_t v) const { return (uint64_t)(v & ~NMEM_OVERFLOW_BIT); }
};

__global__ void fill_pairs_kernel(const uint2* __restrict__ cand, int cap, const uint32_t* __restrict__ nmem,
                                  const uint64_t* __restrict__ row_off, uint32_t m,
                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int lane = threadIdx.x & 31;
  const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < m; r += warps_total) {
    const uint32_t nm = nmem[r];
    if (nm & NMEM_OVERFLOW_BIT) continue;
    const uint64_t off = row_off[r];
    const uint2* cr = cand + (size_t)r * cap;
    for (uint32_t s = lane; s < nm; s += 32) {
      keys[off + s] = cr[s].x & CAND_SLOT_MASK;
      vals[off + s] = r;
    }
  }
}

__global__ void total_kernel(const uint64_t* row_off, const uint32_t* nmem, uint32_t m, uint64_t* total) {
  if (threadIdx.x == 0 && blockIdx.x == 0)
    total[0] = m ? row_off[m - 1] + (uint64_t)(nmem[m - 1] & ~NMEM_OVERFLOW_BIT) : 0;
}

__global__ void offsets_kernel(const uint32_t* __restrict__ keys_sorted, uint64_t total, uint32_t k,
                               uint64_t* __restrict__ offsets) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > k) return;
  uint64_t lo = 0, hi = total;   // first position with key >= c
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (keys_sorted[mid] < c) lo = mid + 1; else hi = mid;
  }
  offsets[c] = lo;
}

template <int METRIC>
int run_resolve_t(spf_ctx* c, const ResolveArgs& a, CsrOut* csr) {
  cudaStream_t st = c->stream;
  DevBuf<uint32_t> ovf_rows, ovf_count;
  SPF_TRY(ovf_rows.alloc(st, a.m));
  SPF_TRY(ovf_count.alloc(st, 1));
  SPF_CUDA(cudaMemsetAsync(ovf_count.p, 0, sizeof(uint32_t), st));
  ResolveDev d{a.P, (uint32_t)a.m, a.C, a.k, a.ld, a.factor, a.cand, a.cand_cnt, a.cap, a.nseg, a.xnorm,
               a.d_cnmax, a.cc, a.best, a.dmin, a.nmem, ovf_rows.p, ovf_count.p, a.want_members ? 1 : 0};
  const unsigned ovf_grid = (unsigned)c->sm_count * 4;
  {
    KernelTimer t(c, "resolve");
    uint64_t blocks = ceil_div(a.m, RS_WARPS);
    if (blocks > (uint64_t)c->sm_count * 16) blocks = (uint64_t)c->sm_count * 16;
    resolve_kernel<METRIC><<<(unsigned)blocks, RS_WARPS * 32, 0, st>>>(d);
    SPF_TRY(check_launch(c, "resolve_kernel"));
    resolve_overflow_kernel<METRIC, 0><<<ovf_grid, 256, 0, st>>>(d, nullptr, nullptr, nullptr);
    SPF_TRY(check_launch(c, "resolve_overflow_kernel<0>"));
  }
  if (!csr) return SPF_OK;

  KernelTimer t(c, "csr");
  // row offsets = exclusive scan of member counts
  DevBuf<uint64_t> row_off, d_total;
  SPF_TRY(row_off.alloc(st, a.m));
  SPF_TRY(d_total.alloc(st, 1));
  cub::TransformInputIterator<uint64_t, CountOp, const uint32_t*> counts(a.nmem, CountOp());
  size_t tmp_bytes = 0;
  SPF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, row_off.p, (int64_t)a.m, st));
  DevBuf<uint8_t> tmp;
  SPF_TRY(tmp.alloc(st, tmp_bytes));
  SPF_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, counts, row_off.p, (int64_t)a.m, st));
  c->launches += 2;
  total_kernel<<<1, 32, 0, st>>>(row_off.p, a.nmem, (uint32_t)a.m, d_total.p);
  SPF_TRY(check_launch(c, "total_kernel"));
  uint64_t total = 0;
  SPF_CUDA(cudaMemcpyAsync(&total, d_total.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));

  DevBuf<uint32_t> keys, vals, keys2, vals2;
  SPF_TRY(keys.alloc(st, total));
  SPF_TRY(vals.alloc(st, total));
  SPF_TRY(keys2.alloc(st, total));
  SPF_TRY(vals2.alloc(st, total));
  DevBuf<uint64_t> offsets;
  SPF_TRY(offsets.alloc(st, (size_t)a.k + 1));
  {
    uint64_t blocks = ceil_div(a.m * 32, 256);
    if (blocks > (uint64_t)c->sm_count * 16) blocks = (uint64_t)c->sm_count * 16;
    fill_pairs_kernel<<<(unsigned)blocks, 256, 0, st>>>(a.cand, a.cap, a.nmem, row_off.p, (uint32_t)a.m,
                                                        keys.p, vals.p);
    SPF_TRY(check_launch(c, "fill_pairs_kernel"));
    resolve_overflow_kernel<METRIC, 1><<<ovf_grid, 256, 0, st>>>(d, row_off.p, keys.p, vals.p);
    SPF_TRY(check_launch(c, "resolve_overflow_kernel<1>"));
  }
  // stable sort by cluster slot keeps the input order inside every cluster (:353-361)
  int end_bit = 1;
  while (end_bit < 32 && (1ull << end_bit) < (uint64_t)a.k) ++end_bit;
  if (total > 0) {
    size_t sort_bytes = 0;
    SPF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys.p, keys2.p, vals.p, vals2.p,
                                             (int64_t)total, 0, end_bit, st));
    DevBuf<uint8_t> stmp;
    SPF_TRY(stmp.alloc(st, sort_bytes));
    SPF_CUDA(cub::DeviceRadixSort::SortPairs(stmp.p, sort_bytes, keys.p, keys2.p, vals.p, vals2.p,
                                             (int64_t)total, 0, end_bit, st));
    c->launches += 4;
  }
  offsets_kernel<<<(unsigned)ceil_div((uint64_t)a.k + 1, 256), 256, 0, st>>>(keys2.p, total, a.k, offsets.p);
  SPF_TRY(check_launch(c, "offsets_kernel"));
  csr->total = total;
  csr->offsets = offsets.take();
  csr->members = vals2.take();
  return SPF_OK;
}

}  // namespace

int run_resolve(spf_ctx* c, const ResolveArgs& a, CsrOut* csr) {
  switch (a.metric) {
    case SPF_METRIC_EUCLIDEAN: return run_resolve_t<SPF_METRIC_EUCLIDEAN>(c, a, csr);
    case SPF_METRIC_MANHATTAN: return run_resolve_t<SPF_METRIC_MANHATTAN>(c, a, csr);
    case SPF_METRIC_CHEBYSHEV: return run_resolve_t<SPF_METRIC_CHEBYSHEV>(c, a, csr);
  }
  return fail(SPF_E_INVALID, "unknown metric %d", a.metric);
}

}  // namespace spf
