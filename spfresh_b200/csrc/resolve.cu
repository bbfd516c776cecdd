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
  uint2* work; uint32_t* work_count; uint32_t work_cap;   // (row, slot) pairs needing an exact distance
  int want_members;
};

__device__ __forceinline__ bool lex_less(float d1, uint32_t j1, float d2, uint32_t j2) {
  return d1 < d2 || (d1 == d2 && j1 < j2);
}

// Warp-aggregated append of (row, slot) to the work list; returns false when the list is full
// (the caller then evaluates the distance inline, so the list is only an accelerator).
__device__ __forceinline__ bool work_push(const ResolveDev& a, bool want, uint32_t row, uint32_t slot, int lane) {
  const unsigned bal = __ballot_sync(0xffffffffu, want);
  if (bal == 0) return true;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(a.work_count, (uint32_t)__popc(bal));
  base = __shfl_sync(0xffffffffu, base, 0);
  const uint32_t pos = base + __popc(bal & ((1u << lane) - 1u));
  if (want && pos < a.work_cap) a.work[pos] = make_uint2(row, slot);
  return !want || pos < a.work_cap;
}

enum { MODE_MIN = 0, MODE_CLASSIFY = 1, MODE_FINAL = 2 };

// One warp per listed point.
//   MODE_MIN      (tensor path) approximate minimum, queue every candidate within 2E of it
//   MODE_CLASSIFY (tensor path) exact (dmin, best); queue the boundary candidates whose
//                 membership cannot be certified from the approximate distance
//   MODE_FINAL    everything decided on exact values; members compacted to the row's front
// Anything that should have been queued but did not fit is evaluated inline by its lane.
template <int METRIC, int MODE>
__global__ void __launch_bounds__(256) resolve_kernel(ResolveDev a) {
  const int lane = threadIdx.x & 31;
  const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
  const float INF = __int_as_float(0x7f800000);
  const uint32_t segcap = (uint32_t)a.cap / (uint32_t)a.nseg;
  for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < a.m; r += warps_total) {
    // the row's buffer holds nseg segments of segcap slots; cnt_s > segcap marks an overflow
    uint32_t cnts[2];
    cnts[0] = a.cand_cnt[(size_t)r * a.nseg];
    cnts[1] = a.nseg > 1 ? a.cand_cnt[(size_t)r * a.nseg + 1] : 0;
    if (cnts[0] > segcap || cnts[1] > segcap) {   // overflowed: the brute-force kernels own this row
      if (MODE == MODE_MIN || (MODE == MODE_FINAL && a.work == nullptr)) {
        if (lane == 0) {
          const uint32_t pos = atomicAdd(a.ovf_count, 1u);
          a.ovf_rows[pos] = r;
          a.nmem[r] = NMEM_OVERFLOW_BIT;
        }
      }
      continue;
    }
    uint2* cr = a.cand + (size_t)r * a.cap;
    const float* x = a.P + (size_t)r * a.ld;
    const float E = a.xnorm ? tc_err_bound(a.xnorm[r], a.cnmax[0], a.ld) : 0.0f;

    float bd = INF;
    uint32_t bj = 0xffffffffu;
    if (MODE == MODE_MIN) {
      float ma = INF;
      for (int sg = 0; sg < a.nseg; ++sg)
        for (uint32_t s = lane; s < cnts[sg]; s += 32) ma = fminf(ma, __uint_as_float(cr[sg * segcap + s].y));
      ma = warp_min(ma);
      const float min_band = ma + 2.0f * E;
      for (int sg = 0; sg < a.nseg; ++sg)
        for (uint32_t s0 = 0; s0 < cnts[sg]; s0 += 32) {
          const uint32_t s = s0 + lane;
          bool want = false;
          if (s < cnts[sg]) {
            const uint2 e = cr[sg * segcap + s];
            want = !(e.x & CAND_EXACT_BIT) && __uint_as_float(e.y) <= min_band;
          }
          work_push(a, want, r, sg * segcap + s, lane);   // leftovers are caught inline in CLASSIFY
        }
      if (lane == 0) a.dmin[r] = min_band;                // handed to MODE_CLASSIFY
      continue;
    }

    // ---- exact (dmin, best) -------------------------------------------------------------------
    if (MODE == MODE_FINAL && a.work != nullptr) {
      bd = a.dmin[r];
      bj = a.best[r];
    } else {
      const float min_band = (MODE == MODE_CLASSIFY) ? a.dmin[r] : INF;   // exact path: all exact already
      for (int sg = 0; sg < a.nseg; ++sg)
        for (uint32_t s = lane; s < cnts[sg]; s += 32) {
          uint2 e = cr[sg * segcap + s];
          float dv = __uint_as_float(e.y);
          if (!(e.x & CAND_EXACT_BIT)) {
            if (!(dv <= min_band)) continue;
            dv = thread_dist<METRIC>(x, a.C + (size_t)(e.x & CAND_SLOT_MASK) * a.ld, a.ld);   // list was full
            e.x |= CAND_EXACT_BIT;
            e.y = __float_as_uint(dv);
            cr[sg * segcap + s] = e;
          }
          const uint32_t j = e.x & CAND_SLOT_MASK;
          if (lex_less(dv, j, bd, bj)) { bd = dv; bj = j; }
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, o);
        const uint32_t oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (lex_less(od, oj, bd, bj)) { bd = od; bj = oj; }
      }
      if (!(bd < INF)) { bd = INF; bj = 0; }   // fold identity (0, +inf): nothing was < inf
      __syncwarp();
      if (lane == 0) {
        a.best[r] = bj;
        a.dmin[r] = bd;
      }
    }
    if (!a.want_members) {
      if (MODE == MODE_FINAL && lane == 0) a.nmem[r] = 1;
      continue;
    }

    // ---- boundary membership -------------------------------------------------------------------
    const float thr = __fmul_rn(bd, a.factor);
    const float* cb = a.C + (size_t)bj * a.ld;
    bool best_listed = false;
    for (int sg = 0; sg < a.nseg; ++sg)
      for (uint32_t s0 = 0; s0 < cnts[sg]; s0 += 32) {
        const uint32_t s = s0 + lane;
        const bool valid = s < cnts[sg];
        uint2 e = make_uint2(0u, 0u);
        bool member = false, ambiguous = false;
        float cc = 0.f;
        if (valid) {
          e = cr[sg * segcap + s];
          const uint32_t j = e.x & CAND_SLOT_MASK;
          const float dv = __uint_as_float(e.y);
          if (j == bj) {
            member = true;
            best_listed = true;
          } else {
            const bool exact = (e.x & CAND_EXACT_BIT) != 0;
            const float lo = exact ? dv : dv - E, hi = exact ? dv : dv + E;
            if (lo < thr) {                         // otherwise certainly d >= thr → not a member
              cc = a.cc ? a.cc[(size_t)bj * a.k + j] : thread_dist<METRIC>(cb, a.C + (size_t)j * a.ld, a.ld);
              if (cc >= lo) {                       // otherwise certainly cc < d → not a member
                if (hi < thr && cc >= hi) member = true;   // certain on both tests (exact ones end here)
                else ambiguous = true;
              }
            }
          }
        }
        if (MODE == MODE_CLASSIFY) {
          work_push(a, ambiguous, r, sg * segcap + s, lane);
        } else {
          if (ambiguous) {                          // not queued (list full): decide inline
            const float dv = thread_dist<METRIC>(x, a.C + (size_t)(e.x & CAND_SLOT_MASK) * a.ld, a.ld);
            member = (dv < thr) && (cc >= dv);
          }
          if (valid) {
            e.x = (e.x & ~CAND_MEMBER_BIT) | (member ? CAND_MEMBER_BIT : 0u);
            cr[sg * segcap + s] = e;
          }
        }
      }
    if (MODE == MODE_CLASSIFY) continue;
    best_listed = __any_sync(0xffffffffu, best_listed);
    __syncwarp();

    // ---- compact the member slots to the front of the row's buffer ------------------------------
    uint32_t out = 0;
    if (!best_listed) {         // only when every distance was inf/NaN: members = {slot 0}
      if (lane == 0) cr[0] = make_uint2(0u, __float_as_uint(bd));
      out = 1;
      __syncwarp();
    } else {
      for (int sg = 0; sg < a.nseg; ++sg)
        for (uint32_t s0 = 0; s0 < cnts[sg]; s0 += 32) {
          const uint32_t s = s0 + lane;
          uint2 e = make_uint2(0u, 0u);
          bool mem = false;
          if (s < cnts[sg]) {
            e = cr[sg * segcap + s];
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

// Exact distances for the queued (row, slot) pairs: 32 pairs per warp, rows fetched with
// coalesced 128-byte requests (pairdist.cuh), result written back into the candidate record.
template <int METRIC>
__global__ void __launch_bounds__(PD_THREADS) work_exact_kernel(ResolveDev a) {
  __shared__ PairDistSmem sm[PD_THREADS / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t count = *a.work_count;
  if (count > a.work_cap) count = a.work_cap;
  const uint32_t nwarps = gridDim.x * (PD_THREADS / 32);
  for (uint32_t base = (blockIdx.x * (PD_THREADS / 32) + warp) * 32; base < count; base += nwarps * 32) {
    const uint32_t i = base + lane;
    const bool valid = i < count;
    const float* pa = nullptr;
    const float* pb = nullptr;
    uint2* rec = nullptr;
    if (valid) {
      const uint2 w = a.work[i];
      rec = a.cand + (size_t)w.x * a.cap + w.y;
      pa = a.P + (size_t)w.x * a.ld;
      pb = a.C + (size_t)(rec->x & CAND_SLOT_MASK) * a.ld;
    }
    const float dv = warp_pair_dist<METRIC>(pa, pb, a.ld, sm[warp]);
    if (valid) *rec = make_uint2(rec->x | CAND_EXACT_BIT, __float_as_uint(dv));
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
  // work list for exact re-evaluation (tensor path only)
  DevBuf<uint2> work;
  DevBuf<uint32_t> work_count;
  const bool approx = a.xnorm != nullptr;
  const uint32_t work_cap = approx ? (uint32_t)std::min<uint64_t>(a.m * 8ull + 1024, 1ull << 30) : 0;
  if (approx) {
    SPF_TRY(work.alloc(st, work_cap));
    SPF_TRY(work_count.alloc(st, 1));
  }
  ResolveDev d{a.P, (uint32_t)a.m, a.C, a.k, a.ld, a.factor, a.cand, a.cand_cnt, a.cap, a.nseg, a.xnorm,
               a.d_cnmax, a.cc, a.best, a.dmin, a.nmem, ovf_rows.p, ovf_count.p,
               approx ? work.p : nullptr, approx ? work_count.p : nullptr, work_cap, a.want_members ? 1 : 0};
  const unsigned ovf_grid = (unsigned)c->sm_count * 4;
  {
    KernelTimer t(c, "resolve");
    uint64_t blocks = ceil_div(a.m * 32, 256);
    if (blocks > (uint64_t)c->sm_count * 16) blocks = (uint64_t)c->sm_count * 16;
    const unsigned pgrid = (unsigned)c->sm_count * 16;
    if (approx) {
      SPF_CUDA(cudaMemsetAsync(work_count.p, 0, sizeof(uint32_t), st));
      resolve_kernel<METRIC, MODE_MIN><<<(unsigned)blocks, 256, 0, st>>>(d);
      SPF_TRY(check_launch(c, "resolve_kernel<MIN>"));
      work_exact_kernel<METRIC><<<pgrid, PD_THREADS, 0, st>>>(d);
      SPF_TRY(check_launch(c, "work_exact_kernel"));
      SPF_CUDA(cudaMemsetAsync(work_count.p, 0, sizeof(uint32_t), st));
      resolve_kernel<METRIC, MODE_CLASSIFY><<<(unsigned)blocks, 256, 0, st>>>(d);
      SPF_TRY(check_launch(c, "resolve_kernel<CLASSIFY>"));
      if (a.want_members) {
        work_exact_kernel<METRIC><<<pgrid, PD_THREADS, 0, st>>>(d);
        SPF_TRY(check_launch(c, "work_exact_kernel"));
      }
    }
    resolve_kernel<METRIC, MODE_FINAL><<<(unsigned)blocks, 256, 0, st>>>(d);
    SPF_TRY(check_launch(c, "resolve_kernel<FINAL>"));
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
