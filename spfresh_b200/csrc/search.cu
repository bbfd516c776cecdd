// search.cu — HBM-resident posting lists and the batched query path.
//
// Replaces SpannIndex::find_k_nearest_neighbor_spann, src/spann/spann_index.rs:148-197
// (reference): kd-tree probe (:164) → exact squared-L2 of the query to every centroid + sorted
// top-nprobe; per-probe file read + decode (src/spann/posting_lists.rs:98-106) → lists resident
// in HBM; per-point distance + `<= thr` filter (:170-179) + stable sort/truncate (:188-193) →
// one streaming pass with a per-query top-k kept in registers.
//
// HBM layout of a posting list ("slots"): vectors are stored in groups of 32, dimension-chunk
// major: group g, chunk c (4 dims), lane l → float4 at ((g * ld/4 + c) * 32 + l).  A warp that
// owns a group reads one fully coalesced 512-byte line per chunk while every lane walks the
// dimensions of its own vector in order, so each distance is the reference's sequential f32
// sum.  ids are stored per slot as well (padded slots hold UINT64_MAX).
#include <errno.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>

#include <string>
#include <vector>

#include <cub/cub.cuh>

#include <algorithm>

#include "comm.cuh"
#include "kernels.cuh"

using namespace spf;

struct spf_index {
  spf_ctx* ctx = nullptr;
  uint32_t d = 0, ld = 0, nlists = 0, list_begin = 0, list_end = 0;
  float* centroids = nullptr;    // nlists x ld (every list, also the ones other ranks own)
  float* vecs = nullptr;         // total_groups * 32 * ld floats, slot layout
  uint64_t* slot_ids = nullptr;  // total_groups * 32
  uint64_t* grp_off = nullptr;   // device nlists+1, group offsets (non-local lists are empty)
  uint32_t* lens = nullptr;      // device nlists, GLOBAL list lengths (encounter index needs all)
  std::vector<uint64_t> h_grp_off;
  std::vector<uint32_t> h_lens;
  uint64_t total_groups = 0, total_vectors = 0;
  uint64_t last_scan_bytes = 0;
  spf::ScanTcSide tc;            // TF32 side structures of the tensor-core scan, made on first use
  // the centroids as one posting list (slot layout) for the tensor-core probe, made on first use
  float* cvecs = nullptr;
  uint64_t* cids = nullptr;
  uint64_t* cgrp = nullptr;      // device {0, groups}
  uint32_t* clens = nullptr;     // device {nlists}
  spf::ScanTcSide ctc;
};

namespace spf {
namespace {

__global__ void pack_lists_kernel(const float* __restrict__ X, uint32_t ld4, const uint64_t* __restrict__ rows,
                                  const uint64_t* __restrict__ loc_off,   // local lists' member offsets (nloc+1)
                                  const uint64_t* __restrict__ grp_off,   // group offsets of local lists (nloc+1)
                                  float* __restrict__ vecs, uint64_t* __restrict__ slot_ids) {
  const uint32_t l = blockIdx.x;
  const uint64_t b = loc_off[l], e = loc_off[l + 1];
  const uint64_t g0 = grp_off[l];
  const float4* X4 = reinterpret_cast<const float4*>(X);
  float4* V4 = reinterpret_cast<float4*>(vecs);
  const uint64_t work = (e - b) * ld4;
  for (uint64_t w = threadIdx.x; w < work; w += blockDim.x) {
    const uint64_t pos = w / ld4;
    const uint32_t c = (uint32_t)(w - pos * ld4);
    const uint64_t row = rows[b + pos];
    const uint64_t g = g0 + (pos >> 5);
    V4[(g * ld4 + c) * 32 + (pos & 31)] = __ldg(X4 + (size_t)row * ld4 + c);
    if (c == 0) slot_ids[g * 32 + (pos & 31)] = row;
  }
}

// Sorted top-nprobe centroids per query by (distance, list id), the prune threshold and the
// encounter-index base of each probed list.  One CTA per query.
__global__ void __launch_bounds__(128)
probe_select_kernel(const float* __restrict__ Dqc, uint32_t nlists, uint32_t nprobe, float prune_factor,
                    const uint32_t* __restrict__ lens, uint32_t* __restrict__ probe, float* __restrict__ thr,
                    uint32_t* __restrict__ seqbase) {
  __shared__ unsigned long long s_red[4];
  __shared__ unsigned long long s_last;
  const uint64_t q = blockIdx.x;
  const float* row = Dqc + q * nlists;
  unsigned long long last = 0;
  bool have_last = false;
  uint32_t seq = 0;
  for (uint32_t p = 0; p < nprobe; ++p) {
    unsigned long long best = ~0ull;
    for (uint32_t j = threadIdx.x; j < nlists; j += blockDim.x) {
      const unsigned long long key = ((unsigned long long)__float_as_uint(row[j]) << 32) | j;
      if ((!have_last || key > last) && key < best) best = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, o);
      if (ob < best) best = ob;
    }
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long b = s_red[0];
      for (int w = 1; w < 4; ++w) if (s_red[w] < b) b = s_red[w];
      s_last = b;
      const uint32_t j = (uint32_t)(b & 0xffffffffull);
      probe[q * nprobe + p] = j;
      seqbase[q * nprobe + p] = seq;
      if (p == 0) {
        // :165  F::from(1.2) * (nearest.distance + F::epsilon())
        const float d0 = __uint_as_float((uint32_t)(b >> 32));
        thr[q] = __fmul_rn(prune_factor, __fadd_rn(d0, 1.1920929e-7f));
      }
    }
    __syncthreads();
    last = s_last;
    have_last = true;
    seq += lens[(uint32_t)(last & 0xffffffffull)];
    __syncthreads();
  }
}

// The same selection by a full block-wide radix sort of the (distance bits << 32 | list id) keys of
// one query (nlists <= 256 * ITEMS): O(nlists) work per query instead of O(nprobe * nlists), which
// is what matters for nprobe in the tens to hundreds.  Same total order, same outputs.
template <int ITEMS>
__global__ void __launch_bounds__(256)
probe_sort_kernel(const float* __restrict__ Dqc, uint32_t nlists, uint32_t nprobe, float prune_factor,
                  const uint32_t* __restrict__ lens, uint32_t* __restrict__ probe, float* __restrict__ thr,
                  uint32_t* __restrict__ seqbase) {
  typedef cub::BlockRadixSort<unsigned long long, 256, ITEMS> Sort;
  typedef cub::BlockScan<uint32_t, 256> Scan;
  __shared__ union { typename Sort::TempStorage sort; typename Scan::TempStorage scan; } tmp;
  const uint64_t q = blockIdx.x;
  const float* row = Dqc + q * nlists;
  unsigned long long key[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t j = threadIdx.x * ITEMS + i;     // blocked arrangement
    // (distance bits, list id) packed into 44 bits: fewer radix passes (list id < 256 * ITEMS <= 4096)
    key[i] = j < nlists ? (((unsigned long long)__float_as_uint(row[j]) << 12) | j) : ~0ull;
  }
  Sort(tmp.sort).Sort(key, 0, 44);
  __syncthreads();
  // encounter-index base of every selected list = exclusive prefix of the lengths in probe order
  uint32_t len[ITEMS], base[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t p = threadIdx.x * ITEMS + i;
    len[i] = (p < nprobe && key[i] != ~0ull) ? lens[(uint32_t)(key[i] & 0xfffull)] : 0u;
  }
  Scan(tmp.scan).ExclusiveSum(len, base);
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t p = threadIdx.x * ITEMS + i;
    if (p < nprobe) {
      probe[q * nprobe + p] = (uint32_t)(key[i] & 0xfffull);
      seqbase[q * nprobe + p] = base[i];
    }
  }
  if (threadIdx.x == 0) {
    // :165  F::from(1.2) * (nearest.distance + F::epsilon())
    const float d0 = __uint_as_float((uint32_t)(key[0] >> 12));
    thr[q] = __fmul_rn(prune_factor, __fadd_rn(d0, 1.1920929e-7f));
  }
}

// The same outputs for nprobe << nlists without sorting everything: a bitwise search finds the
// nprobe-th smallest (distance bits, list id) key of the query (one counting step over the keys in
// registers per differing distance bit, 12 more over the list ids only when equal distances straddle
// the cut), the nprobe keys up to it are compacted into shared memory and only those are sorted.
// Keys are unique, so exactly nprobe keys are selected.  ITEMS: keys per thread (nlists <= 256 *
// ITEMS), OUT: selected keys per thread (nprobe <= 256 * OUT).
template <int ITEMS, int OUT>
__global__ void __launch_bounds__(256)
probe_topn_kernel(const float* __restrict__ Dqc, uint32_t nlists, uint32_t nprobe, float prune_factor,
                  const uint32_t* __restrict__ lens, uint32_t* __restrict__ probe, float* __restrict__ thr,
                  uint32_t* __restrict__ seqbase) {
  typedef cub::BlockRadixSort<unsigned long long, 256, OUT> Sort;
  typedef cub::BlockScan<uint32_t, 256> Scan;
  __shared__ union { typename Sort::TempStorage sort; typename Scan::TempStorage scan; } tmp;
  __shared__ unsigned long long s_sel[256 * OUT];
  __shared__ uint32_t s_cnt[2][8];
  const uint64_t q = blockIdx.x;
  const float* row = Dqc + q * nlists;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // striped keys (coalesced loads; the order is irrelevant here): distance bits in registers, the
  // list id j = i * 256 + threadIdx.x implicit.  Pad keys (j >= nlists) are (0xffffffff, j): last.
  uint32_t dv[ITEMS];
  uint32_t vo = 0u, va = ~0u;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t j = i * 256 + threadIdx.x;
    dv[i] = j < nlists ? __float_as_uint(row[j]) : ~0u;
    vo |= dv[i];
    va &= dv[i];
  }
  auto block_sum = [&](uint32_t c, int slot) {       // sum over the block; slots alternate between calls
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) s_cnt[slot][warp] = c;
    __syncthreads();
    uint32_t tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s_cnt[slot][w];
    return tot;
  };
  // bits that all keys share need no counting
  vo = __reduce_or_sync(0xffffffffu, vo);
  va = __reduce_and_sync(0xffffffffu, va);
  if (lane == 0) { s_cnt[0][warp] = vo; s_cnt[1][warp] = va; }
  __syncthreads();
  vo = 0u; va = ~0u;
#pragma unroll
  for (int w = 0; w < 8; ++w) { vo |= s_cnt[0][w]; va &= s_cnt[1][w]; }
  __syncthreads();
  const uint32_t diff = vo ^ va;
  const int hb = diff ? 31 - __clz((int)diff) : -1;
  // distance bits of the nprobe-th smallest key, most significant differing bit first: with the bits
  // above `bit` fixed to `prefix`, count the keys whose bit is 0; the wanted key is among them if at
  // least `want` keys are
  uint32_t prefix = hb >= 31 ? 0u : (hb < 0 ? va : (va & ~((2u << hb) - 1u)));
  uint32_t want = nprobe;
  for (int bit = hb; bit >= 0; --bit) {
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) c += ((dv[i] ^ prefix) >> bit) == 0u ? 1u : 0u;
    const uint32_t tot = block_sum(c, bit & 1);
    if (want > tot) { want -= tot; prefix |= 1u << bit; }
  }
  // keys with smaller distance bits are all selected; of the `ties` keys with equal bits the `want`
  // smallest list ids are (usually ties == want == 1)
  uint32_t tc = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) tc += dv[i] == prefix ? 1u : 0u;
  __syncthreads();
  const uint32_t ties = block_sum(tc, 0);
  uint32_t jmax = 0xffffffffu;
  if (ties != want) {                                 // the want-th smallest id among the ties, bitwise again
    uint32_t jp = 0;
    __syncthreads();
    for (int bit = 11; bit >= 0; --bit) {
      uint32_t c = 0;
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const uint32_t j = i * 256 + threadIdx.x;
        c += (dv[i] == prefix && ((j ^ jp) >> bit) == 0u) ? 1u : 0u;
      }
      const uint32_t tot = block_sum(c, bit & 1);
      if (want > tot) { want -= tot; jp |= 1u << bit; }
    }
    jmax = jp;
  }
  // compact the selected keys (exactly nprobe of them) and sort those
  uint32_t mine = 0, base0 = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t j = i * 256 + threadIdx.x;
    mine += (dv[i] < prefix || (dv[i] == prefix && j <= jmax)) ? 1u : 0u;
  }
  __syncthreads();
  Scan(tmp.scan).ExclusiveSum(mine, base0);
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t j = i * 256 + threadIdx.x;
    if (dv[i] < prefix || (dv[i] == prefix && j <= jmax))
      s_sel[base0++] = j < nlists ? (((unsigned long long)dv[i] << 12) | j) : ~0ull;
  }
  __syncthreads();
  unsigned long long sk[OUT];
#pragma unroll
  for (int i = 0; i < OUT; ++i) {
    const uint32_t p = threadIdx.x * OUT + i;       // blocked arrangement
    sk[i] = p < nprobe ? s_sel[p] : ~0ull;
  }
  __syncthreads();
  Sort(tmp.sort).Sort(sk, 0, 44);
  __syncthreads();
  uint32_t len[OUT], base[OUT];
#pragma unroll
  for (int i = 0; i < OUT; ++i) {
    const uint32_t p = threadIdx.x * OUT + i;
    len[i] = (p < nprobe && sk[i] != ~0ull) ? lens[(uint32_t)(sk[i] & 0xfffull)] : 0u;
  }
  Scan(tmp.scan).ExclusiveSum(len, base);
#pragma unroll
  for (int i = 0; i < OUT; ++i) {
    const uint32_t p = threadIdx.x * OUT + i;
    if (p < nprobe) {
      probe[q * nprobe + p] = (uint32_t)(sk[i] & 0xfffull);
      seqbase[q * nprobe + p] = base[i];
    }
  }
  if (threadIdx.x == 0) {
    // :165  F::from(1.2) * (nearest.distance + F::epsilon())
    const float d0 = __uint_as_float((uint32_t)(sk[0] >> 12));
    thr[q] = __fmul_rn(prune_factor, __fadd_rn(d0, 1.1920929e-7f));
  }
}

template <int R>
__device__ __forceinline__ void topk_insert(unsigned long long (&key)[R], unsigned long long (&pay)[R],
                                            unsigned long long ck, unsigned long long cp, int lane) {
  int pos = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) pos += __popc(__ballot_sync(0xffffffffu, key[r] < ck));
#pragma unroll
  for (int r = R - 1; r >= 0; --r) {
    unsigned long long uk = __shfl_up_sync(0xffffffffu, key[r], 1);
    unsigned long long up = __shfl_up_sync(0xffffffffu, pay[r], 1);
    if (r > 0) {
      const unsigned long long pk = __shfl_sync(0xffffffffu, key[r - 1], 31);
      const unsigned long long pp = __shfl_sync(0xffffffffu, pay[r - 1], 31);
      if (lane == 0) { uk = pk; up = pp; }
    }
    const int e = r * 32 + lane;
    if (e > pos) { key[r] = uk; pay[r] = up; }
    else if (e == pos) { key[r] = ck; pay[r] = cp; }
  }
}

template <int R>
__device__ __forceinline__ unsigned long long topk_kth(const unsigned long long (&key)[R], uint32_t K) {
  unsigned long long v = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const unsigned long long t = __shfl_sync(0xffffffffu, key[r], (K - 1) & 31);
    if ((int)((K - 1) >> 5) == r) v = t;
  }
  return v;
}


constexpr int SCAN_WARPS = 8;

// One CTA per query.  Warps take (probed list, group) units round-robin; a lane computes the
// exact squared L2 of its slot's vector (spann_index.rs:172), keeps it if <= thr (:176) and
// the warp maintains the K smallest (distance, encounter index) keys.
template <int R>
__device__ __forceinline__ void scan_query(const ScanArgs& a, const uint64_t q) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_q = reinterpret_cast<float4*>(smem_raw);                       // ld/4 float4
  uint32_t* s_gpre = reinterpret_cast<uint32_t*>(s_q + a.ld / 4);          // nprobe+1
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(
      (reinterpret_cast<uintptr_t>(s_gpre + a.nprobe + 1) + 7) & ~(uintptr_t)7);   // SCAN_WARPS*K
  unsigned long long* s_pay = s_keys + SCAN_WARPS * a.K;

  if (a.only != nullptr && a.only[q] == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ld4 = a.ld / 4;
  const uint32_t* probe = a.probe + q * a.nprobe;
  const uint32_t* seqb = a.seqbase + q * a.nprobe;
  for (uint32_t c = threadIdx.x; c < ld4; c += blockDim.x)
    s_q[c] = reinterpret_cast<const float4*>(a.Q + q * a.ld)[c];
  if (threadIdx.x == 0) {
    uint32_t acc = 0;
    unsigned long long bytes = 0;
    for (uint32_t p = 0; p < a.nprobe; ++p) {
      s_gpre[p] = acc;
      const uint32_t l = probe[p];
      const uint32_t ng = (uint32_t)(a.grp_off[l + 1] - a.grp_off[l]);   // 0 for lists of other ranks
      acc += ng;
      if (ng) bytes += (unsigned long long)a.lens[l] * a.d * 4ull;
    }
    s_gpre[a.nprobe] = acc;
    if (bytes && a.out_col == 0) atomicAdd(a.bytes, bytes);
  }
  __syncthreads();
  const float thr = a.thr[q];
  const uint32_t units = s_gpre[a.nprobe];
  const uint32_t ostride = a.out_stride ? a.out_stride : a.K;
  // a later pass of a K > 128 search: only what lies behind the previous pass's last result (~0: it was not full)
  const unsigned long long after = a.out_col ? a.out_keys[q * ostride + a.out_col - 1] : 0ull;
  const bool later = a.out_col != 0;

  unsigned long long key[R], pay[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { key[r] = ~0ull; pay[r] = ~0ull; }
  unsigned long long kth = ~0ull;

  uint32_t p = 0;
  const float4* V4 = reinterpret_cast<const float4*>(a.vecs);
  for (uint32_t u = warp; u < units; u += SCAN_WARPS) {
    while (u >= s_gpre[p + 1]) ++p;
    const uint32_t l = probe[p];
    const uint32_t g = u - s_gpre[p];
    const uint64_t G = a.grp_off[l] + g;
    const float4* base = V4 + G * ld4 * 32 + lane;
    float acc = 0.0f;
#pragma unroll 8
    for (uint32_t c = 0; c < ld4; ++c) {
      const float4 v = __ldg(base + (size_t)c * 32);
      const float4 qv = s_q[c];
      acc = dist_step<SPF_METRIC_EUCLIDEAN>(acc, qv.x, v.x);
      acc = dist_step<SPF_METRIC_EUCLIDEAN>(acc, qv.y, v.y);
      acc = dist_step<SPF_METRIC_EUCLIDEAN>(acc, qv.z, v.z);
      acc = dist_step<SPF_METRIC_EUCLIDEAN>(acc, qv.w, v.w);
    }
    const uint32_t pos = g * 32 + lane;
    const bool valid = pos < a.lens[l];
    const unsigned long long ck =
        ((unsigned long long)__float_as_uint(acc) << 32) | (unsigned long long)(seqb[p] + pos);
    unsigned bal = __ballot_sync(0xffffffffu, valid && acc <= thr && ck < kth && (!later || ck > after));
    while (bal) {
      const int src = __ffs(bal) - 1;
      bal &= bal - 1;
      const unsigned long long k2 = __shfl_sync(0xffffffffu, ck, src);
      if (k2 < kth) {
        topk_insert<R>(key, pay, k2, G * 32 + src, lane);
        kth = topk_kth<R>(key, a.K);
      }
    }
  }
  // merge the per-warp lists: warp 0 inserts the others' entries
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const uint32_t e = r * 32 + lane;
    if (e < a.K) { s_keys[warp * a.K + e] = key[r]; s_pay[warp * a.K + e] = pay[r]; }
  }
  __syncthreads();
  if (warp != 0) return;
  for (int w = 1; w < SCAN_WARPS; ++w) {
    for (uint32_t e = 0; e < a.K; ++e) {
      const unsigned long long k2 = s_keys[w * a.K + e];
      if (k2 >= kth) break;           // lists are ascending
      topk_insert<R>(key, pay, k2, s_pay[w * a.K + e], lane);
      kth = topk_kth<R>(key, a.K);
    }
  }
  uint32_t count = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const uint32_t e = r * 32 + lane;
    const bool ok = e < a.K && key[r] != ~0ull;
    count += __popc(__ballot_sync(0xffffffffu, ok));
    if (e < a.K) {
      const size_t o = q * ostride + a.out_col + e;
      a.out_ids[o] = ok ? a.slot_ids[pay[r]] : ~0ull;
      a.out_dists[o] = ok ? __uint_as_float((uint32_t)(key[r] >> 32)) : __int_as_float(0x7f800000);
      a.out_keys[o] = key[r];
      a.out_slots[o] = ok ? pay[r] : ~0ull;
    }
  }
  if (lane == 0) a.out_counts[q] = (later ? a.out_counts[q] : 0u) + count;
}

template <int R>
__global__ void __launch_bounds__(SCAN_WARPS * 32)
scan_kernel(ScanArgs a) {
  scan_query<R>(a, blockIdx.x);
}

// The exact fallback behind the tensor scan: only the queries with only[q] != 0 run, and usually none
// is flagged.  One CTA looks at 256 flags and leaves at once when none is set (a 100 k-query batch is
// 391 CTAs instead of 100 000 that each read one flag: 0.065 -> ~0.005 ms), otherwise it takes its
// flagged queries one after the other.
template <int R>
__global__ void __launch_bounds__(SCAN_WARPS * 32)
scan_flagged_kernel(ScanArgs a, uint64_t nq) {
  const uint64_t q0 = (uint64_t)blockIdx.x * (SCAN_WARPS * 32);
  const uint64_t mine = q0 + threadIdx.x;
  const int flag = mine < nq && a.only[mine] != 0;
  if (!__syncthreads_or(flag)) return;
  for (uint32_t i = 0; i < (uint32_t)(SCAN_WARPS * 32) && q0 + i < nq; ++i) {
    if (a.only[q0 + i] == 0) continue;                    // uniform over the CTA
    scan_query<R>(a, q0 + i);
    __syncthreads();                                      // shared memory is reused by the next query
  }
}

// ---- list-major scan ---------------------------------------------------------------------------
// With many queries in flight most posting lists are probed by several of them.  Inverting the
// probe table (sort of (list, query-probe) pairs) lets one CTA own one list and evaluate it
// against batches of QB queries at a time: every vector chunk that is loaded is used QB times,
// so the scan stops being bound by L2/HBM traffic and runs at the FP32 issue rate of the exact
// (un-fused, sequential) distance.  Results are identical to the query-major kernel: the same
// sequential f32 sums, the same `<= thr` filter, the same (distance, encounter index) keys.
constexpr int LS_WARPS = 8;
constexpr int LS_QB = 8;              // queries a warp evaluates per pass over the list

struct ListScanArgs {
  ScanArgs s;
  const uint32_t* pair_sorted;        // (q * nprobe + p), grouped by probed list
  const uint32_t* list_off;           // nlists + 1 offsets into pair_sorted
  const uint32_t* unit_off;           // nlists + 1: first unit (= batch of LS_QB queries) of each list
  unsigned long long* unit_keys;      // (nq * nprobe) x K per-(query, probe) partial top-k
  unsigned long long* unit_slots;
};

// units per list = ceil(queries probing it / LS_QB), 0 for lists this rank does not hold
__global__ void unit_counts_kernel(const uint32_t* __restrict__ list_off, const uint64_t* __restrict__ grp_off,
                                   uint32_t nlists, uint32_t* __restrict__ counts) {
  const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l > nlists) return;
  uint32_t n = 0;
  if (l < nlists && grp_off[l + 1] != grp_off[l]) n = (list_off[l + 1] - list_off[l] + LS_QB - 1) / LS_QB;
  counts[l] = n;
}

// Sorted insertion of one key into a warp-distributed ascending list of 32 keys (lane = rank).
__device__ __forceinline__ void topk_insert_key(unsigned long long& key, unsigned long long ck, int lane) {
  const int pos = __popc(__ballot_sync(0xffffffffu, key < ck));
  const unsigned long long uk = __shfl_up_sync(0xffffffffu, key, 1);
  if (lane > pos) key = uk;
  else if (lane == pos) key = ck;
}

// One CTA per unit = (list, batch of up to LS_QB queries probing it).  The batch's query vectors sit
// in shared memory; the warps split the list's 32-vector groups round-robin, a lane computes the
// exact squared L2 of its slot's vector to all LS_QB queries (spann_index.rs:172), and each warp
// keeps the K smallest (distance, encounter index) keys per query (the slot of a result follows
// from its encounter index, so no payload is carried); the warps' partial lists are merged at
// the end, one query per warp.  The vector chunks are double-buffered in registers: the loads of
// the next four chunks are in flight while the current four are consumed.
//
// WARP_UNIT = true is the variant for short lists (a few groups each): every warp is a unit of its
// own and walks all groups of its list, so there is no cross-warp merge and no block barrier.
template <bool WARP_UNIT>
__global__ void __launch_bounds__(LS_WARPS * 32, 2)
scan_lists_kernel(ListScanArgs la, uint32_t nlists) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const ScanArgs& a = la.s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // unit -> (list, batch): last list whose first unit is <= unit
  const uint32_t unit = WARP_UNIT ? blockIdx.x * LS_WARPS + warp : blockIdx.x;
  uint32_t lo = 0, hi = nlists;
  if (unit >= la.unit_off[nlists]) return;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (la.unit_off[mid] <= unit) lo = mid; else hi = mid;
  }
  const uint32_t l = lo;
  const uint32_t batch = unit - la.unit_off[l];
  const uint32_t ng = (uint32_t)(a.grp_off[l + 1] - a.grp_off[l]);
  const uint32_t p0 = la.list_off[l] + batch * LS_QB;
  const uint32_t nb = min((uint32_t)LS_QB, la.list_off[l + 1] - p0);
  const uint32_t ld4 = a.ld / 4;
  // query vectors: LS_QB x ld4 per unit (per CTA, or per warp in the WARP_UNIT variant)
  float4* s_q = reinterpret_cast<float4*>(smem_raw) + (WARP_UNIT ? (size_t)warp * LS_QB * ld4 : 0);
  unsigned long long* s_keys =
      reinterpret_cast<unsigned long long*>(reinterpret_cast<float4*>(smem_raw) + (size_t)LS_QB * ld4);   // [warp][qi][K]
  const uint64_t G0 = a.grp_off[l];
  const uint32_t len = a.lens[l];
  const float4* V4 = reinterpret_cast<const float4*>(a.vecs);
  if (WARP_UNIT ? lane == 0 : threadIdx.x == 0) atomicAdd(a.bytes, (unsigned long long)nb * len * a.d * 4ull);
  const uint32_t tid = WARP_UNIT ? (uint32_t)lane : threadIdx.x, nthr = WARP_UNIT ? 32u : blockDim.x;

  float thr[LS_QB];
  uint32_t seq[LS_QB];
  unsigned long long key[LS_QB], kth[LS_QB];
#pragma unroll
  for (int qi = 0; qi < LS_QB; ++qi) {
    key[qi] = ~0ull; kth[qi] = ~0ull;
    thr[qi] = -1.0f; seq[qi] = 0;
    const float4* src = nullptr;
    if ((uint32_t)qi < nb) {
      const uint32_t pair = la.pair_sorted[p0 + qi];
      const uint32_t q = pair / a.nprobe;
      thr[qi] = a.thr[q];
      seq[qi] = a.seqbase[pair];
      src = reinterpret_cast<const float4*>(a.Q + (size_t)q * a.ld);
    }
    for (uint32_t c = tid; c < ld4; c += nthr)
      s_q[qi * ld4 + c] = src ? src[c] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (WARP_UNIT) __syncwarp(); else __syncthreads();
  for (uint32_t g = WARP_UNIT ? 0u : (uint32_t)warp; g < ng; g += WARP_UNIT ? 1u : (uint32_t)LS_WARPS) {
    const float4* base = V4 + (G0 + g) * ld4 * 32 + lane;
    float acc[LS_QB];
#pragma unroll
    for (int qi = 0; qi < LS_QB; ++qi) acc[qi] = 0.0f;
    float4 va[4], vb[4];
    auto load4 = [&](float4 (&v)[4], uint32_t c0) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = (c0 + u < ld4) ? __ldg(base + (size_t)(c0 + u) * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto use4 = [&](const float4 (&v)[4], uint32_t c0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (c0 + u < ld4) {
#pragma unroll
          for (int qi = 0; qi < LS_QB; ++qi) {
            const float4 qv = s_q[qi * ld4 + c0 + u];
            acc[qi] = dist_step<SPF_METRIC_EUCLIDEAN>(acc[qi], qv.x, v[u].x);
            acc[qi] = dist_step<SPF_METRIC_EUCLIDEAN>(acc[qi], qv.y, v[u].y);
            acc[qi] = dist_step<SPF_METRIC_EUCLIDEAN>(acc[qi], qv.z, v[u].z);
            acc[qi] = dist_step<SPF_METRIC_EUCLIDEAN>(acc[qi], qv.w, v[u].w);
          }
        }
      }
    };
    load4(va, 0);
    for (uint32_t c0 = 0; c0 < ld4; c0 += 8) {
      load4(vb, c0 + 4);
      use4(va, c0);
      load4(va, c0 + 8);
      use4(vb, c0 + 4);
    }
    const uint32_t pos = g * 32 + lane;
    const bool valid = pos < len;
#pragma unroll
    for (int qi = 0; qi < LS_QB; ++qi) {
      const unsigned long long ck =
          ((unsigned long long)__float_as_uint(acc[qi]) << 32) | (unsigned long long)(seq[qi] + pos);
      unsigned bal = __ballot_sync(0xffffffffu, valid && acc[qi] <= thr[qi] && ck < kth[qi]);
      while (bal) {
        const int src = __ffs(bal) - 1;
        bal &= bal - 1;
        const unsigned long long k2 = __shfl_sync(0xffffffffu, ck, src);
        if (k2 < kth[qi]) {
          topk_insert_key(key[qi], k2, lane);
          kth[qi] = __shfl_sync(0xffffffffu, key[qi], (a.K - 1) & 31);
        }
      }
    }
  }
  if (WARP_UNIT) {          // this warp saw the whole list: its lists are the unit results
#pragma unroll
    for (int qi = 0; qi < LS_QB; ++qi) {
      if ((uint32_t)qi < nb && (uint32_t)lane < a.K) {
        const uint32_t pair = la.pair_sorted[p0 + qi];
        la.unit_keys[(size_t)pair * a.K + lane] = key[qi];
        la.unit_slots[(size_t)pair * a.K + lane] =
            key[qi] == ~0ull ? ~0ull : G0 * 32 + ((uint32_t)(key[qi] & 0xffffffffull) - seq[qi]);
      }
    }
    return;
  }
  // merge the warps' partial lists: warp w owns query w of the batch
#pragma unroll
  for (int qi = 0; qi < LS_QB; ++qi)
    if ((uint32_t)lane < a.K) s_keys[((size_t)warp * LS_QB + qi) * a.K + lane] = key[qi];
  __syncthreads();
  static_assert(LS_WARPS == LS_QB, "one query per warp in the final merge");
  if ((uint32_t)warp >= nb) return;
  unsigned long long mk = ~0ull, mkth = ~0ull;
  for (int w = 0; w < LS_WARPS; ++w) {
    const unsigned long long* sk = s_keys + ((size_t)w * LS_QB + warp) * a.K;
    for (uint32_t e = 0; e < a.K; ++e) {
      const unsigned long long k2 = sk[e];
      if (k2 >= mkth) break;          // partial lists are ascending
      topk_insert_key(mk, k2, lane);
      mkth = __shfl_sync(0xffffffffu, mk, (a.K - 1) & 31);
    }
  }
  const uint32_t mypair = la.pair_sorted[p0 + warp];
  if ((uint32_t)lane < a.K) {
    la.unit_keys[(size_t)mypair * a.K + lane] = mk;
    // slot of a result: first slot of the list + (encounter index - encounter base of this probe)
    la.unit_slots[(size_t)mypair * a.K + lane] =
        mk == ~0ull ? ~0ull : G0 * 32 + ((uint32_t)(mk & 0xffffffffull) - a.seqbase[mypair]);
  }
}

// Per query: merge the (ascending) partial top-k of its nprobe probed lists.  One warp per query.
__global__ void __launch_bounds__(256)
merge_units_kernel(ListScanArgs la, uint64_t nq) {
  const ScanArgs& a = la.s;
  const int lane = threadIdx.x & 31;
  const uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  unsigned long long key[1] = {~0ull}, pay[1] = {~0ull};
  unsigned long long kth = ~0ull;
  for (uint32_t p = 0; p < a.nprobe; ++p) {
    const uint32_t lst = a.probe[q * a.nprobe + p];
    if (a.grp_off[lst + 1] == a.grp_off[lst]) continue;          // other rank's list (or empty): no unit ran
    const unsigned long long* uk = la.unit_keys + ((size_t)q * a.nprobe + p) * a.K;
    const unsigned long long* us = la.unit_slots + ((size_t)q * a.nprobe + p) * a.K;
    for (uint32_t e = 0; e < a.K; ++e) {
      const unsigned long long k2 = uk[e];
      if (k2 >= kth) break;           // unit lists are ascending
      topk_insert<1>(key, pay, k2, us[e], lane);
      kth = __shfl_sync(0xffffffffu, key[0], (a.K - 1) & 31);
    }
  }
  const bool ok = (uint32_t)lane < a.K && key[0] != ~0ull;
  const uint32_t count = __popc(__ballot_sync(0xffffffffu, ok));
  if ((uint32_t)lane < a.K) {
    a.out_ids[q * a.K + lane] = ok ? a.slot_ids[pay[0]] : ~0ull;
    a.out_dists[q * a.K + lane] = ok ? __uint_as_float((uint32_t)(key[0] >> 32)) : __int_as_float(0x7f800000);
    a.out_keys[q * a.K + lane] = key[0];
    a.out_slots[q * a.K + lane] = ok ? pay[0] : ~0ull;
  }
  if (lane == 0) a.out_counts[q] = count;
}

// Row-major rows → one posting list in slot layout (ids = row numbers, pad slots zero / UINT64_MAX).
__global__ void rows_to_slots_kernel(const float* __restrict__ X, uint32_t ld4, uint32_t nrows, uint32_t nslots,
                                     float* __restrict__ vecs, uint64_t* __restrict__ slot_ids) {
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= (uint64_t)nslots * ld4) return;
  const uint32_t pos = (uint32_t)(w / ld4), c = (uint32_t)(w - (uint64_t)pos * ld4);
  const float4 v = pos < nrows ? __ldg(reinterpret_cast<const float4*>(X) + (size_t)pos * ld4 + c)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
  reinterpret_cast<float4*>(vecs)[((size_t)(pos >> 5) * ld4 + c) * 32 + (pos & 31)] = v;
  if (c == 0) slot_ids[pos] = pos < nrows ? (uint64_t)pos : ~0ull;
}

// Probe table from the sorted (distance, list id) keys of the tensor-core probe: the probed lists,
// the prune threshold (:165) and the encounter-index base of every probe.  A query with fewer than
// nprobe keys (non-finite distances) raises `redo`: the exact probe kernels own such batches.
__global__ void probe_from_keys_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ counts,
                                       uint64_t nq, uint32_t nprobe, float prune_factor,
                                       const uint32_t* __restrict__ lens, uint32_t* __restrict__ probe,
                                       float* __restrict__ thr, uint32_t* __restrict__ seqbase, int* __restrict__ redo) {
  const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  if (counts[q] < nprobe) { *redo = 1; return; }
  uint32_t seq = 0;
  for (uint32_t p = 0; p < nprobe; ++p) {
    const unsigned long long k = keys[q * nprobe + p];
    const uint32_t j = (uint32_t)(k & 0xffffffffull);
    probe[q * nprobe + p] = j;
    seqbase[q * nprobe + p] = seq;
    seq += lens[j];
    if (p == 0) thr[q] = __fmul_rn(prune_factor, __fadd_rn(__uint_as_float((uint32_t)(k >> 32)), 1.1920929e-7f));
  }
}

__global__ void iota_u32_kernel(uint32_t* p, uint64_t n) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) p[t] = (uint32_t)t;
}

__global__ void list_offsets_kernel(const uint32_t* __restrict__ keys_sorted, uint64_t total, uint32_t nlists,
                                    uint32_t* __restrict__ offsets) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > nlists) return;
  uint64_t lo = 0, hi = total;   // first position with key >= c
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (keys_sorted[mid] < c) lo = mid + 1; else hi = mid;
  }
  offsets[c] = (uint32_t)lo;
}

__global__ void gather_vectors_kernel(const float* __restrict__ vecs, uint32_t ld, uint32_t d,
                                      const unsigned long long* __restrict__ slots, uint64_t nres,
                                      float* __restrict__ out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nres * d) return;
  const uint64_t r = t / d;
  const uint32_t i = (uint32_t)(t - r * d);
  const unsigned long long s = slots[r];
  float v = 0.0f;
  if (s != ~0ull) {
    const uint64_t G = s >> 5;
    const uint32_t lane = (uint32_t)(s & 31);
    v = vecs[((G * (ld / 4) + (i >> 2)) * 32 + lane) * 4 + (i & 3)];
  }
  out[t] = v;
}

int make_dir(const char* dir) {
  if (mkdir(dir, 0777) != 0 && errno != EEXIST) return -1;
  return 0;
}

void put_u64(std::vector<unsigned char>& b, uint64_t v) {
  for (int i = 0; i < 8; ++i) b.push_back((unsigned char)(v >> (8 * i)));
}

bool get_u64(const std::vector<unsigned char>& b, size_t& o, uint64_t* v) {
  if (o + 8 > b.size()) return false;
  uint64_t r = 0;
  for (int i = 0; i < 8; ++i) r |= (uint64_t)b[o + i] << (8 * i);
  o += 8;
  *v = r;
  return true;
}

bool read_file(const std::string& path, std::vector<unsigned char>& out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize(sz > 0 ? (size_t)sz : 0);
  const bool ok = sz <= 0 || fread(out.data(), 1, (size_t)sz, f) == (size_t)sz;
  fclose(f);
  return ok;
}

// uploads host slot arrays and finishes the index object
int index_finish(spf_index* idx, const std::vector<float>* h_vecs, const std::vector<uint64_t>* h_ids) {
  spf_ctx* c = idx->ctx;
  cudaStream_t st = c->stream;
  SPF_CUDA(cudaMalloc((void**)&idx->grp_off, ((size_t)idx->nlists + 1) * sizeof(uint64_t)));
  SPF_CUDA(cudaMalloc((void**)&idx->lens, (size_t)(idx->nlists ? idx->nlists : 1) * sizeof(uint32_t)));
  SPF_CUDA(cudaMemcpyAsync(idx->grp_off, idx->h_grp_off.data(), ((size_t)idx->nlists + 1) * sizeof(uint64_t),
                           cudaMemcpyHostToDevice, st));
  SPF_CUDA(cudaMemcpyAsync(idx->lens, idx->h_lens.data(), (size_t)idx->nlists * sizeof(uint32_t),
                           cudaMemcpyHostToDevice, st));
  if (h_vecs) {
    const size_t nslots = (size_t)idx->total_groups * 32;
    SPF_CUDA(cudaMalloc((void**)&idx->vecs, (nslots ? nslots : 1) * idx->ld * sizeof(float)));
    SPF_CUDA(cudaMalloc((void**)&idx->slot_ids, (nslots ? nslots : 1) * sizeof(uint64_t)));
    SPF_CUDA(cudaMemcpyAsync(idx->vecs, h_vecs->data(), nslots * idx->ld * sizeof(float), cudaMemcpyHostToDevice, st));
    SPF_CUDA(cudaMemcpyAsync(idx->slot_ids, h_ids->data(), nslots * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  }
  SPF_CUDA(cudaStreamSynchronize(st));
  return SPF_OK;
}

template <int R>
int launch_scan(spf_ctx* c, const ScanArgs& a, uint64_t nq) {
  const size_t smem = (size_t)a.ld * 4 + ((size_t)a.nprobe + 1) * 4 + 8 + (size_t)SCAN_WARPS * a.K * 16;
  if (smem > 48 * 1024)
    SPF_CUDA(cudaFuncSetAttribute(scan_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  scan_kernel<R><<<(unsigned)nq, SCAN_WARPS * 32, smem, c->stream>>>(a);
  return check_launch(c, "scan_kernel");
}

// launch of the flagged-queries form (a.only must be set)
template <int R>
int launch_scan_flagged(spf_ctx* c, const ScanArgs& a, uint64_t nq) {
  const size_t smem = (size_t)a.ld * 4 + ((size_t)a.nprobe + 1) * 4 + 8 + (size_t)SCAN_WARPS * a.K * 16;
  if (smem > 48 * 1024)
    SPF_CUDA(cudaFuncSetAttribute(scan_flagged_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  scan_flagged_kernel<R><<<(unsigned)ceil_div(nq, (uint64_t)(SCAN_WARPS * 32)), SCAN_WARPS * 32, smem, c->stream>>>(a, nq);
  return check_launch(c, "scan_flagged_kernel");
}

}  // namespace
}  // namespace spf

extern "C" {

int spf_index_pack(spf_dataset* ds, const uint64_t* offsets, const uint64_t* members,
                   const uint64_t* centroid_rows, uint32_t nlists, uint32_t list_begin, uint32_t list_end,
                   spf_index** out) {
  return spf::guarded([&]() -> int {
  if (!ds || !offsets || !centroid_rows || !out) return fail(SPF_E_INVALID, "spf_index_pack: NULL argument");
  *out = nullptr;
  if (nlists == 0) return fail(SPF_E_INVALID, "nlists must be > 0");
  if (list_begin > list_end || list_end > nlists) return fail(SPF_E_INVALID, "bad list range [%u,%u)", list_begin, list_end);
  if (offsets[nlists] && !members) return fail(SPF_E_INVALID, "members is NULL");
  for (uint32_t l = 0; l < nlists; ++l) {
    if (offsets[l] > offsets[l + 1]) return fail(SPF_E_INVALID, "offsets must be non-decreasing");
    if (offsets[l + 1] - offsets[l] >= (1ull << 32)) return fail(SPF_E_INVALID, "list %u too long", l);
    if (centroid_rows[l] >= ds->n) return fail(SPF_E_INVALID, "centroid row %llu >= n", (unsigned long long)centroid_rows[l]);
  }
  for (uint64_t t = offsets[list_begin]; t < offsets[list_end]; ++t)
    if (members[t] >= ds->n) return fail(SPF_E_INVALID, "member row %llu >= n", (unsigned long long)members[t]);
  spf_ctx* c = ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  spf_index* idx = new (std::nothrow) spf_index();
  if (!idx) return fail(SPF_E_OOM, "out of host memory");
  idx->ctx = c; idx->d = ds->d; idx->ld = ds->ld; idx->nlists = nlists;
  idx->list_begin = list_begin; idx->list_end = list_end;
  idx->h_grp_off.assign((size_t)nlists + 1, 0);
  idx->h_lens.resize(nlists);
  uint64_t g = 0;
  for (uint32_t l = 0; l < nlists; ++l) {
    const uint64_t len = offsets[l + 1] - offsets[l];
    idx->h_lens[l] = (uint32_t)len;
    idx->h_grp_off[l] = g;
    if (l >= list_begin && l < list_end) { g += (len + 31) / 32; idx->total_vectors += len; }
  }
  idx->h_grp_off[nlists] = g;
  idx->total_groups = g;

  auto cleanup = [&](int rc) { spf_index_free(idx); return rc; };
  const uint32_t nloc = list_end - list_begin;
  const uint64_t nmem = offsets[list_end] - offsets[list_begin];
  const size_t nslots = (size_t)g * 32;
  // centroids (all lists)
  DevBuf<uint64_t> d_crow, d_rows, d_loc_off, d_grp_loc;
  int rc = d_crow.alloc(st, nlists);
  if (rc < 0) return cleanup(rc);
  if (cudaMalloc((void**)&idx->centroids, (size_t)nlists * idx->ld * sizeof(float)) != cudaSuccess)
    return cleanup(fail(SPF_E_OOM, "centroid allocation failed"));
  cudaMemcpyAsync(d_crow.p, centroid_rows, (size_t)nlists * sizeof(uint64_t), cudaMemcpyHostToDevice, st);
  rc = launch_gather_rows(c, ds->x, ds->ld, d_crow.p, nlists, idx->centroids);
  if (rc < 0) return cleanup(rc);
  // local lists
  if (cudaMalloc((void**)&idx->vecs, (nslots ? nslots : 1) * idx->ld * sizeof(float)) != cudaSuccess ||
      cudaMalloc((void**)&idx->slot_ids, (nslots ? nslots : 1) * sizeof(uint64_t)) != cudaSuccess)
    return cleanup(fail(SPF_E_OOM, "posting-list allocation of %zu slots failed", nslots));
  cudaMemsetAsync(idx->vecs, 0, (nslots ? nslots : 1) * idx->ld * sizeof(float), st);
  cudaMemsetAsync(idx->slot_ids, 0xff, (nslots ? nslots : 1) * sizeof(uint64_t), st);
  if (nloc && nmem) {
    std::vector<uint64_t> loc_off(nloc + 1), grp_loc(nloc + 1);
    for (uint32_t l = 0; l <= nloc; ++l) {
      loc_off[l] = offsets[list_begin + l] - offsets[list_begin];
      grp_loc[l] = idx->h_grp_off[list_begin + l];
    }
    if ((rc = d_rows.alloc(st, nmem)) < 0 || (rc = d_loc_off.alloc(st, nloc + 1)) < 0 ||
        (rc = d_grp_loc.alloc(st, nloc + 1)) < 0)
      return cleanup(rc);
    cudaMemcpyAsync(d_rows.p, members + offsets[list_begin], nmem * sizeof(uint64_t), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_loc_off.p, loc_off.data(), (nloc + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_grp_loc.p, grp_loc.data(), (nloc + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st);
    pack_lists_kernel<<<nloc, 256, 0, st>>>(ds->x, idx->ld / 4, d_rows.p, d_loc_off.p, d_grp_loc.p, idx->vecs, idx->slot_ids);
    rc = check_launch(c, "pack_lists_kernel");
    if (rc < 0) return cleanup(rc);
    if (cudaStreamSynchronize(st) != cudaSuccess) return cleanup(fail(SPF_E_CUDA, "index pack failed on the device"));
  }
  rc = index_finish(idx, nullptr, nullptr);
  if (rc < 0) return cleanup(rc);
  *out = idx;
  return SPF_OK;
  });
}

int spf_index_load_dir(spf_ctx* c, const char* dir, const float* centroids, uint32_t nlists, uint32_t d,
                       spf_index** out) {
  return spf::guarded([&]() -> int {
  if (!c || !dir || !centroids || !out) return fail(SPF_E_INVALID, "spf_index_load_dir: NULL argument");
  *out = nullptr;
  if (nlists == 0 || d == 0) return fail(SPF_E_INVALID, "nlists and d must be > 0");
  // cluster_ids.bin: u64 m, m x u64 (posting_lists.rs:47-59,108-113); order is arbitrary
  std::vector<unsigned char> buf;
  const std::string base(dir);
  if (!read_file(base + "/cluster_ids.bin", buf)) return fail(SPF_E_IO, "cannot read %s/cluster_ids.bin", dir);
  size_t o = 0;
  uint64_t m = 0;
  if (!get_u64(buf, o, &m)) return fail(SPF_E_IO, "cluster_ids.bin is truncated");
  std::vector<char> present(nlists, 0);
  for (uint64_t i = 0; i < m; ++i) {
    uint64_t id;
    if (!get_u64(buf, o, &id)) return fail(SPF_E_IO, "cluster_ids.bin is truncated");
    if (id >= nlists) return fail(SPF_E_IO, "cluster id %llu >= nlists %u", (unsigned long long)id, nlists);
    present[id] = 1;
  }
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  spf_index* idx = new (std::nothrow) spf_index();
  if (!idx) return fail(SPF_E_OOM, "out of host memory");
  idx->ctx = c; idx->d = d; idx->ld = round_up(d, 4); idx->nlists = nlists;
  idx->list_begin = 0; idx->list_end = nlists;
  idx->h_grp_off.assign((size_t)nlists + 1, 0);
  idx->h_lens.assign(nlists, 0);
  auto cleanup = [&](int rc) { spf_index_free(idx); return rc; };
  const uint32_t ld = idx->ld, ld4 = ld / 4;
  std::vector<float> h_vecs;
  std::vector<uint64_t> h_ids;
  uint64_t g = 0;
  for (uint32_t l = 0; l < nlists; ++l) {
    idx->h_grp_off[l] = g;
    if (!present[l]) continue;        // get_posting_list → Ok(None) (posting_lists.rs:99-101)
    char name[64];
    snprintf(name, sizeof(name), "/posting_list_%u.bin", l);
    if (!read_file(base + name, buf)) return cleanup(fail(SPF_E_IO, "cannot read %s%s", dir, name));
    o = 0;
    uint64_t len = 0;
    if (!get_u64(buf, o, &len)) return cleanup(fail(SPF_E_IO, "%s is truncated", name));
    if (len >= (1ull << 32)) return cleanup(fail(SPF_E_IO, "%s: list too long", name));
    // the count comes from the file: check it against the file size before anything is sized by it
    if (len > (buf.size() - o) / (16ull + 4ull * d))
      return cleanup(fail(SPF_E_IO, "%s is truncated or corrupt (%llu vectors do not fit %zu bytes)", name,
                          (unsigned long long)len, buf.size()));
    const uint64_t ng = (len + 31) / 32;
    h_vecs.resize((size_t)(g + ng) * 32 * ld, 0.0f);
    h_ids.resize((size_t)(g + ng) * 32, ~0ull);
    for (uint64_t pos = 0; pos < len; ++pos) {
      uint64_t id, dd;
      if (!get_u64(buf, o, &id) || !get_u64(buf, o, &dd)) return cleanup(fail(SPF_E_IO, "%s is truncated", name));
      if (dd != d) return cleanup(fail(SPF_E_IO, "%s: vector length %llu != d %u", name, (unsigned long long)dd, d));
      if (o + 4ull * d > buf.size()) return cleanup(fail(SPF_E_IO, "%s is truncated", name));
      const uint64_t G = g + (pos >> 5);
      const uint32_t lane = (uint32_t)(pos & 31);
      for (uint32_t i = 0; i < d; ++i) {
        uint32_t bits = (uint32_t)buf[o] | ((uint32_t)buf[o + 1] << 8) | ((uint32_t)buf[o + 2] << 16) | ((uint32_t)buf[o + 3] << 24);
        o += 4;
        float v;
        memcpy(&v, &bits, 4);
        h_vecs[((G * ld4 + (i >> 2)) * 32 + lane) * 4 + (i & 3)] = v;
      }
      h_ids[G * 32 + lane] = id;
    }
    idx->h_lens[l] = (uint32_t)len;
    idx->total_vectors += len;
    g += ng;
  }
  idx->h_grp_off[nlists] = g;
  idx->total_groups = g;
  // centroids
  if (cudaMalloc((void**)&idx->centroids, (size_t)nlists * ld * sizeof(float)) != cudaSuccess)
    return cleanup(fail(SPF_E_OOM, "centroid allocation failed"));
  cudaMemsetAsync(idx->centroids, 0, (size_t)nlists * ld * sizeof(float), c->stream);
  cudaMemcpy2DAsync(idx->centroids, (size_t)ld * 4, centroids, (size_t)d * 4, (size_t)d * 4, nlists,
                    cudaMemcpyHostToDevice, c->stream);
  int rc = index_finish(idx, &h_vecs, &h_ids);
  if (rc < 0) return cleanup(rc);
  *out = idx;
  return SPF_OK;
  });
}

int spf_index_save_dir(const spf_index* idx, const char* dir) {
  return spf::guarded([&]() -> int {
  if (!idx || !dir) return fail(SPF_E_INVALID, "spf_index_save_dir: NULL argument");
  spf_ctx* c = idx->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  if (make_dir(dir) != 0) return fail(SPF_E_IO, "cannot create directory %s", dir);
  const size_t nslots = (size_t)idx->total_groups * 32;
  const uint32_t ld = idx->ld, ld4 = ld / 4, d = idx->d;
  std::vector<float> h_vecs(nslots * ld);
  std::vector<uint64_t> h_ids(nslots);
  if (nslots) {
    SPF_CUDA(cudaMemcpyAsync(h_vecs.data(), idx->vecs, nslots * ld * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    SPF_CUDA(cudaMemcpyAsync(h_ids.data(), idx->slot_ids, nslots * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    SPF_CUDA(cudaStreamSynchronize(c->stream));
  }
  const std::string base(dir);
  std::vector<unsigned char> buf;
  std::vector<uint64_t> ids_written;
  for (uint32_t l = idx->list_begin; l < idx->list_end; ++l) {
    // posting_list_{id}.bin = bincode(Vec<PointData>) (posting_lists.rs:71-90)
    buf.clear();
    const uint64_t len = idx->h_lens[l];
    put_u64(buf, len);
    for (uint64_t pos = 0; pos < len; ++pos) {
      const uint64_t G = idx->h_grp_off[l] + (pos >> 5);
      const uint32_t lane = (uint32_t)(pos & 31);
      put_u64(buf, h_ids[G * 32 + lane]);
      put_u64(buf, d);
      for (uint32_t i = 0; i < d; ++i) {
        uint32_t bits;
        const float v = h_vecs[((G * ld4 + (i >> 2)) * 32 + lane) * 4 + (i & 3)];
        memcpy(&bits, &v, 4);
        for (int b = 0; b < 4; ++b) buf.push_back((unsigned char)(bits >> (8 * b)));
      }
    }
    char name[64];
    snprintf(name, sizeof(name), "/posting_list_%u.bin", l);
    FILE* f = fopen((base + name).c_str(), "wb");
    if (!f || fwrite(buf.data(), 1, buf.size(), f) != buf.size()) {
      if (f) fclose(f);
      return fail(SPF_E_IO, "cannot write %s%s", dir, name);
    }
    fclose(f);
    ids_written.push_back(l);
  }
  // cluster_ids.bin lists every list of the directory.  A list-sharded index holds only the range
  // [list_begin, list_end): ids already recorded there by another shard are kept (read, merge,
  // write to a temporary, rename), so shards saved one after the other build up the complete
  // directory.  Concurrent saves into one directory must be serialised by the caller.
  std::vector<uint64_t> all_ids;
  {
    std::vector<unsigned char> old;
    if (read_file(base + "/cluster_ids.bin", old)) {
      size_t o = 0;
      uint64_t m = 0;
      if (get_u64(old, o, &m) && m <= (old.size() - o) / 8) {
        for (uint64_t i = 0; i < m; ++i) {
          uint64_t id = 0;
          get_u64(old, o, &id);
          if (id < idx->list_begin || id >= idx->list_end) all_ids.push_back(id);
        }
      }
    }
  }
  all_ids.insert(all_ids.end(), ids_written.begin(), ids_written.end());
  std::sort(all_ids.begin(), all_ids.end());
  all_ids.erase(std::unique(all_ids.begin(), all_ids.end()), all_ids.end());
  buf.clear();
  put_u64(buf, all_ids.size());
  for (uint64_t id : all_ids) put_u64(buf, id);
  const std::string tmp = base + "/cluster_ids.bin.tmp", fin = base + "/cluster_ids.bin";
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f || fwrite(buf.data(), 1, buf.size(), f) != buf.size()) {
    if (f) fclose(f);
    return fail(SPF_E_IO, "cannot write %s/cluster_ids.bin", dir);
  }
  fclose(f);
  if (rename(tmp.c_str(), fin.c_str()) != 0) return fail(SPF_E_IO, "cannot replace %s/cluster_ids.bin", dir);
  return SPF_OK;
  });
}

void spf_index_free(spf_index* idx) {
  if (!idx) return;
  cudaSetDevice(idx->ctx->device);
  cudaStreamSynchronize(idx->ctx->stream);
  if (idx->centroids) cudaFree(idx->centroids);
  if (idx->vecs) cudaFree(idx->vecs);
  if (idx->slot_ids) cudaFree(idx->slot_ids);
  if (idx->grp_off) cudaFree(idx->grp_off);
  if (idx->lens) cudaFree(idx->lens);
  spf::scan_tc_release(&idx->tc);
  spf::scan_tc_release(&idx->ctc);
  if (idx->cvecs) cudaFree(idx->cvecs);
  if (idx->cids) cudaFree(idx->cids);
  if (idx->cgrp) cudaFree(idx->cgrp);
  if (idx->clens) cudaFree(idx->clens);
  delete idx;
}

uint32_t spf_index_lists(const spf_index* idx) { return idx ? idx->nlists : 0; }
uint64_t spf_index_vectors(const spf_index* idx) { return idx ? idx->total_vectors : 0; }
uint64_t spf_index_last_scan_bytes(const spf_index* idx) { return idx ? idx->last_scan_bytes : 0; }

// The two phases of a search on device buffers (shared by spf_search_batch and the list-sharded
// spf_search_sharded).  Qp: nq x ld query rows (zero padded).
static int search_probe(spf_index* idx, const float* Qp, uint64_t nq, uint32_t nprobe, float prune_factor,
                        uint32_t* probe, float* thr, uint32_t* seqbase) {
  spf_ctx* c = idx->ctx;
  cudaStream_t st = c->stream;
  const uint32_t ld = idx->ld, d = idx->d, nlists = idx->nlists;
  // centroid probe.  Large batches with nprobe <= 32: the centroids are one posting list that every
  // query probes, so the tensor-core candidate scan (scan_tc.cu) with K = nprobe yields the sorted
  // (distance, list id) keys; otherwise, and for batches with non-finite distances, dense exact
  // distances in query chunks followed by a selection kernel.
  bool probed = false;
  const uint32_t cslots = round_up(nlists, 32);
  const bool tc_probe_ok = c->params.scan_tc != 0 && scan_tc_supported(c, ld, cslots, nprobe <= 32 ? nprobe : 1, nq) &&
                           (c->params.scan_tc == 2 || (nq >= 2048 && nlists >= 512));
  auto centroid_side = [&]() -> int {      // the centroids as one posting list + TF32 side structures, once
    if (idx->ctc.ready) return SPF_OK;
    const uint64_t hg[2] = {0, cslots / 32};
    SPF_CUDA(cudaMalloc((void**)&idx->cvecs, (size_t)cslots * ld * sizeof(float)));
    SPF_CUDA(cudaMalloc((void**)&idx->cids, (size_t)cslots * sizeof(uint64_t)));
    SPF_CUDA(cudaMalloc((void**)&idx->cgrp, 2 * sizeof(uint64_t)));
    SPF_CUDA(cudaMalloc((void**)&idx->clens, sizeof(uint32_t)));
    SPF_CUDA(cudaMemcpyAsync(idx->cgrp, hg, sizeof(hg), cudaMemcpyHostToDevice, st));
    SPF_CUDA(cudaMemcpyAsync(idx->clens, &nlists, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    rows_to_slots_kernel<<<(unsigned)ceil_div((uint64_t)cslots * (ld / 4), 256), 256, 0, st>>>(
        idx->centroids, ld / 4, nlists, cslots, idx->cvecs, idx->cids);
    SPF_TRY(check_launch(c, "rows_to_slots_kernel"));
    SPF_CUDA(cudaStreamSynchronize(st));     // hg / nlists are stack variables
    return scan_tc_prepare(c, idx->cvecs, idx->cids, cslots, ld, &idx->ctc);
  };
  if (tc_probe_ok && nprobe > 32 && nlists <= 4096) {
    // dense TF32 s of every query against every centroid, selection with the certified bound
    KernelTimer t(c, "probe");
    SPF_TRY(centroid_side());
    DevBuf<int> redo;
    SPF_TRY(redo.alloc(st, 1));
    SPF_CUDA(cudaMemsetAsync(redo.p, 0, sizeof(int), st));
    SPF_TRY(probe_tc_dense(c, idx->ctc, idx->centroids, Qp, nq, ld, nlists, nprobe, prune_factor, idx->lens, probe,
                           thr, seqbase, redo.p));
    int h_redo = 0;
    SPF_CUDA(cudaMemcpyAsync(&h_redo, redo.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SPF_CUDA(cudaStreamSynchronize(st));
    probed = h_redo == 0;
  } else if (tc_probe_ok && nprobe <= 32) {
    KernelTimer t(c, "probe");
    SPF_TRY(centroid_side());
    DevBuf<uint32_t> zero, pairs, loff2, p_counts;
    DevBuf<float> thr_inf, p_dists;
    DevBuf<uint64_t> p_ids;
    DevBuf<unsigned long long> p_keys, p_slots, p_bytes;
    DevBuf<uint8_t> p_flag;
    DevBuf<int> redo;
    SPF_TRY(zero.alloc(st, nq));
    SPF_TRY(pairs.alloc(st, nq));
    SPF_TRY(loff2.alloc(st, 2));
    SPF_TRY(thr_inf.alloc(st, nq));
    SPF_TRY(p_ids.alloc(st, (size_t)nq * nprobe));
    SPF_TRY(p_dists.alloc(st, (size_t)nq * nprobe));
    SPF_TRY(p_keys.alloc(st, (size_t)nq * nprobe));
    SPF_TRY(p_slots.alloc(st, (size_t)nq * nprobe));
    SPF_TRY(p_counts.alloc(st, nq));
    SPF_TRY(p_bytes.alloc(st, 1));
    SPF_TRY(p_flag.alloc(st, nq));
    SPF_TRY(redo.alloc(st, 1));
    const uint32_t h_loff[2] = {0u, (uint32_t)nq};
    SPF_CUDA(cudaMemsetAsync(zero.p, 0, nq * sizeof(uint32_t), st));
    SPF_CUDA(cudaMemsetAsync(redo.p, 0, sizeof(int), st));
    SPF_CUDA(cudaMemcpyAsync(loff2.p, h_loff, sizeof(h_loff), cudaMemcpyHostToDevice, st));
    SPF_TRY(launch_fill_f32(c, thr_inf.p, nq, __builtin_inff()));
    iota_u32_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>(pairs.p, nq);
    SPF_TRY(check_launch(c, "iota_u32_kernel"));
    ScanTcCall pc;
    pc.s.vecs = idx->cvecs; pc.s.slot_ids = idx->cids; pc.s.grp_off = idx->cgrp; pc.s.lens = idx->clens;
    pc.s.ld = ld; pc.s.d = d; pc.s.Q = Qp; pc.s.probe = zero.p; pc.s.thr = thr_inf.p; pc.s.seqbase = zero.p;
    pc.s.nprobe = 1; pc.s.K = nprobe;
    pc.s.out_ids = p_ids.p; pc.s.out_dists = p_dists.p; pc.s.out_counts = p_counts.p; pc.s.out_keys = p_keys.p;
    pc.s.out_slots = p_slots.p; pc.s.bytes = p_bytes.p; pc.s.only = nullptr;
    pc.side = &idx->ctc; pc.pair_sorted = pairs.p; pc.list_off = loff2.p; pc.nlists = 1; pc.nq = nq;
    pc.qflag = p_flag.p; pc.is_probe = true;
    SPF_TRY(scan_tc_run(c, pc));
    ScanArgs fa = pc.s;
    fa.only = p_flag.p;
    SPF_TRY(launch_scan_flagged<1>(c, fa, nq));
    probe_from_keys_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>(p_keys.p, p_counts.p, nq, nprobe, prune_factor,
                                                                        idx->lens, probe, thr, seqbase, redo.p);
    SPF_TRY(check_launch(c, "probe_from_keys_kernel"));
    int h_redo = 0;
    SPF_CUDA(cudaMemcpyAsync(&h_redo, redo.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SPF_CUDA(cudaStreamSynchronize(st));
    probed = h_redo == 0;
  }
  if (!probed) {
    uint64_t chunk = (256ull << 20) / nlists;
    if (chunk == 0) chunk = 1;
    if (chunk > nq) chunk = nq;
    DevBuf<float> Dqc;
    SPF_TRY(Dqc.alloc(st, (size_t)chunk * nlists));
    KernelTimer t(c, "probe");
    for (uint64_t q0 = 0; q0 < nq; q0 += chunk) {
      const uint64_t nc = (nq - q0) < chunk ? (nq - q0) : chunk;
      SPF_TRY(launch_assign_exact(c, SPF_METRIC_EUCLIDEAN, Qp + q0 * ld, nc, idx->centroids, nlists, ld, 1.0f,
                                  nullptr, Dqc.p));
      if (nprobe > 16 && nlists <= 1024) {
        probe_sort_kernel<4><<<(unsigned)nc, 256, 0, st>>>(Dqc.p, nlists, nprobe, prune_factor, idx->lens,
                                                          probe + q0 * nprobe, thr + q0, seqbase + q0 * nprobe);
      } else if (nprobe > 16 && nprobe <= 256 && nlists <= 4096) {        // select the nprobe smallest, sort only those
        probe_topn_kernel<16, 1><<<(unsigned)nc, 256, 0, st>>>(Dqc.p, nlists, nprobe, prune_factor, idx->lens,
                                                              probe + q0 * nprobe, thr + q0, seqbase + q0 * nprobe);
      } else if (nprobe > 16 && nlists <= 4096) {
        probe_topn_kernel<16, 4><<<(unsigned)nc, 256, 0, st>>>(Dqc.p, nlists, nprobe, prune_factor, idx->lens,
                                                              probe + q0 * nprobe, thr + q0, seqbase + q0 * nprobe);
      } else {
        probe_select_kernel<<<(unsigned)nc, 128, 0, st>>>(Dqc.p, nlists, nprobe, prune_factor, idx->lens,
                                                          probe + q0 * nprobe, thr + q0, seqbase + q0 * nprobe);
      }
      SPF_TRY(check_launch(c, "probe_select_kernel"));
    }
  }
  return SPF_OK;
}

static int search_scan(spf_index* idx, const float* Qp, uint64_t nq, uint32_t k, uint32_t nprobe, uint32_t* probe,
                       const float* thr, const uint32_t* seqbase, uint64_t* o_ids, float* o_dists, uint32_t* o_counts,
                       unsigned long long* o_keys, unsigned long long* o_slots, unsigned long long* d_bytes,
                       const spf_comm* comm = nullptr) {
  spf_ctx* c = idx->ctx;
  cudaStream_t st = c->stream;
  const uint32_t ld = idx->ld, d = idx->d, nlists = idx->nlists;
  // list-sharded search: exactly one min-reduction of the per-query bound per scan on every rank.  The
  // tensor scan does it behind its tau pass; a rank on any other path (or without probed lists)
  // contributes +inf here at the end.
  bool bound_exchanged = false;
  struct BoundGuard {
    spf_ctx* c; const spf_comm* comm; uint64_t nq; bool* done; int rc = SPF_OK;
    int finish() {
      if (!comm || comm->world == 1 || *done) return SPF_OK;
      *done = true;
      DevBuf<float> inf;
      SPF_TRY(inf.alloc(c->stream, nq));
      SPF_CUDA(cudaMemsetAsync(inf.p, 0x7f, nq * sizeof(float), c->stream));   // 0x7f7f7f7f: a huge finite float
      return comm_allreduce_min_f32(c, comm, inf.p, nq);
    }
  } bound_guard{c, comm, nq, &bound_exchanged};
  ScanArgs a;
  a.vecs = idx->vecs; a.slot_ids = idx->slot_ids; a.grp_off = idx->grp_off; a.lens = idx->lens;
  a.ld = ld; a.d = d; a.Q = Qp; a.probe = probe; a.thr = thr; a.seqbase = seqbase;
  a.nprobe = nprobe; a.K = k;
  a.out_ids = o_ids; a.out_dists = o_dists; a.out_counts = o_counts; a.out_keys = o_keys;
  a.out_slots = o_slots; a.bytes = d_bytes; a.only = nullptr;
  // list-major scan when the batch is large enough for lists to be shared between queries; with
  // tens of probing queries per list the TF32 candidate scan (scan_tc.cu) takes over
  const uint64_t npairs = nq * nprobe;
  const uint32_t nloc = idx->list_end - idx->list_begin;
  const bool use_tc = c->params.scan_tc != 0 && scan_tc_supported(c, ld, idx->total_groups * 32, k, npairs) &&
                      (c->params.scan_tc == 2 || npairs >= 16ull * (nloc ? nloc : 1));
  // (rows too long for the list-major kernel's shared-memory tiles stay on the query-major kernel)
  const bool list_major = !use_tc && k <= 32 && npairs < (1ull << 32) && c->params.scan_list_major != 0 &&
                          (c->params.scan_list_major == 2 || npairs >= 2ull * nlists) &&
                          (size_t)LS_QB * ld * sizeof(float) + (size_t)LS_WARPS * LS_QB * k * 8 <= 200 * 1024;
  DevBuf<uint32_t> pk, pv, pk2, pv2, loff;
  DevBuf<unsigned long long> ukeys, uslots;
  DevBuf<uint8_t> stmp, qflag;
  if (use_tc) SPF_TRY(scan_tc_prepare(c, idx->vecs, idx->slot_ids, idx->total_groups * 32, ld, &idx->tc));
  if (list_major || use_tc) {
    // invert the probe table: (q * nprobe + p) grouped by probed list
    SPF_TRY(pv.alloc(st, npairs));
    SPF_TRY(pk2.alloc(st, npairs));
    SPF_TRY(pv2.alloc(st, npairs));
    SPF_TRY(loff.alloc(st, (size_t)nlists + 1));
    iota_u32_kernel<<<(unsigned)ceil_div(npairs, 256), 256, 0, st>>>(pv.p, npairs);
    SPF_TRY(check_launch(c, "iota_u32_kernel"));
    int end_bit = 1;
    while (end_bit < 32 && (1ull << end_bit) < (uint64_t)nlists) ++end_bit;
    size_t sort_bytes = 0;
    SPF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, probe, pk2.p, pv.p, pv2.p, (int64_t)npairs, 0, end_bit, st));
    SPF_TRY(stmp.alloc(st, sort_bytes));
    SPF_CUDA(cub::DeviceRadixSort::SortPairs(stmp.p, sort_bytes, probe, pk2.p, pv.p, pv2.p, (int64_t)npairs, 0, end_bit, st));
    c->launches += 3;
    list_offsets_kernel<<<(unsigned)ceil_div((uint64_t)nlists + 1, 256), 256, 0, st>>>(pk2.p, npairs, nlists, loff.p);
    SPF_TRY(check_launch(c, "list_offsets_kernel"));
  }
  if (use_tc) {
    KernelTimer t(c, "scan");
    SPF_TRY(qflag.alloc(st, nq));
    ScanTcCall tc;
    tc.s = a; tc.side = &idx->tc; tc.pair_sorted = pv2.p; tc.list_off = loff.p; tc.nlists = nlists; tc.nq = nq;
    tc.qflag = qflag.p;
    tc.comm = (comm && comm->world > 1) ? comm : nullptr;
    tc.bound_exchanged = &bound_exchanged;
    SPF_TRY(scan_tc_run(c, tc));
    SPF_TRY(bound_guard.finish());                         // no local units: the tau pass (and its reduction) did not run
    // flagged queries (no certified bound / bucket overflow): exact query-major scan
    KernelTimer t2(c, "scan_tc_fallback");
    ScanArgs fa = a;
    fa.only = qflag.p;
    SPF_TRY(launch_scan_flagged<1>(c, fa, nq));
  } else if (list_major) {
    KernelTimer t(c, "scan");
    SPF_TRY(ukeys.alloc(st, npairs * k));
    SPF_TRY(uslots.alloc(st, npairs * k));
    // units: batches of LS_QB queries per list; CTA -> unit through the scanned unit counts
    DevBuf<uint32_t> ucnt, uoff;
    SPF_TRY(ucnt.alloc(st, (size_t)nlists + 1));
    SPF_TRY(uoff.alloc(st, (size_t)nlists + 1));
    unit_counts_kernel<<<(unsigned)ceil_div((uint64_t)nlists + 1, 256), 256, 0, st>>>(loff.p, idx->grp_off, nlists, ucnt.p);
    SPF_TRY(check_launch(c, "unit_counts_kernel"));
    size_t scan_bytes = 0;
    SPF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, ucnt.p, uoff.p, (int)(nlists + 1), st));
    DevBuf<uint8_t> stmp2;
    SPF_TRY(stmp2.alloc(st, scan_bytes));
    SPF_CUDA(cub::DeviceScan::ExclusiveSum(stmp2.p, scan_bytes, ucnt.p, uoff.p, (int)(nlists + 1), st));
    c->launches += 2;
    ListScanArgs la;
    la.s = a; la.pair_sorted = pv2.p; la.list_off = loff.p; la.unit_off = uoff.p;
    la.unit_keys = ukeys.p; la.unit_slots = uslots.p;
    const uint64_t max_units = npairs / LS_QB + nlists + 1;   // upper bound; surplus units exit at once
    // short lists (a few 32-vector groups each): one warp per unit; long lists: one CTA per unit
    // (the per-warp query tiles of the first variant must fit in shared memory: d <= 800)
    const bool warp_units = nloc > 0 && idx->total_groups / nloc < 24 &&
                            (size_t)LS_WARPS * LS_QB * ld * sizeof(float) <= 200 * 1024;
    if (warp_units) {
      const size_t smem = (size_t)LS_WARPS * LS_QB * ld * sizeof(float);
      SPF_CUDA(cudaFuncSetAttribute(scan_lists_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      scan_lists_kernel<true><<<(unsigned)ceil_div(max_units, LS_WARPS), LS_WARPS * 32, smem, st>>>(la, nlists);
    } else {
      const size_t smem = (size_t)LS_QB * ld * sizeof(float) + (size_t)LS_WARPS * LS_QB * k * 8;
      if (smem > 200 * 1024) return fail(SPF_E_INVALID, "dimension too large for the list-major scan");
      SPF_CUDA(cudaFuncSetAttribute(scan_lists_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      scan_lists_kernel<false><<<(unsigned)max_units, LS_WARPS * 32, smem, st>>>(la, nlists);
    }
    SPF_TRY(check_launch(c, "scan_lists_kernel"));
    merge_units_kernel<<<(unsigned)ceil_div(nq * 32, 256), 256, 0, st>>>(la, nq);
    SPF_TRY(check_launch(c, "merge_units_kernel"));
  } else {
    KernelTimer t(c, "scan");
    if (k <= 32) SPF_TRY(launch_scan<1>(c, a, nq));
    else if (k <= 64) SPF_TRY(launch_scan<2>(c, a, nq));
    else if (k <= 128) SPF_TRY(launch_scan<4>(c, a, nq));
    else {
      // the reference has no cap on k: passes of 128 results, each taking the keys behind the previous one
      for (uint32_t col = 0; col < k; col += 128) {
        ScanArgs pa = a;
        pa.K = k - col < 128u ? k - col : 128u;
        pa.out_stride = k;
        pa.out_col = col;
        SPF_TRY(launch_scan<4>(c, pa, nq));
      }
    }
  }
  SPF_TRY(bound_guard.finish());
  return SPF_OK;
}

// Host queries -> device rows `Qd` (row pitch ld, padding already zeroed in stream order) and the centroid
// probe of those queries.
static int upload_and_probe(spf_index* idx, const float* queries, uint64_t nq, uint32_t nprobe, float prune_factor,
                            float* Qd, uint32_t* probe, float* thr, uint32_t* seqbase) {
  spf_ctx* c = idx->ctx;
  cudaStream_t st = c->stream;
  const uint32_t ld = idx->ld, d = idx->d;
  // Large batches: the queries travel in pieces on the auxiliary stream and the centroid probe of a piece
  // (independent per query) runs while the next piece is still on the bus — 100 k x 128 queries are 0.93 ms
  // of PCIe time that otherwise sits in front of the first kernel.  Every probe call has ~0.25 ms of fixed
  // cost (a host round trip, two dozen small kernels), so two pieces are the optimum: 100 k queries, nprobe
  // 8: 7.81 ms (one piece), 7.55 ms (two), 7.85 ms (four).
  const uint64_t npieces = c->params.search_upload_pieces > 1 ? (uint64_t)c->params.search_upload_pieces : 1;
  const uint64_t piece = nq >= 32768 && npieces > 1 ? ((nq + npieces - 1) / npieces + 127) / 128 * 128 : nq;
  if (piece >= nq) {
    SPF_CUDA(cudaMemcpy2DAsync(Qd, (size_t)ld * 4, queries, (size_t)d * 4, (size_t)d * 4, nq, cudaMemcpyHostToDevice, st));
    SPF_TRY(search_probe(idx, Qd, nq, nprobe, prune_factor, probe, thr, seqbase));
  } else {
    struct Events {
      cudaEvent_t e[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
      ~Events() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
    } ev;
    if (!c->aux_stream) SPF_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    for (cudaEvent_t& x : ev.e) SPF_CUDA(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
    SPF_CUDA(cudaEventRecord(ev.e[8], st));                            // Q is allocated (and zeroed) in stream order
    SPF_CUDA(cudaStreamWaitEvent(c->aux_stream, ev.e[8], 0));
    int rc = SPF_OK;
    uint32_t np = 0;
    for (uint64_t q0 = 0; q0 < nq; q0 += piece, ++np) {
      const uint64_t n = nq - q0 < piece ? nq - q0 : piece;
      if (cudaMemcpy2DAsync(Qd + q0 * ld, (size_t)ld * 4, queries + q0 * d, (size_t)d * 4, (size_t)d * 4, n,
                            cudaMemcpyHostToDevice, c->aux_stream) != cudaSuccess ||
          cudaEventRecord(ev.e[np], c->aux_stream) != cudaSuccess) {
        rc = fail(SPF_E_CUDA, "query upload: %s", cudaGetErrorString(cudaGetLastError()));
        break;
      }
    }
    uint32_t ip = 0;
    for (uint64_t q0 = 0; rc == SPF_OK && ip < np; q0 += piece, ++ip) {
      const uint64_t n = nq - q0 < piece ? nq - q0 : piece;
      if (cudaStreamWaitEvent(st, ev.e[ip], 0) != cudaSuccess) { rc = fail(SPF_E_CUDA, "cudaStreamWaitEvent failed"); break; }
      rc = search_probe(idx, Qd + q0 * ld, n, nprobe, prune_factor, probe + q0 * nprobe, thr + q0, seqbase + q0 * nprobe);
    }
    if (rc != SPF_OK) {                                                // nothing may still write into Q when it is freed
      cudaStreamSynchronize(c->aux_stream);
      return rc;
    }
  }
  return SPF_OK;
}

int spf_search_batch(spf_index* idx, const float* queries, uint64_t nq, uint32_t k, uint32_t nprobe,
                     float prune_factor, uint64_t* ids, float* dists, uint32_t* counts, float* vectors,
                     uint64_t* keys) {
  return spf::guarded([&]() -> int {
  if (!idx || !queries || !ids || !dists || !counts) return fail(SPF_E_INVALID, "spf_search_batch: NULL argument");
  if (k == 0 || k > 1024) return fail(SPF_E_INVALID, "k must be in [1,1024]");
  if (nq == 0) return SPF_OK;
  if (nprobe == 0) nprobe = k;                       // spann_index.rs:164 nearest_n(query, k)
  if (nprobe > idx->nlists) nprobe = idx->nlists;
  if (nprobe > 1024) return fail(SPF_E_INVALID, "nprobe must be <= 1024");
  spf_ctx* c = idx->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  c->kernel_ms.clear();
  const uint32_t ld = idx->ld, d = idx->d, nlists = idx->nlists;

  DevBuf<float> Q, thr, o_dists;
  DevBuf<uint32_t> probe, seqbase, o_counts;
  DevBuf<uint64_t> o_ids;
  DevBuf<unsigned long long> o_keys, o_slots, d_bytes;
  SPF_TRY(Q.alloc(st, (size_t)nq * ld));
  SPF_TRY(thr.alloc(st, nq));
  SPF_TRY(probe.alloc(st, (size_t)nq * nprobe));
  SPF_TRY(seqbase.alloc(st, (size_t)nq * nprobe));
  SPF_TRY(o_ids.alloc(st, (size_t)nq * k));
  SPF_TRY(o_dists.alloc(st, (size_t)nq * k));
  SPF_TRY(o_keys.alloc(st, (size_t)nq * k));
  SPF_TRY(o_slots.alloc(st, (size_t)nq * k));
  SPF_TRY(o_counts.alloc(st, nq));
  SPF_TRY(d_bytes.alloc(st, 1));
  SPF_CUDA(cudaMemsetAsync(d_bytes.p, 0, sizeof(unsigned long long), st));
  if (ld != d) SPF_CUDA(cudaMemsetAsync(Q.p, 0, (size_t)nq * ld * sizeof(float), st));
  SPF_TRY(upload_and_probe(idx, queries, nq, nprobe, prune_factor, Q.p, probe.p, thr.p, seqbase.p));
  SPF_TRY(search_scan(idx, Q.p, nq, k, nprobe, probe.p, thr.p, seqbase.p, o_ids.p, o_dists.p, o_counts.p, o_keys.p, o_slots.p,
                      d_bytes.p));
  DevBuf<float> o_vec;
  if (vectors) {
    SPF_TRY(o_vec.alloc(st, (size_t)nq * k * d));
    const uint64_t tot = nq * k * d;
    gather_vectors_kernel<<<(unsigned)ceil_div(tot, 256), 256, 0, st>>>(idx->vecs, ld, d, o_slots.p, nq * k, o_vec.p);
    SPF_TRY(check_launch(c, "gather_vectors_kernel"));
    SPF_CUDA(cudaMemcpyAsync(vectors, o_vec.p, tot * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  SPF_CUDA(cudaMemcpyAsync(ids, o_ids.p, (size_t)nq * k * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaMemcpyAsync(dists, o_dists.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaMemcpyAsync(counts, o_counts.p, (size_t)nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  if (keys) SPF_CUDA(cudaMemcpyAsync(keys, o_keys.p, (size_t)nq * k * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  unsigned long long bytes = 0;
  SPF_CUDA(cudaMemcpyAsync(&bytes, d_bytes.p, sizeof(bytes), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  idx->last_scan_bytes = bytes;
  return SPF_OK;
  });
}

// Device-side merge of the ranks' partial top-k for this rank's slice of the batch: per query the k
// smallest keys (distance bits << 32 | encounter index, globally consistent) over the `parts`
// sorted partial lists — the reference's final stable sort + truncate (spann_index.rs:188-193)
// applied across list shards.  One thread per query.
__global__ void topk_merge_kernel(uint32_t parts, uint64_t nq, uint32_t k, const unsigned long long* __restrict__ keys,
                                  const uint64_t* __restrict__ ids, const float* __restrict__ dists,
                                  const uint32_t* __restrict__ counts, uint64_t* __restrict__ out_ids,
                                  float* __restrict__ out_dists, uint32_t* __restrict__ out_counts) {
  const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  constexpr uint32_t MAXP = 16;
  uint32_t cur[MAXP];
  for (uint32_t p = 0; p < parts; ++p) cur[p] = 0;
  uint32_t n = 0;
  while (n < k) {
    int bp = -1;
    unsigned long long bk = 0;
    for (uint32_t p = 0; p < parts; ++p) {
      if (cur[p] >= counts[(size_t)p * nq + q]) continue;
      const unsigned long long kk = keys[((size_t)p * nq + q) * k + cur[p]];
      if (bp < 0 || kk < bk) { bp = (int)p; bk = kk; }      // strict <: the lower rank wins (keys are unique anyway)
    }
    if (bp < 0) break;
    const size_t src = ((size_t)bp * nq + q) * k + cur[bp];
    out_ids[q * k + n] = ids[src];
    out_dists[q * k + n] = dists[src];
    ++cur[bp];
    ++n;
  }
  out_counts[q] = n;
  for (uint32_t e = n; e < k; ++e) {
    out_ids[q * k + e] = ~0ull;
    out_dists[q * k + e] = __int_as_float(0x7f800000);
  }
}

// List-sharded batched search over the ranks of `comm`: every rank holds the
// posting lists [list_begin, list_end) it packed and all centroids.  Each rank brings nq_local
// queries of the batch (the same count on every rank); the call returns the global top-k of THOSE
// queries.  On the device: the query slices are all-gathered, each rank probes its own slice and
// the probe tables are all-gathered (the probe is computed once, not world times), every rank
// scans its lists for the whole batch, the partial top-k travel to the rank that owns the query
// (one personalised exchange), and a device merge produces the result.
int spf_search_sharded(spf_index* idx, spf_comm* comm, const float* queries, uint64_t nq_local, uint32_t k,
                       uint32_t nprobe, float prune_factor, uint64_t* ids, float* dists, uint32_t* counts) {
  return spf::guarded([&]() -> int {
  if (!idx || !queries || !ids || !dists || !counts) return fail(SPF_E_INVALID, "spf_search_sharded: NULL argument");
  if (comm && comm->ctx != idx->ctx) return fail(SPF_E_INVALID, "communicator and index belong to different contexts");
  if (k == 0 || k > 1024) return fail(SPF_E_INVALID, "k must be in [1,1024]");
  if (nq_local == 0) return fail(SPF_E_INVALID, "every rank must bring at least one query");
  if (nprobe == 0) nprobe = k;
  if (nprobe > idx->nlists) nprobe = idx->nlists;
  if (nprobe > 1024) return fail(SPF_E_INVALID, "nprobe must be <= 1024");
  const uint32_t world = comm ? (uint32_t)comm->world : 1u, rank = comm ? (uint32_t)comm->rank : 0u;
  if (world > 16) return fail(SPF_E_INVALID, "at most 16 ranks");
  spf_ctx* c = idx->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  c->kernel_ms.clear();
  const uint32_t ld = idx->ld, d = idx->d;
  const uint64_t nq = nq_local * world;

  DevBuf<float> Q, thr, o_dists, r_dists, m_dists;
  DevBuf<uint32_t> probe, seqbase, o_counts, r_counts, m_counts;
  DevBuf<uint64_t> o_ids, r_ids, m_ids;
  DevBuf<unsigned long long> o_keys, o_slots, d_bytes, r_keys;
  SPF_TRY(Q.alloc(st, (size_t)nq * ld));
  SPF_TRY(thr.alloc(st, nq));
  SPF_TRY(probe.alloc(st, (size_t)nq * nprobe));
  SPF_TRY(seqbase.alloc(st, (size_t)nq * nprobe));
  SPF_TRY(o_ids.alloc(st, (size_t)nq * k));
  SPF_TRY(o_dists.alloc(st, (size_t)nq * k));
  SPF_TRY(o_keys.alloc(st, (size_t)nq * k));
  SPF_TRY(o_slots.alloc(st, (size_t)nq * k));
  SPF_TRY(o_counts.alloc(st, nq));
  SPF_TRY(d_bytes.alloc(st, 1));
  SPF_TRY(r_ids.alloc(st, (size_t)nq * k));
  SPF_TRY(r_dists.alloc(st, (size_t)nq * k));
  SPF_TRY(r_keys.alloc(st, (size_t)nq * k));
  SPF_TRY(r_counts.alloc(st, nq));
  SPF_TRY(m_ids.alloc(st, (size_t)nq_local * k));
  SPF_TRY(m_dists.alloc(st, (size_t)nq_local * k));
  SPF_TRY(m_counts.alloc(st, nq_local));
  SPF_CUDA(cudaMemsetAsync(d_bytes.p, 0, sizeof(unsigned long long), st));
  float* Qmine = Q.p + (size_t)rank * nq_local * ld;
  if (ld != d) SPF_CUDA(cudaMemsetAsync(Qmine, 0, (size_t)nq_local * ld * sizeof(float), st));
  // own slice: upload and probe (in pieces for large slices), then the slices travel to every rank
  SPF_TRY(upload_and_probe(idx, queries, nq_local, nprobe, prune_factor, Qmine, probe.p + (size_t)rank * nq_local * nprobe,
                           thr.p + (size_t)rank * nq_local, seqbase.p + (size_t)rank * nq_local * nprobe));
  {
    KernelTimer t(c, "exchange");
    SPF_TRY(comm_allgather(c, comm, Qmine, Q.p, (size_t)nq_local * ld * sizeof(float)));
  }
  {
    KernelTimer t(c, "exchange");
    SPF_TRY(comm_group_start(comm));
    int rc = comm_allgather(c, comm, probe.p + (size_t)rank * nq_local * nprobe, probe.p, (size_t)nq_local * nprobe * 4);
    if (rc == SPF_OK) rc = comm_allgather(c, comm, seqbase.p + (size_t)rank * nq_local * nprobe, seqbase.p, (size_t)nq_local * nprobe * 4);
    if (rc == SPF_OK) rc = comm_allgather(c, comm, thr.p + (size_t)rank * nq_local, thr.p, (size_t)nq_local * 4);
    const int rc2 = comm_group_end(comm);
    if (rc != SPF_OK) return rc;
    SPF_TRY(rc2);
  }
  SPF_TRY(search_scan(idx, Q.p, nq, k, nprobe, probe.p, thr.p, seqbase.p, o_ids.p, o_dists.p, o_counts.p, o_keys.p, o_slots.p,
                      d_bytes.p, comm));
  {
    KernelTimer t(c, "exchange");                          // the four tables travel as one aggregated operation
    SPF_TRY(comm_group_start(comm));
    int rc = comm_alltoall(c, comm, o_keys.p, r_keys.p, (size_t)nq_local * k * 8);
    if (rc == SPF_OK) rc = comm_alltoall(c, comm, o_ids.p, r_ids.p, (size_t)nq_local * k * 8);
    if (rc == SPF_OK) rc = comm_alltoall(c, comm, o_dists.p, r_dists.p, (size_t)nq_local * k * 4);
    if (rc == SPF_OK) rc = comm_alltoall(c, comm, o_counts.p, r_counts.p, (size_t)nq_local * 4);
    const int rc2 = comm_group_end(comm);
    if (rc != SPF_OK) return rc;
    SPF_TRY(rc2);
  }
  {
    KernelTimer t(c, "merge");
    topk_merge_kernel<<<(unsigned)ceil_div(nq_local, 128), 128, 0, st>>>(world, nq_local, k, r_keys.p, r_ids.p, r_dists.p,
                                                                         r_counts.p, m_ids.p, m_dists.p, m_counts.p);
    SPF_TRY(check_launch(c, "topk_merge_kernel"));
  }
  SPF_CUDA(cudaMemcpyAsync(ids, m_ids.p, (size_t)nq_local * k * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaMemcpyAsync(dists, m_dists.p, (size_t)nq_local * k * sizeof(float), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaMemcpyAsync(counts, m_counts.p, (size_t)nq_local * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  unsigned long long bytes = 0;
  SPF_CUDA(cudaMemcpyAsync(&bytes, d_bytes.p, sizeof(bytes), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  idx->last_scan_bytes = bytes;
  return SPF_OK;
  });
}

// Host-side merge of per-rank partial top-k (the reference's final stable sort, spann_index.rs:
// 188-193, applied across list shards).  Keys are globally consistent, so the k smallest win.
int spf_topk_merge(uint32_t parts, uint64_t nq, uint32_t k, const uint64_t* keys, const uint64_t* ids,
                   const float* dists, const uint32_t* counts, uint64_t* out_ids, float* out_dists,
                   uint32_t* out_counts) {
  return spf::guarded([&]() -> int {
  if (!keys || !ids || !dists || !counts || !out_ids || !out_dists || !out_counts)
    return fail(SPF_E_INVALID, "spf_topk_merge: NULL argument");
  if (parts == 0 || k == 0) return fail(SPF_E_INVALID, "parts and k must be > 0");
  std::vector<uint32_t> cur(parts);
  for (uint64_t q = 0; q < nq; ++q) {
    for (uint32_t p = 0; p < parts; ++p) cur[p] = 0;
    uint32_t n = 0;
    while (n < k) {
      int bp = -1;
      uint64_t bk = 0;
      for (uint32_t p = 0; p < parts; ++p) {
        if (cur[p] >= counts[(size_t)p * nq + q]) continue;
        const uint64_t kk = keys[((size_t)p * nq + q) * k + cur[p]];
        if (bp < 0 || kk < bk) { bp = (int)p; bk = kk; }
      }
      if (bp < 0) break;
      const size_t src = ((size_t)bp * nq + q) * k + cur[bp];
      out_ids[q * k + n] = ids[src];
      out_dists[q * k + n] = dists[src];
      ++cur[bp];
      ++n;
    }
    out_counts[q] = n;
    for (uint32_t e = n; e < k; ++e) {
      out_ids[q * k + e] = ~0ull;
      out_dists[q * k + e] = __builtin_inff();
    }
  }
  return SPF_OK;
  });
}

}  // extern "C"
