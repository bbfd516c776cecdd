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
// recomputation when its whole interval [d-E, d+E] lies on one side of the test; otherwise its
// direct-form distance is recomputed (same rounding sequence as the reference).
//
// Three flat kernels (none of them serialises a warp on one point's dependent chain):
//   classify   warp per point, one sweep over its group records: finds the elements that can be the
//              argmin ("band": within 2E of the producer's approximate minimum), bounds the
//              threshold from the band, and sorts every other element into certain member /
//              certainly not / needs an exact value; survivors go to a short list (16 B entries)
//   exact_eval gathers all short-list entries that need an exact distance and evaluates them 32 at
//              a time per warp, one pair per lane (terms staged in parallel, one sequential chain
//              per lane) — the only place distances are recomputed
//   finalize   warp per point, lane per short-list entry: exact (dmin, best), the remaining tests
//              on exact values, member list
#include <cub/cub.cuh>

#include <algorithm>

#include "kernels.cuh"
#include "pairdist.cuh"

namespace spf {


// Short-list entry (classify → exact_eval → finalize).
struct __align__(16) ShortEnt {
  uint32_t j;       // centroid slot
  float v;          // distance: exact when SE_EXACT, else approximate (within E)
  float cc;         // d(c_best, c_j) when the kind is SE_TEST_CC
  uint32_t flags;
};
constexpr uint32_t SE_EXACT = 1u;       // v is the exact direct-form distance
constexpr uint32_t SE_NEED_EVAL = 2u;   // exact_eval must replace v
constexpr uint32_t SE_KIND_MASK = 3u << 2;
constexpr uint32_t SE_BAND = 0u << 2;       // can be the argmin (always exact after exact_eval)
constexpr uint32_t SE_MEMBER = 1u << 2;     // certain boundary member whatever the exact values are
constexpr uint32_t SE_TEST_CC = 2u << 2;    // d < thr and cc >= d still to be decided; cc stored
constexpr uint32_t SE_TEST_NOCC = 3u << 2;  // same, best was not known yet: cc fetched by finalize

// Per-call state shared by the chunks of one assign (kernels.cuh: resolve_begin / _chunk / _finish).
struct ResolveState {
  DevBuf<uint32_t> ovf_rows, ovf_count, sl_cnt, work_count;
  DevBuf<uint32_t> memlist;      // m_total x (1 << sl_shift) member slots, kept until the CSR is built
  DevBuf<ShortEnt> sl;           // one chunk's short lists
  DevBuf<uint2> work;
  int sl_shift = 0;
  uint32_t work_cap = 0;
  uint64_t chunk_rows = 0;
};

namespace {

struct ResolveDev {
  const float* P; uint32_t m; const float* C; uint32_t k; uint32_t ld;
  float factor;
  CandRec* rec; const RowInfo* info; int cap; int nseg;
  const float* xnorm; const float* xres; const float* cstat; const float* cc;
  const float* seed;
  const float* penalty;    // balanced assignment (extension): cost = fl(d + penalty[j]); NULL otherwise
  uint32_t eld;            // row length entering the certified error bound (ld, plus slack in balanced mode)
  uint32_t* best; float* dmin; uint32_t* nmem;
  uint32_t* ovf_rows; uint32_t* ovf_count;
  int want_members;
  ShortEnt* sl; uint32_t* sl_cnt; int sl_shift;   // short list: 1 << sl_shift (<= 64) entries per point
  uint32_t* memlist;                              // member slots of this chunk's points, same stride
  // work list of exact evaluations: (short-list index, centroid slot) pairs appended by classify
  uint2* work; uint32_t* work_count; uint32_t work_cap;
  uint32_t row_base;   // first row of this chunk in the whole point list (ovf_rows holds global rows)
};

__device__ __forceinline__ bool lex_less(float d1, uint32_t j1, float d2, uint32_t j2) {
  return d1 < d2 || (d1 == d2 && j1 < j2);
}

constexpr int RS_WARPS = 8;           // warps per CTA of classify / finalize
constexpr int RS_EL = 8;              // candidate elements a lane holds per sweep step (2 records)
constexpr int EV_WARPS = 2;           // warps per CTA of exact_eval (33 KB of staging each)
constexpr int EV_CHUNK = 128;         // dimensions staged per step (one 16-byte cp.async per lane and row)
constexpr int EV_STRIDE = EV_CHUNK + 4;
constexpr int EV_QUEUE = 64;          // queued pairs per warp (32 are evaluated at a time)

// The element term and the running combination of the three metrics, split so that the terms
// of one pair can be produced by 32 lanes in parallel while ONE lane folds them in dimension
// order: acc = comb(acc, term(a_i, b_i)) for i = 0, 1, ... is exactly the reference's sequential
// f32 chain (src/distances/distance.rs:16-43; un-fused sub / mul / add).
template <int METRIC>
__device__ __forceinline__ float dist_term(float a, float b) {
  const float df = __fsub_rn(a, b);
  return METRIC == SPF_METRIC_EUCLIDEAN ? __fmul_rn(df, df) : fabsf(df);
}
template <int METRIC>
__device__ __forceinline__ float dist_comb(float acc, float t) {
  return METRIC == SPF_METRIC_CHEBYSHEV ? fmaxf(acc, t) : __fadd_rn(acc, t);
}

// ---------------------------------------------------------------------------------------------
// classify
// ---------------------------------------------------------------------------------------------
constexpr int CLS_STAGE = 256;        // candidate elements a point may keep after the coarse filter
constexpr int CLS_QUEUE = 96;         // pending exact evaluations per warp (flushed 32 at a time)

struct ClsSmem {
  uint32_t j[CLS_STAGE];              // centroid slot of a staged element
  float v[CLS_STAGE];                 // its distance (approximate ones already shifted by |x|^2)
  uint2 q[CLS_QUEUE];
};

// Warp per point.  Level A walks the point's group records (a lane holds one record of each
// segment per step) and keeps the ELEMENTS that can matter — in the band of possible minima or
// with an interval reaching below the loosest possible threshold — compacted into shared memory.
// Level B then works lane-per-element on the survivors (a dozen or two): band statistics, the
// threshold interval, the centroid-centroid test, and the short-list entry of every element that
// is not certainly out.  Records are either all exact (CUDA-core producer) or all approximate
// (tensor producer), so exactness is a property of the call, not of the element.
template <bool APPROX>
__global__ void __launch_bounds__(RS_WARPS * 32, 3) classify_kernel(ResolveDev a) {
  __shared__ ClsSmem smem[RS_WARPS];
  ClsSmem& sm = smem[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const unsigned below = (1u << lane) - 1u;
  const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
  const float INF = __int_as_float(0x7f800000);
  const uint32_t segcap = (uint32_t)a.cap / (uint32_t)a.nseg;
  const uint32_t slcap = 1u << a.sl_shift;
  constexpr bool approx = APPROX;   // tensor producer (a.xnorm != nullptr): a property of the call
  const uint32_t ent_flags = approx ? 0u : SE_EXACT;
  const float cnmax = approx ? a.cstat[0] : 0.0f, dcmax = approx ? a.cstat[1] : 0.0f;

  struct Rec2 { float4 t0, t1; uint32_t g0, g1; };
  auto load_recs = [&](const CandRec* cr, uint32_t i, uint32_t n0, uint32_t n1) {
    Rec2 rc;
    rc.t0 = rc.t1 = make_float4(0.f, 0.f, 0.f, 0.f);
    rc.g0 = rc.g1 = 0;
    // records are read exactly once: streaming loads keep them from evicting the k x k centroid
    // matrix (read with a data-dependent pattern later in this kernel) out of L2
    if (i < n0 && i < segcap) { rc.t0 = __ldcs(&cr[i].t); rc.g0 = __ldcs(&cr[i].g); }
    if (i < n1 && i < segcap) { rc.t1 = __ldcs(&cr[segcap + i].t); rc.g1 = __ldcs(&cr[segcap + i].g); }
    return rc;
  };
  // Two-level prefetch: a row's info word and norms are fetched two rows ahead, its first 32
  // records of each segment (only the live ones, the counts are known by then) one row ahead, so
  // the dependent HBM round trips of the next rows overlap the work on the current one.
  struct Pre1 { RowInfo info; float xn, xr; };
  auto prefetch1 = [&](uint32_t rr) {
    Pre1 p;
    p.info = make_uint4(0u, 0u, 0u, 0u);
    p.xn = p.xr = 0.f;
    if (rr < a.m) {
      p.info = a.info[rr];
      if (approx) { p.xn = a.xnorm[rr]; p.xr = a.xres[rr]; }
    }
    return p;
  };
  auto prefetch2 = [&](uint32_t rr, const Pre1& p1) {
    if (rr < a.m)
      return load_recs(a.rec + (size_t)rr * a.cap, (uint32_t)lane, p1.info.x, a.nseg > 1 ? p1.info.z : 0u);
    return Rec2{make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), 0u, 0u};
  };
  uint32_t nq = 0;                                   // queued exact evaluations (warp-uniform)
  // moves 32 queued items (or all of them when `all`) to the global work list
  auto flush_queue = [&](bool all) {
    while (nq >= 32u || (all && nq > 0)) {
      const uint32_t take = nq < 32u ? nq : 32u;
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(a.work_count, take);
      base = __shfl_sync(0xffffffffu, base, 0);
      // items that do not fit stay flagged SE_NEED_EVAL: finalize recomputes those itself
      if ((uint32_t)lane < take && base + lane < a.work_cap) a.work[base + lane] = sm.q[nq - take + lane];
      nq -= take;
      __syncwarp();
    }
  };
  uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  Pre1 cur1 = prefetch1(r);
  Pre1 nx1 = prefetch1(r + warps_total);
  Rec2 cur_rc = prefetch2(r, cur1);
  for (; r < a.m; r += warps_total) {
    const RowInfo info = cur1.info;
    const float xn = cur1.xn, xr = cur1.xr;
    const Rec2 rc0 = cur_rc;
    cur1 = nx1;
    nx1 = prefetch1(r + 2 * warps_total);
    cur_rc = prefetch2(r + warps_total, cur1);
    const uint32_t cnt0 = info.x, cnt1 = a.nseg > 1 ? info.z : 0u;
    bool overflow = cnt0 > segcap || cnt1 > segcap;   // the dense fallback owns such rows
    uint32_t steps = overflow ? 0u : (max(cnt0, cnt1) + 31u) >> 5;
    const CandRec* cr = a.rec + (size_t)r * a.cap;
    const float E = approx ? tc_err_bound(xn, xr, cnmax, dcmax, a.eld) : 0.0f;
    // a seeded candidate pass is only valid when the observed maximum reaches the seed's bound
    if (approx && a.seed != nullptr) {
      const float sd = a.seed[r];
      if (sd < INF && !(fmaxf(__uint_as_float(info.y), __uint_as_float(info.w)) >= tc_seed_bound(xn, sd, E, cnmax))) {
        overflow = true;
        steps = 0;
      }
    }
    // smallest distance the producer saw: approximate on the tensor path (the info words hold the
    // largest s = x.c - |c|^2/2 per column half, d = |x|^2 - 2 s; the true minimum is within E of
    // it, so only elements within 2E can be the argmin), exact otherwise
    const float ma = approx ? fmaf(-2.0f, fmaxf(__uint_as_float(info.y), __uint_as_float(info.w)), xn)
                            : __uint_as_float(info.y);
    const float band = __fadd_ru(ma, 2.0f * E);
    // loosest possible threshold: thr = fl(dmin * factor) with dmin <= ma + E
    const float thi_loose = a.want_members ? __fmul_ru(__fadd_ru(ma, E), a.factor) : 0.0f;
    // an element matters only if it is in the band or its interval reaches below the threshold
    const float vbound = fmaxf(band, __fadd_ru(thi_loose, E));

    // ---- level A: element filter, survivors compacted into shared memory ----------------------
    uint32_t nel = 0;
    __syncwarp();
    for (uint32_t st = 0; st < steps; ++st) {
      const uint32_t i = st * 32 + lane;
      const Rec2 rc = st == 0 ? rc0 : load_recs(cr, i, cnt0, cnt1);
      const float tv[8] = {rc.t0.x, rc.t0.y, rc.t0.z, rc.t0.w, rc.t1.x, rc.t1.y, rc.t1.z, rc.t1.w};
      const uint32_t jb0 = (rc.g0 & REC_G_MASK) << 2, jb1 = (rc.g1 & REC_G_MASK) << 2;
      float v[8];
      uint32_t pass = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        v[q] = approx ? fmaf(-2.0f, tv[q], xn) : tv[q];
        pass |= v[q] <= vbound ? (1u << q) : 0u;                // NaN never passes (never a member)
      }
      // live elements, per record: the record exists and the slot is a real centroid (only the last
      // group of four can reach past k)
      const uint32_t lm0 = i < cnt0 ? (jb0 + 3u < a.k ? 0xfu : (jb0 < a.k ? (1u << (a.k - jb0)) - 1u : 0u)) : 0u;
      const uint32_t lm1 = i < cnt1 ? (jb1 + 3u < a.k ? 0xfu : (jb1 < a.k ? (1u << (a.k - jb1)) - 1u : 0u)) : 0u;
      pass &= lm0 | (lm1 << 4);
      const uint32_t mine = (uint32_t)__popc(pass);
      uint32_t incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      uint32_t pos = nel + incl - mine;
      nel += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if ((pass >> q) & 1u) {
          if (pos < (uint32_t)CLS_STAGE) {
            sm.j[pos] = (q < 4 ? jb0 : jb1) + (uint32_t)(q & 3);
            sm.v[pos] = v[q];
          }
          ++pos;
        }
      }
    }
    __syncwarp();
    if (nel > (uint32_t)CLS_STAGE) { overflow = true; nel = 0; }
    const uint32_t steps2 = (nel + 31u) >> 5;

    // ---- level B, sweep 1: the band: how many elements, are they all exact, which one ------------
    uint32_t n_in = 0, j_any = 0;
    float bd = INF;
    uint32_t bj = 0xffffffffu;
    // the first 32 staged elements stay in registers for sweep 2
    const uint32_t j0 = (uint32_t)lane < nel ? sm.j[lane] : 0u;
    const float v0 = (uint32_t)lane < nel ? sm.v[lane] : INF;
    for (uint32_t st = 0; st < steps2; ++st) {
      const uint32_t e = st * 32 + lane;
      const bool have = e < nel;
      const uint32_t j = st == 0 ? j0 : (have ? sm.j[e] : 0u);
      const float v = st == 0 ? v0 : (have ? sm.v[e] : INF);
      const bool inb = have && v <= band;
      n_in += (uint32_t)__popc(__ballot_sync(0xffffffffu, inb));
      if (inb) {
        j_any = j;
        if (!approx && lex_less(v, j, bd, bj)) { bd = v; bj = j; }
      }
    }
    bool best_known, bd_known;
    if (!approx) {             // every band element is exact: the best and its distance are known
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, o);
        const uint32_t oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (lex_less(od, oj, bd, bj)) { bd = od; bj = oj; }
      }
      if (!(bd < INF)) {                        // fold identity (0, +inf): nothing was < inf
        // (a finite producer minimum without a record at it can only come from a seed that was not
        // the distance of a real centroid: the dense fallback decides such a row)
        if (ma < INF) overflow = true;
        bd = INF; bj = 0;
      }
      best_known = bd_known = true;
    } else if (n_in == 1) {    // a single approximate element can be the argmin: it is the best
      bj = __reduce_max_sync(0xffffffffu, j_any);
      best_known = true;
      bd_known = false;
    } else if (n_in == 0) {    // no finite candidate at all: the fold identity
      bd = INF; bj = 0;
      best_known = bd_known = true;
    } else {
      best_known = bd_known = false;
    }
    // bounds of thr = fl(dmin * factor): dmin lies in [max(ma - E, 0), ma + E]
    const float tlo = bd_known ? __fmul_rn(bd, a.factor) : __fmul_rd(fmaxf(__fadd_rd(ma, -E), 0.0f), a.factor);
    const float thi = bd_known ? tlo : thi_loose;
    const float* ccrow = (best_known && a.cc) ? a.cc + (size_t)bj * a.k : nullptr;

    // ---- level B, sweep 2: classification, survivors appended to the short list --------------------
    ShortEnt* sl = a.sl + ((size_t)r << a.sl_shift);
    uint32_t out = 0;
    for (uint32_t st = 0; st < steps2; ++st) {
      const uint32_t e = st * 32 + lane;
      const bool have = e < nel;
      const uint32_t j = st == 0 ? j0 : (have ? sm.j[e] : 0u);
      const float v = st == 0 ? v0 : (have ? sm.v[e] : INF);
      const bool inb = have && v <= band;
      const float lo = approx ? __fadd_rd(v, -E) : v, hi = approx ? __fadd_ru(v, E) : v;
      bool keep = inb;
      uint32_t fl = SE_BAND | (approx ? SE_NEED_EVAL : SE_EXACT);
      float ccv = 0.f;
      if (a.want_members && have && !inb && lo < thi) {      // else certainly d >= thr → not a member
        const bool t1c = hi < tlo;                            // certainly d < thr
        if (ccrow != nullptr) {
          ccv = ccrow[j];
          if (ccv >= lo) {                                    // otherwise certainly cc < d → not a member
            keep = true;
            const bool certain = t1c && ccv >= hi;
            fl = certain ? (SE_MEMBER | ent_flags) : (SE_TEST_CC | (approx ? SE_NEED_EVAL : SE_EXACT));
          }
        } else {                                              // best (or the cc matrix) not available here
          keep = true;
          fl = SE_TEST_NOCC | (approx ? (t1c ? 0u : SE_NEED_EVAL) : SE_EXACT);
        }
      }
      // append: lane-per-entry, ballot compaction
      const unsigned kb = __ballot_sync(0xffffffffu, keep);
      const uint32_t pos = out + (uint32_t)__popc(kb & below);
      if (keep && pos < slcap) {
        ShortEnt w;
        w.j = j; w.v = v; w.cc = ccv; w.flags = fl;
        sl[pos] = w;
      }
      // queue the entries that need an exact value: slot (30 bits) + the entry kind (top 2 bits),
      // so exact_eval can store the final flags
      const bool need = keep && pos < slcap && (fl & SE_NEED_EVAL);
      const unsigned nbal = __ballot_sync(0xffffffffu, need);
      if (nbal) {
        const uint32_t ntot = (uint32_t)__popc(nbal);
        if (nq + ntot > (uint32_t)CLS_QUEUE) flush_queue(true);
        if (need) sm.q[nq + (uint32_t)__popc(nbal & below)] = make_uint2((r << a.sl_shift) + pos, j | ((fl & SE_KIND_MASK) << 28));
        nq += ntot;
        __syncwarp();
        flush_queue(false);
      }
      out += (uint32_t)__popc(kb);
    }
    overflow = overflow || out > slcap;
    if (lane == 0) {
      if (overflow) {
        const uint32_t p2 = atomicAdd(a.ovf_count, 1u);
        a.ovf_rows[p2] = a.row_base + r;
        a.nmem[r] = NMEM_OVERFLOW_BIT;
        a.sl_cnt[r] = 0;
      } else {
        a.sl_cnt[r] = out;
      }
    }
  }
  flush_queue(true);
}

// ---------------------------------------------------------------------------------------------
// exact_eval
// ---------------------------------------------------------------------------------------------
struct EvalSmem {
  float xs[32][EV_STRIDE];   // staged point rows (one 128-dimension chunk), one row per queued pair
  float cs[32][EV_STRIDE];   // staged centroid rows
  uint32_t qd[EV_QUEUE];     // short-list index of the queued pair (point = qd >> sl_shift)
  uint32_t qj[EV_QUEUE];     // centroid slot
};

// Evaluates the first np (<= 32) queued pairs.  The warp copies both rows of every pair into
// shared memory with cp.async (one coalesced 512-byte request per row, all 2*np of them in flight
// together, no registers held); then lane p walks pair p in dimension order, so every value is
// the reference's sequential f32 sum, and stores the exact distance into the short-list entry.
template <int METRIC>
__device__ __forceinline__ void eval_queue(const ResolveDev& a, EvalSmem& sm, uint32_t np, int lane) {
  float acc = 0.0f;
  for (uint32_t c0 = 0; c0 < a.ld; c0 += EV_CHUNK) {
    const uint32_t col = c0 + lane * 4;
    if (col < a.ld) {                                // ld is a multiple of 4
      for (uint32_t p = 0; p < np; ++p) {
        const float* x = a.P + (size_t)(sm.qd[p] >> a.sl_shift) * a.ld + col;
        const float* y = a.C + (size_t)(sm.qj[p] & 0x3fffffffu) * a.ld + col;
        const uint32_t dx = (uint32_t)__cvta_generic_to_shared(&sm.xs[p][lane * 4]);
        const uint32_t dy = (uint32_t)__cvta_generic_to_shared(&sm.cs[p][lane * 4]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dx), "l"(x) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dy), "l"(y) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if ((uint32_t)lane < np) {
      const int n4 = (int)(((a.ld - c0) < (uint32_t)EV_CHUNK ? (a.ld - c0) : (uint32_t)EV_CHUNK) >> 2);
      const float4* xp = reinterpret_cast<const float4*>(&sm.xs[lane][0]);
      const float4* yp = reinterpret_cast<const float4*>(&sm.cs[lane][0]);
#pragma unroll 8
      for (int i = 0; i < n4; ++i) {
        const float4 xv = xp[i], yv = yp[i];
        acc = dist_step<METRIC>(acc, xv.x, yv.x);
        acc = dist_step<METRIC>(acc, xv.y, yv.y);
        acc = dist_step<METRIC>(acc, xv.z, yv.z);
        acc = dist_step<METRIC>(acc, xv.w, yv.w);
      }
    }
    __syncwarp();
  }
  if ((uint32_t)lane < np) {
    ShortEnt* w = a.sl + sm.qd[lane];
    if (a.penalty) acc = __fadd_rn(acc, a.penalty[sm.qj[lane] & 0x3fffffffu]);
    w->v = acc;
    w->flags = ((sm.qj[lane] >> 30) << 2) | SE_EXACT;
  }
  __syncwarp();
}

template <int METRIC>
__global__ void __launch_bounds__(EV_WARPS * 32) exact_eval_kernel(ResolveDev a) {
  extern __shared__ __align__(16) unsigned char ev_raw[];
  EvalSmem& sm = reinterpret_cast<EvalSmem*>(ev_raw)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
  const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t total = min(*a.work_count, a.work_cap);
  // 32 work items per warp and round; the next round's items are fetched before this round runs
  uint32_t base = warp_global * 32;
  uint2 item = (base + lane < total) ? a.work[base + lane] : make_uint2(0u, 0u);
  for (; base < total; base += warps_total * 32) {
    const uint32_t np = min(32u, total - base);
    sm.qd[lane] = item.x;
    sm.qj[lane] = item.y;
    const uint32_t nb = base + warps_total * 32;
    item = (nb + lane < total) ? a.work[nb + lane] : make_uint2(0u, 0u);
    __syncwarp();
    eval_queue<METRIC>(a, sm, np, lane);
  }
}

// ---------------------------------------------------------------------------------------------
// finalize
// ---------------------------------------------------------------------------------------------
// A short list has a dozen entries on average (at most 64), so a point gets half a warp: two points
// per warp, FN_MAX entries per lane.  All warp-wide primitives run with the full mask and stay inside
// the 16-lane group by construction (xor / up shuffles with offsets < 16, ballots masked per group).
template <int METRIC, int FN_LANES>    // FN_LANES: lanes per point
__global__ void __launch_bounds__(RS_WARPS * 32) finalize_kernel(ResolveDev a) {
  constexpr int FN_GROUPS = 32 / FN_LANES;
  constexpr int FN_MAX = 64 / FN_LANES;  // short-list entries per lane (short list <= 64 entries)
  const int lane = threadIdx.x & 31;
  const int gl = lane & (FN_LANES - 1);                                  // lane within the point's group
  const int grp = lane / FN_LANES;                                       // which point of the warp
  const unsigned gmask = (FN_LANES == 32 ? 0xffffffffu : ((1u << FN_LANES) - 1u)) << (grp * FN_LANES);
  const uint32_t stride = ((gridDim.x * blockDim.x) >> 5) * FN_GROUPS;   // points per grid sweep
  const float INF = __int_as_float(0x7f800000);
  const bool approx = a.xnorm != nullptr;
  // two-level prefetch: a row's entry count and overflow mark two rows ahead, its live entries one
  // row ahead
  struct PreE { ShortEnt en[FN_MAX]; };
  auto prefetch_n = [&](uint32_t rr) { return rr < a.m ? make_uint2(a.sl_cnt[rr], a.nmem[rr]) : make_uint2(0u, 0u); };
  auto prefetch_e = [&](uint32_t rr, uint32_t n) {
    PreE p;
#pragma unroll
    for (int u = 0; u < FN_MAX; ++u) { p.en[u].j = 0; p.en[u].v = 0.f; p.en[u].cc = 0.f; p.en[u].flags = SE_MEMBER; }
    if (rr < a.m) {
      const ShortEnt* sl = a.sl + ((size_t)rr << a.sl_shift);
#pragma unroll
      for (int u = 0; u < FN_MAX; ++u)
        if ((uint32_t)(u * FN_LANES + gl) < n) p.en[u] = sl[u * FN_LANES + gl];
    }
    return p;
  };
  const uint32_t r_first = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * FN_GROUPS;   // the warp's first point
  uint32_t r = r_first + grp;
  uint2 cur_n = prefetch_n(r);
  uint2 nx_n = prefetch_n(r + stride);
  PreE cur_e = prefetch_e(r, cur_n.x);
  for (uint32_t rw = r_first; rw < a.m; rw += stride, r += stride) {      // warp-uniform trip count
    const uint32_t nm = cur_n.y;
    const PreE cur = cur_e;
    // a group without a point (past the end) or whose row belongs to the dense fallback idles through
    // the shuffles with an empty list and writes nothing
    const bool live = r < a.m && !(nm & NMEM_OVERFLOW_BIT);
    const uint32_t n = live ? cur_n.x : 0u;
    cur_n = nx_n;
    nx_n = prefetch_n(r + 2 * stride);
    cur_e = prefetch_e(r + stride, cur_n.x);
    ShortEnt en[FN_MAX];
    float bd = INF;
    uint32_t bj = 0xffffffffu;
#pragma unroll
    for (int u = 0; u < FN_MAX; ++u) {
      const uint32_t e = u * FN_LANES + gl;
      en[u] = cur.en[u];
      if (e < n && (en[u].flags & SE_NEED_EVAL)) {   // did not fit the work list (rare): recompute here
        en[u].v = thread_dist<METRIC>(a.P + (size_t)r * a.ld, a.C + (size_t)en[u].j * a.ld, a.ld);
        if (a.penalty) en[u].v = __fadd_rn(en[u].v, a.penalty[en[u].j]);
        en[u].flags = (en[u].flags & ~SE_NEED_EVAL) | SE_EXACT;
      }
      if (e < n && (en[u].flags & SE_KIND_MASK) == SE_BAND && lex_less(en[u].v, en[u].j, bd, bj)) { bd = en[u].v; bj = en[u].j; }
    }
#pragma unroll
    for (int o = FN_LANES / 2; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, o);
      const uint32_t oj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (lex_less(od, oj, bd, bj)) { bd = od; bj = oj; }
    }
    if (!(bd < INF)) { bd = INF; bj = 0; }   // fold identity (0, +inf): nothing was < inf
    if (gl == 0 && live) {
      a.best[r] = bj;
      a.dmin[r] = bd;
    }
    if (!a.want_members) {                   // uniform over the launch
      if (gl == 0 && live) a.nmem[r] = 1;
      continue;
    }
    const float thr = __fmul_rn(bd, a.factor);
    const float* x = a.P + (size_t)(live ? r : 0u) * a.ld;
    float E = 0.0f;
    uint32_t member = 0;
    bool best_listed = false;
#pragma unroll
    for (int u = 0; u < FN_MAX; ++u) {
      const uint32_t e = u * FN_LANES + gl;
      if (e >= n) continue;
      const uint32_t kind = en[u].flags & SE_KIND_MASK, j = en[u].j;
      const bool ex = (en[u].flags & SE_EXACT) != 0;
      float dv = en[u].v;
      bool mem = false;
      if (j == bj) {
        mem = true;
        best_listed = true;
      } else if (kind == SE_MEMBER) {
        mem = true;
      } else if (kind == SE_TEST_CC) {               // exact by construction
        mem = dv < thr && en[u].cc >= dv;
      } else {                                       // a band element that lost, or best was unknown
        bool t1 = true;                              // an approximate value reaching here was certified < thr
        if (ex) t1 = dv < thr;
        if (t1) {
          const float cc = a.cc ? a.cc[(size_t)bj * a.k + j]
                                : thread_dist<METRIC>(a.C + (size_t)bj * a.ld, a.C + (size_t)j * a.ld, a.ld);
          if (ex) {
            mem = cc >= dv;
          } else {
            if (E == 0.0f && approx) E = tc_err_bound(a.xnorm[r], a.xres[r], a.cstat[0], a.cstat[1], a.eld);
            const float lo = __fadd_rd(dv, -E), hi = __fadd_ru(dv, E);
            if (cc >= hi) mem = true;
            else if (cc >= lo) {                     // undecidable on the interval: recompute (rare)
              dv = thread_dist<METRIC>(x, a.C + (size_t)j * a.ld, a.ld);
              if (a.penalty) dv = __fadd_rn(dv, a.penalty[j]);
              mem = dv < thr && cc >= dv;
            }
          }
        }
      }
      member |= mem ? (1u << u) : 0u;
    }
    // member list (uint32 slots) of the row: prefix sum over the group's lanes
    __syncwarp();
    const uint32_t mine = (uint32_t)__popc(member);
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < FN_LANES; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
      if (gl >= o) incl += up;
    }
    uint32_t out = __shfl_sync(0xffffffffu, incl, grp * FN_LANES + FN_LANES - 1);
    best_listed = (__ballot_sync(0xffffffffu, best_listed) & gmask) != 0u;
    if (live) {
      uint32_t* memlist = a.memlist + ((size_t)r << a.sl_shift);
      uint32_t pos = incl - mine;
#pragma unroll
      for (int u = 0; u < FN_MAX; ++u)
        if ((member >> u) & 1u) memlist[pos++] = en[u].j;
      if (!best_listed) {       // only when every distance was inf/NaN: members = {slot 0}
        if (gl == 0) memlist[0] = 0u;
        out = 1;
      }
      if (gl == 0) a.nmem[r] = out;
    }
  }
}

// Dense fallback for rows whose candidate or short-list buffers overflowed (or whose norms are
// not finite): their exact distances to all k centroids come from the CUDA-core direct-form
// kernel (dense mode), `drow` is that row of the batch.  One CTA per row; PHASE 0 writes
// best / dmin / nmem, PHASE 1 writes the member slots as (key, val) pairs.
template <int METRIC, int PHASE>
__global__ void __launch_bounds__(256)
overflow_rows_kernel(ResolveDev a, const float* __restrict__ dense, const uint32_t* __restrict__ rows,
                     uint32_t nrows, const uint64_t* __restrict__ row_off, uint32_t* __restrict__ keys,
                     uint32_t* __restrict__ vals) {
  __shared__ float s_bd[8];
  __shared__ uint32_t s_bj[8];
  __shared__ unsigned s_cnt;
  const float INF = __int_as_float(0x7f800000);
  for (uint32_t o = blockIdx.x; o < nrows; o += gridDim.x) {
    const uint32_t r = rows[o];
    const float* drow = dense + (size_t)o * a.k;
    float bd = INF;
    uint32_t bj = 0xffffffffu;
    if (PHASE == 0) {
      for (uint32_t j = threadIdx.x; j < a.k; j += blockDim.x) {
        const float dv = a.penalty ? __fadd_rn(drow[j], a.penalty[j]) : drow[j];
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
        const float dv = a.penalty ? __fadd_rn(drow[j], a.penalty[j]) : drow[j];
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

__global__ void gather_rows32_kernel(const float* __restrict__ src, uint32_t ld4, const uint32_t* __restrict__ idx,
                                     uint32_t m, float* __restrict__ dst) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint64_t)m * ld4) return;
  const uint32_t r = (uint32_t)(t / ld4), c = (uint32_t)(t - (uint64_t)r * ld4);
  reinterpret_cast<float4*>(dst)[t] = __ldg(reinterpret_cast<const float4*>(src) + (size_t)idx[r] * ld4 + c);
}

constexpr uint32_t OVF_BATCH = 4096;   // overflow rows per dense batch (4096 x k x 4 bytes of scratch)

// Runs PHASE over all overflow rows, OVF_BATCH at a time.  When `keep` is given (n_ovf x k floats,
// used when that fits the budget) PHASE 0 stores every batch's dense block there and PHASE 1 reuses
// it; otherwise the block is recomputed per phase so the scratch stays bounded.
template <int METRIC, int PHASE>
int run_overflow(spf_ctx* c, const ResolveDev& d, uint32_t n_ovf, float* keep, const uint64_t* row_off,
                 uint32_t* keys, uint32_t* vals) {
  if (n_ovf == 0) return SPF_OK;
  cudaStream_t st = c->stream;
  const uint32_t nb_max = n_ovf < OVF_BATCH ? n_ovf : OVF_BATCH;
  DevBuf<float> rows, dense;
  const bool compute = !(keep && PHASE == 1);
  if (compute) SPF_TRY(rows.alloc(st, (size_t)nb_max * d.ld));
  if (!keep) SPF_TRY(dense.alloc(st, (size_t)nb_max * d.k));
  for (uint32_t b0 = 0; b0 < n_ovf; b0 += OVF_BATCH) {
    const uint32_t nb = (n_ovf - b0) < OVF_BATCH ? (n_ovf - b0) : OVF_BATCH;
    float* blk = keep ? keep + (size_t)b0 * d.k : dense.p;
    if (compute) {
      const uint64_t total = (uint64_t)nb * (d.ld / 4);
      gather_rows32_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(d.P, d.ld / 4, d.ovf_rows + b0, nb, rows.p);
      SPF_TRY(check_launch(c, "gather_rows32_kernel"));
      SPF_TRY(launch_assign_exact(c, METRIC, rows.p, nb, d.C, d.k, d.ld, 1.0f, nullptr, blk));
    }
    const unsigned grid = nb < (uint32_t)c->sm_count * 8 ? nb : (unsigned)c->sm_count * 8;
    overflow_rows_kernel<METRIC, PHASE><<<grid, 256, 0, st>>>(d, blk, d.ovf_rows + b0, nb, row_off, keys, vals);
    SPF_TRY(check_launch(c, "overflow_rows_kernel"));
  }
  return SPF_OK;
}

struct CountOp {
  __host__ __device__ uint64_t operator()(uint32_t v) const { return (uint64_t)(v & ~NMEM_OVERFLOW_BIT); }
};

// Thread per row: a row's member slots start on a 16-byte boundary (stride 1 << sl_shift >= 4 slots
// or a single-slot list), so they are read four at a time; consecutive rows write consecutive
// ranges of the pair arrays.  (A warp per row left 26 of 32 lanes idle at the usual 1-6 members.)
__global__ void fill_pairs_kernel(const uint32_t* __restrict__ memlist, int sl_shift, const uint32_t* __restrict__ nmem,
                                  const uint64_t* __restrict__ row_off, uint32_t m,
                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += stride) {
    const uint32_t nm = nmem[r];
    if (nm & NMEM_OVERFLOW_BIT) continue;
    const uint64_t off = row_off[r];
    const uint32_t* mem = memlist + ((size_t)r << sl_shift);
    if (sl_shift >= 2) {
      for (uint32_t s = 0; s < nm; s += 4) {
        const uint4 q = *reinterpret_cast<const uint4*>(mem + s);
        const uint32_t qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (s + i < nm) { keys[off + s + i] = qv[i]; vals[off + s + i] = r; }
      }
    } else {
      for (uint32_t s = 0; s < nm; ++s) { keys[off + s] = mem[s]; vals[off + s] = r; }
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
int resolve_chunk_t(spf_ctx* c, ResolveState* s, const ResolveArgs& a, uint64_t r0) {
  cudaStream_t st = c->stream;
  if (a.m > s->chunk_rows) return fail(SPF_E_INVALID, "resolve_chunk: chunk larger than planned");
  SPF_CUDA(cudaMemsetAsync(s->work_count.p, 0, sizeof(uint32_t), st));
  SPF_CUDA(cudaMemsetAsync(a.nmem, 0, a.m * sizeof(uint32_t), st));
  ResolveDev d{a.P, (uint32_t)a.m, a.C, a.k, a.ld, a.factor, a.cand.rec, a.cand.info, a.cand.cap, a.nseg,
               a.xnorm, a.xres, a.d_cstat, a.cc, a.seed, a.penalty, a.eld ? a.eld : a.ld, a.best, a.dmin, a.nmem, s->ovf_rows.p,
               s->ovf_count.p,
               a.want_members ? 1 : 0, s->sl.p, s->sl_cnt.p, s->sl_shift,
               a.want_members ? s->memlist.p + ((size_t)r0 << s->sl_shift) : nullptr,
               s->work.p, s->work_count.p, s->work_cap, (uint32_t)r0};
  KernelTimer t(c, "resolve");
  // persistent grids sized to exactly one resident wave (classify: 3 CTAs / SM by registers,
  // finalize: 4), rows are taken grid-stride
  uint64_t blocks = ceil_div(a.m, RS_WARPS);
  const int fn_lanes = c->params.finalize_lanes == 8 ? 8 : (c->params.finalize_lanes == 32 ? 32 : 16);
  uint64_t blocks_cls = blocks, blocks_fin = ceil_div(a.m, RS_WARPS * (32 / fn_lanes));
  int cls_per_sm = 2;
  if (a.xnorm) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cls_per_sm, classify_kernel<true>, RS_WARPS * 32, 0);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cls_per_sm, classify_kernel<false>, RS_WARPS * 32, 0);
  if (cls_per_sm < 1) cls_per_sm = 1;
  if (blocks_cls > (uint64_t)c->sm_count * cls_per_sm) blocks_cls = (uint64_t)c->sm_count * cls_per_sm;
  if (blocks_fin > (uint64_t)c->sm_count * 8) blocks_fin = (uint64_t)c->sm_count * 8;
  {
    KernelTimer t2(c, "classify");
    if (a.xnorm) classify_kernel<true><<<(unsigned)blocks_cls, RS_WARPS * 32, 0, st>>>(d);
    else classify_kernel<false><<<(unsigned)blocks_cls, RS_WARPS * 32, 0, st>>>(d);
    SPF_TRY(check_launch(c, "classify_kernel"));
  }
  if (a.xnorm) {   // the exact path never queues work
    KernelTimer t2(c, "exact_eval");
    const size_t ev_smem = EV_WARPS * sizeof(EvalSmem);
    SPF_CUDA(cudaFuncSetAttribute(exact_eval_kernel<METRIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ev_smem));
    exact_eval_kernel<METRIC><<<(unsigned)c->sm_count * 3, EV_WARPS * 32, ev_smem, st>>>(d);
    SPF_TRY(check_launch(c, "exact_eval_kernel"));
  }
  {
    KernelTimer t2(c, "finalize");
    if (fn_lanes == 8) finalize_kernel<METRIC, 8><<<(unsigned)blocks_fin, RS_WARPS * 32, 0, st>>>(d);
    else if (fn_lanes == 32) finalize_kernel<METRIC, 32><<<(unsigned)blocks_fin, RS_WARPS * 32, 0, st>>>(d);
    else finalize_kernel<METRIC, 16><<<(unsigned)blocks_fin, RS_WARPS * 32, 0, st>>>(d);
    SPF_TRY(check_launch(c, "finalize_kernel"));
  }
  return SPF_OK;
}

// Overflow rows (all chunks) through the dense fallback, then the cluster-major CSR.  `a` describes
// the whole point list (P, outputs); the candidate fields are not used.
template <int METRIC>
int resolve_finish_t(spf_ctx* c, ResolveState* s, const ResolveArgs& a, CsrOut* csr) {
  cudaStream_t st = c->stream;
  ResolveDev d{a.P, (uint32_t)a.m, a.C, a.k, a.ld, a.factor, nullptr, nullptr, 0, 1,
               a.xnorm, a.xres, a.d_cstat, a.cc, nullptr, a.penalty, a.eld ? a.eld : a.ld, a.best, a.dmin, a.nmem, s->ovf_rows.p,
               s->ovf_count.p,
               a.want_members ? 1 : 0, s->sl.p, s->sl_cnt.p, s->sl_shift, s->memlist.p, s->work.p,
               s->work_count.p, s->work_cap, 0u};
  uint32_t n_ovf = 0;
  SPF_CUDA(cudaMemcpyAsync(&n_ovf, s->ovf_count.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  // dense blocks of the overflow rows are kept between the two phases when they fit 16 GB
  DevBuf<float> keep;
  const bool keep_dense = csr != nullptr && n_ovf > 0 && (uint64_t)n_ovf * a.k * sizeof(float) <= (16ull << 30);
  if (keep_dense) SPF_TRY(keep.alloc(st, (size_t)n_ovf * a.k));
  {
    KernelTimer t2(c, "overflow");
    SPF_TRY((run_overflow<METRIC, 0>(c, d, n_ovf, keep.p, nullptr, nullptr, nullptr)));
  }
  c->last_overflow_rows = n_ovf;
  if (!csr) return SPF_OK;

  KernelTimer t(c, "csr");
  // row offsets = exclusive scan of member counts
  DevBuf<uint64_t> row_off, d_total;
  SPF_TRY(row_off.alloc_cached(c, "csr_row_off", a.m));
  SPF_TRY(d_total.alloc(st, 1));
  cub::TransformInputIterator<uint64_t, CountOp, const uint32_t*> counts(a.nmem, CountOp());
  size_t tmp_bytes = 0;
  uint64_t total = 0;
  {
  KernelTimer ts(c, "csr_scan");
  SPF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, row_off.p, (int64_t)a.m, st));
  DevBuf<uint8_t> tmp;
  SPF_TRY(tmp.alloc_cached(c, "csr_scan_tmp", tmp_bytes));
  SPF_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, counts, row_off.p, (int64_t)a.m, st));
  c->launches += 2;
  total_kernel<<<1, 32, 0, st>>>(row_off.p, a.nmem, (uint32_t)a.m, d_total.p);
  SPF_TRY(check_launch(c, "total_kernel"));
  SPF_CUDA(cudaMemcpyAsync(&total, d_total.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  }

  DevBuf<uint32_t> keys, vals, keys2, vals2;
  SPF_TRY(keys.alloc_cached(c, "csr_keys", total));
  SPF_TRY(vals.alloc_cached(c, "csr_vals", total));
  SPF_TRY(keys2.alloc_cached(c, "csr_keys2", total));
  SPF_TRY(vals2.alloc(st, total));
  DevBuf<uint64_t> offsets;
  SPF_TRY(offsets.alloc(st, (size_t)a.k + 1));
  {
    KernelTimer ts(c, "csr_fill");
    uint64_t blocks = ceil_div(a.m, 256);
    if (blocks > (uint64_t)c->sm_count * 16) blocks = (uint64_t)c->sm_count * 16;
    fill_pairs_kernel<<<(unsigned)blocks, 256, 0, st>>>(s->memlist.p, s->sl_shift, a.nmem, row_off.p, (uint32_t)a.m,
                                                        keys.p, vals.p);
    SPF_TRY(check_launch(c, "fill_pairs_kernel"));
    SPF_TRY((run_overflow<METRIC, 1>(c, d, n_ovf, keep.p, row_off.p, keys.p, vals.p)));
  }
  // stable sort by cluster slot keeps the input order inside every cluster (:353-361)
  int end_bit = 1;
  while (end_bit < 32 && (1ull << end_bit) < (uint64_t)a.k) ++end_bit;
  if (total > 0) {
    KernelTimer ts(c, "csr_sort");
    size_t sort_bytes = 0;
    SPF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys.p, keys2.p, vals.p, vals2.p,
                                             (int64_t)total, 0, end_bit, st));
    DevBuf<uint8_t> stmp;
    SPF_TRY(stmp.alloc_cached(c, "csr_sort_tmp", sort_bytes));
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

int resolve_begin(spf_ctx* c, uint64_t m_total, uint64_t chunk_rows, bool approx, bool want_members,
                  ResolveState** out) {
  cudaStream_t st = c->stream;
  ResolveState* s = new (std::nothrow) ResolveState();
  if (!s) return fail(SPF_E_OOM, "out of host memory");
  s->chunk_rows = chunk_rows;
  s->sl_shift = 0;
  while ((1 << s->sl_shift) < c->params.short_cap) ++s->sl_shift;
  const uint64_t work_cap64 = c->params.work_cap > 0 ? (uint64_t)c->params.work_cap
                                                     : (approx ? chunk_rows * 8 + 1024 : 32);
  s->work_cap = work_cap64 > 0xffffff00ull ? 0xffffff00u : (uint32_t)work_cap64;
  int rc = SPF_OK;
  if ((chunk_rows << s->sl_shift) >= (1ull << 32)) rc = fail(SPF_E_INVALID, "assign: chunk too large");
  if (rc >= 0) rc = s->ovf_rows.alloc_cached(c, "ovf_rows", m_total);
  if (rc >= 0) rc = s->ovf_count.alloc(st, 1);
  if (rc >= 0) rc = s->sl.alloc_cached(c, "short_lists", (size_t)chunk_rows << s->sl_shift);
  if (rc >= 0) rc = s->sl_cnt.alloc_cached(c, "short_counts", chunk_rows);
  if (rc >= 0 && want_members) rc = s->memlist.alloc_cached(c, "member_slots", (size_t)m_total << s->sl_shift);
  if (rc >= 0) rc = s->work.alloc_cached(c, "eval_work", s->work_cap);
  if (rc >= 0) rc = s->work_count.alloc(st, 1);
  if (rc >= 0 && cudaMemsetAsync(s->ovf_count.p, 0, sizeof(uint32_t), st) != cudaSuccess)
    rc = fail(SPF_E_CUDA, "cudaMemsetAsync failed");
  if (rc < 0) { delete s; return rc; }
  *out = s;
  return SPF_OK;
}

void resolve_free(ResolveState* s) { delete s; }

int resolve_chunk(spf_ctx* c, ResolveState* s, const ResolveArgs& a, uint64_t r0) {
  switch (a.metric) {
    case SPF_METRIC_EUCLIDEAN: return resolve_chunk_t<SPF_METRIC_EUCLIDEAN>(c, s, a, r0);
    case SPF_METRIC_MANHATTAN: return resolve_chunk_t<SPF_METRIC_MANHATTAN>(c, s, a, r0);
    case SPF_METRIC_CHEBYSHEV: return resolve_chunk_t<SPF_METRIC_CHEBYSHEV>(c, s, a, r0);
  }
  return fail(SPF_E_INVALID, "unknown metric %d", a.metric);
}

int resolve_finish(spf_ctx* c, ResolveState* s, const ResolveArgs& a, CsrOut* csr) {
  switch (a.metric) {
    case SPF_METRIC_EUCLIDEAN: return resolve_finish_t<SPF_METRIC_EUCLIDEAN>(c, s, a, csr);
    case SPF_METRIC_MANHATTAN: return resolve_finish_t<SPF_METRIC_MANHATTAN>(c, s, a, csr);
    case SPF_METRIC_CHEBYSHEV: return resolve_finish_t<SPF_METRIC_CHEBYSHEV>(c, s, a, csr);
  }
  return fail(SPF_E_INVALID, "unknown metric %d", a.metric);
}

}  // namespace spf
