// scan_tc.cu — tcgen05 / TMEM / TMA candidate scan of the posting lists (sm_100a).
//
// Replaces the per-point distance loop of SpannIndex::find_k_nearest_neighbor_spann,
// src/spann/spann_index.rs:168-181 (reference), for batches large enough that a posting list is
// probed by many queries.  d(q, v) = |q|^2 - 2 q.v + |v|^2 is a dense contraction between the
// queries probing a list and the list's vectors: it runs on the 5th-gen tensor cores as one TF32
// pass (fp32 accumulation in TMEM) whose only job is to SELECT.  With E = tc_err_bound a certified
// bound on |d_tf32 - d_ref| (kernels.cuh) the result is the reference's, bit for bit:
//
//   bound pass   per (query, probe) pair the 16 (or 32) largest 32-column chunk maxima of
//                s = q'.v' - |v|^2/2.  Chunk maxima belong to distinct vectors, so the K-th largest
//                over a query's pairs, s_K, certifies that K probed vectors have
//                d_ref <= |q|^2 - 2 s_K + E.  A vector can be among the K smallest keys that pass
//                `dist <= threshold` (:176) only if d_tf32 - E <= min(thr_q, |q|^2 - 2 s_K + E).
//   candidates   either the bound pass also stores every chunk maximum (4 KB per 128 x 256 tile) and
//                the few (pair, 32-slot group) items whose maximum passes are re-evaluated exactly
//                ("group refinement": no second GEMM), or — centroid probe, rows longer than 256
//                floats, chunk maxima over the memory budget — a second GEMM pass emits the single
//                (query, slot) hits: typically K plus a handful per query.
//   select       exact distances (the reference's sequential f32 sum, :172), `dist <= threshold`,
//                the K smallest (distance, encounter index) keys — the stable sort + truncate of
//                :188-193.
//
// Queries without a certified bound (non-finite norms) or with more candidates than their bucket
// holds are flagged and re-run by the exact query-major kernel (search.cu), so the tensor path never
// decides a comparison on an approximate value.  The centroid probe is the same computation with the
// centroids as one posting list (K = nprobe <= 32), or, for larger nprobe, a dense store of s
// followed by a certified bitwise selection (probe_dense_select_kernel).
//
// Kernel anatomy (persistent, one CTA per SM, 320 threads), per unit = (list, <= 128 probing pairs):
//   warp 0     TMA producer: the unit's 128 gathered query rows (A, stationary for the unit) and a
//              4-stage ring of 256-slot x 32-float tiles of the list (B), SWIZZLE_128B, K-major,
//              plus the K-extension rows that carry -|v|^2/2
//   warp 1     TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=256, K=8 per MMA)
//   warps 2-9  epilogue: warp w reads TMEM lane quarter w%4 and column half (w-2)/4 of the
//              double-buffered 2 x 256 column accumulator, 32 columns per tcgen05.ld
#include <cub/cub.cuh>

#include "comm.cuh"
#include "tc_ptx.cuh"

namespace spf {

using namespace tc;

namespace {

constexpr int KB_MAX = 4;          // stationary A supports ld <= 128
constexpr int NSTAGE = 4;          // B ring depth
constexpr int A_KB_BYTES = BM * BK * 4;       // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 4;    // 32 KB
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_A = KB_MAX * A_KB_BYTES;                  // 64 KB
constexpr int SMEM_B = NSTAGE * B_STAGE_BYTES;               // 128 KB
constexpr int NEXT = 2;                                      // ring depth of the extension tiles
constexpr int AEXT_BYTES = BM * EXT_K * 4;                   // 4 KB, constant for the whole kernel
constexpr int BEXT_BYTES = BN * EXT_K * 4;                   // 8 KB per tile
constexpr int SMEM_AEXT_OFF = SMEM_A + SMEM_B;
constexpr int SMEM_BEXT_OFF = SMEM_AEXT_OFF + AEXT_BYTES;
constexpr int SMEM_BAR_OFF = SMEM_BEXT_OFF + NEXT * BEXT_BYTES;
constexpr int SMEM_TOTAL = SMEM_BAR_OFF + 256 + 1024;        // + alignment slack

constexpr int TOPR_MAX = 32;       // chunk maxima kept per (pair, column half) in phase A: 16 or 32, >= K
constexpr int UNIT_ROWS = BM;      // pairs per unit
constexpr uint32_t NOPAIR = 0xffffffffu;

struct UnitDesc { uint32_t slot0, nslots, tile0, pad; };   // tile0: first tile of the unit in the chunk-maxima array

struct ScanTcArgs {
  uint32_t nunits, kb, nprobe, cap;
  uint32_t u0;                     // rank of this launch's first unit (row scalars are kept for all units)
  uint32_t ubase;                  // first unit of this launch inside `desc` / the gathered rows (split launches)
  const UnitDesc* desc;            // per unit of this launch
  float4* cmax;                    // bound pass out (optional): per (tile, column half, row) its 4 chunk maxima
  uint32_t* cmask;                 // ... and per chunk 8 bits: which 4-slot subgroups can still pass (see below)
  const float2* rowes;             // per row: (E, slop) of its query — the running threshold of the masks
  uint32_t topk;                   // K of the search (<= TOPR)
  const float* rowthr;             // per row: threshold on s (phase B), -inf / +inf = enabled / disabled (phase A)
  const uint32_t* rowseq;          // per row: encounter base of the pair - slot0 (mod 2^32)
  const uint32_t* rowpair;         // per row: q * nprobe + p, NOPAIR for padding rows
  float* pairtop;                  // phase A out: (pair, half) x TOPR (16 or 32) chunk maxima, descending
  uint32_t* qcnt; uint2* bucket;   // phase B out: per query candidate count and (slot, encounter index) entries
  float* dense; uint32_t dense_ld; // MODE 2 out: s of row `pair` against every slot, dense_ld floats per row
};

// MODE 0: bound pass, 1: emit pass, 2: dense store of s (centroid probe with nprobe > 32)
template <int MODE, int TOPR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_e, ScanTcArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* smem_a = smem;
  unsigned char* smem_b = smem + SMEM_A;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BAR_OFF);
  uint64_t* a_full = bars;                 // [KB_MAX]
  uint64_t* a_empty = bars + KB_MAX;       // [KB_MAX]
  uint64_t* b_full = bars + 2 * KB_MAX;    // [NSTAGE]
  uint64_t* b_empty = b_full + NSTAGE;     // [NSTAGE]
  uint64_t* t_full = b_empty + NSTAGE;     // [2]
  uint64_t* t_empty = t_full + 2;          // [2]
  uint64_t* e_full = t_empty + 2;          // [NEXT]
  uint64_t* e_empty = e_full + NEXT;       // [NEXT]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(e_empty + NEXT);
  unsigned char* smem_aext = smem + SMEM_AEXT_OFF;
  unsigned char* smem_bext = smem + SMEM_BEXT_OFF;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_e) : "memory");
    for (int i = 0; i < KB_MAX; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], NUM_EPI_WARPS); }
    for (int i = 0; i < NEXT; ++i) { mbar_init(&e_full[i], 1); mbar_init(&e_empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // A side of the K extension: every row {1,1,0,0, 1,1,0,0}.  Both 16-byte halves are equal, so the
  // SWIZZLE_32B permutation leaves the tile unchanged and it can be written directly.
  for (int i = threadIdx.x; i < BM * 2; i += NUM_THREADS)
    reinterpret_cast<float4*>(smem_aext)[i] = make_float4(1.0f, 1.0f, 0.0f, 0.0f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes → visible to the MMA
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // ld <= 128: the unit's query tile (<= 4 K blocks) stays in shared memory for the whole unit.
  // Longer rows: the query K block is streamed with the list's K block, one ring stage for both
  // (the 64 KB that hold the stationary tile are then four 16 KB stages).
  const bool stream_a = a.kb > (uint32_t)KB_MAX;
  static_assert(KB_MAX == NSTAGE, "the stationary A tile and the streamed A stages share one region");

  if (warp == 0) {
    // =============================== TMA producer ===========================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, it = 0, ecount = 0;
      for (uint32_t u = a.ubase + blockIdx.x; u < a.ubase + a.nunits; u += gridDim.x) {
        const UnitDesc ud = a.desc[u];
        const uint32_t ntiles = (ud.nslots + BN - 1) / BN;
        if (ntiles == 0) continue;
        for (uint32_t t = 0; t < ntiles; ++t, ++ecount) {
          const int row0 = (int)(ud.slot0 + t * BN);
          for (uint32_t kb = 0; kb < a.kb; ++kb) {
            if (!stream_a && t == 0) {
              // the unit's query K block, requested as soon as the previous unit's last tile is done with
              // THIS block (not with all four): the first list stages of the new unit are already in
              // flight while the previous unit's last MMAs and epilogue run, so the ring never drains
              mbar_wait(&a_empty[kb], (it & 1) ^ 1);
              mbar_expect_tx(&a_full[kb], A_KB_BYTES);
              tma_load_2d(smem_a + kb * A_KB_BYTES, &map_a, &a_full[kb], (int)(kb * BK), (int)(u * UNIT_ROWS));
            }
            mbar_wait(&b_empty[stage], phase ^ 1);
            if (stream_a) {                               // the query K block rides in the same ring stage
              mbar_expect_tx(&b_full[stage], A_KB_BYTES + B_STAGE_BYTES);
              tma_load_2d(smem_a + stage * A_KB_BYTES, &map_a, &b_full[stage], (int)(kb * BK), (int)(u * UNIT_ROWS));
            } else {
              mbar_expect_tx(&b_full[stage], B_STAGE_BYTES);
            }
            tma_load_2d(smem_b + stage * B_STAGE_BYTES, &map_b, &b_full[stage], (int)(kb * BK), row0);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
          const uint32_t es = ecount % NEXT, eu = ecount / NEXT;
          mbar_wait(&e_empty[es], (eu & 1) ^ 1);
          mbar_expect_tx(&e_full[es], BEXT_BYTES);
          tma_load_2d(smem_bext + es * BEXT_BYTES, &map_e, &e_full[es], 0, row0);
        }
        ++it;
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ==============================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, it = 0, tcount = 0;
      for (uint32_t u = a.ubase + blockIdx.x; u < a.ubase + a.nunits; u += gridDim.x) {
        const UnitDesc ud = a.desc[u];
        const uint32_t ntiles = (ud.nslots + BN - 1) / BN;
        if (ntiles == 0) continue;
        for (uint32_t t = 0; t < ntiles; ++t, ++tcount) {
          const uint32_t buf = tcount & 1, use = tcount >> 1;
          mbar_wait(&t_empty[buf], (use & 1) ^ 1);        // epilogue drained this accumulator
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * BN;
          for (uint32_t kb = 0; kb < a.kb; ++kb) {
            if (!stream_a && t == 0) mbar_wait(&a_full[kb], it & 1);
            mbar_wait(&b_full[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem_a + (stream_a ? stage : kb) * A_KB_BYTES);
            const uint32_t b_addr = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              tc_mma_tf32(tmem_d, make_smem_desc(a_addr + k * UMMA_K * 4), make_smem_desc(b_addr + k * UMMA_K * 4),
                          IDESC_TF32, (kb | (uint32_t)k) != 0 ? 1u : 0u);
            }
            tc_commit(&b_empty[stage]);                     // frees the B stage once these MMAs retire
            if (!stream_a && t + 1 == ntiles) tc_commit(&a_empty[kb]);   // last use of this A block
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
          {   // K extension: accumulator += -|v|^2/2
            const uint32_t es = tcount % NEXT, eu = tcount / NEXT;
            mbar_wait(&e_full[es], eu & 1);
            tc_fence_after();
            tc_mma_tf32(tmem_d, make_smem_desc32(smem_u32(smem_aext)),
                        make_smem_desc32(smem_u32(smem_bext + es * BEXT_BYTES)), IDESC_TF32, 1u);
            tc_commit(&e_empty[es]);
          }
          tc_commit(&t_full[buf]);                        // accumulator complete → epilogue
        }
        ++it;
      }
    }
  } else {
    // =============================== epilogue warps ==========================================
    const uint32_t quarter = warp & 3;                    // TMEM lane quarter this warp may access
    const uint32_t half = (uint32_t)(warp - 2) >> 2;      // column half of every accumulator
    const uint32_t lrow = quarter * 32 + lane;
    const float INF = __int_as_float(0x7f800000);
    uint32_t tcount = 0;
    for (uint32_t u = a.ubase + blockIdx.x; u < a.ubase + a.nunits; u += gridDim.x) {
      const UnitDesc ud = a.desc[u];
      const uint32_t ntiles = (ud.nslots + BN - 1) / BN;
      if (ntiles == 0) continue;
      const size_t row = (size_t)(a.u0 + u) * UNIT_ROWS + lrow;
      const float thr_s = a.rowthr[row];
      const bool enabled = thr_s < INF;
      const bool warp_enabled = __any_sync(0xffffffffu, enabled);
      const uint32_t rseq = a.rowseq[row];
      const uint32_t pair = a.rowpair[row];
      const float2 es = (MODE == 0 && a.cmask != nullptr) ? a.rowes[row] : make_float2(0.f, 0.f);
      float top[TOPR];
#pragma unroll
      for (int i = 0; i < TOPR; ++i) top[i] = -INF;

      for (uint32_t t = 0; t < ntiles; ++t, ++tcount) {
        const uint32_t buf = tcount & 1, use = tcount >> 1;
        mbar_wait(&t_full[buf], use & 1);
        tc_fence_after();
        if (!warp_enabled) {                              // no live row in this lane quarter
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[buf]);
          continue;
        }
        const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + buf * BN + half * (BN / 2);
        const uint32_t colbase = t * BN + half * (BN / 2);   // first slot of this half tile, relative to slot0

        // One 32-column chunk of s.  Chunks are aligned with the 32-slot groups of the list, so a
        // chunk lies either inside the list or entirely behind its end (where the tile holds the
        // next list's vectors).
        // Subgroup masks (bound pass with kept chunk maxima): the final threshold on s of this row's query
        // is max(by_thr, kth - E) - slop with kth = the K-th largest chunk maximum over ALL of the query's
        // pairs (tau_kernel), which is at least the K-th largest this row has seen so far.  A 4-slot
        // subgroup whose maximum does not exceed that running bound (one more slop below, so that the
        // two evaluations need not round alike) cannot hold a passing element; the group refinement
        // then evaluates only the marked subgroups of a flagged chunk instead of all 32 vectors.
        float th_run = -INF;
        if (MODE == 0 && a.cmask != nullptr) {
          float kv = top[TOPR - 1];
#pragma unroll
          for (int i = 0; i < TOPR - 1; ++i)
            if ((uint32_t)(i + 1) == a.topk) kv = top[i];
          th_run = ((kv - es.x) - es.y) - es.y;
        }
        uint32_t mword = 0;
        auto process = [&](uint32_t (&rr)[32], int c) -> float {
          const uint32_t cb = colbase + (uint32_t)c * 32u;
          if (cb >= ud.nslots || !enabled) return -INF;
          if (MODE == 2) {                                // 128 contiguous bytes per thread: four full sectors
            float4* o = reinterpret_cast<float4*>(a.dense + (size_t)pair * a.dense_ld + cb);
#pragma unroll
            for (int g = 0; g < 8; ++g)
              o[g] = make_float4(__uint_as_float(rr[g * 4 + 0]), __uint_as_float(rr[g * 4 + 1]),
                                 __uint_as_float(rr[g * 4 + 2]), __uint_as_float(rr[g * 4 + 3]));
            return -INF;
          }
          float q[8];
#pragma unroll
          for (int g = 0; g < 8; ++g)
            q[g] = fmaxf(fmaxf(__uint_as_float(rr[g * 4 + 0]), __uint_as_float(rr[g * 4 + 1])),
                         fmaxf(__uint_as_float(rr[g * 4 + 2]), __uint_as_float(rr[g * 4 + 3])));
          float m = fmaxf(fmaxf(fmaxf(q[0], q[1]), fmaxf(q[2], q[3])), fmaxf(fmaxf(q[4], q[5]), fmaxf(q[6], q[7])));
          m = fmaxf(m, -INF);                             // an all-NaN chunk counts as empty
          if (MODE == 0 && a.cmask != nullptr) {
            uint32_t mk = 0;
#pragma unroll
            for (int g = 0; g < 8; ++g) mk |= q[g] > th_run ? (1u << g) : 0u;
            mword |= mk << (8 * c);
          }
          if (MODE == 1) {
            if (m > thr_s) {
              // which of the 32 columns pass: straight-line mask, then one iteration per hit (only the
              // column index is needed, so the registers are never indexed dynamically)
              uint32_t hits = 0;
#pragma unroll
              for (int e = 0; e < 32; ++e) hits |= __uint_as_float(rr[e]) > thr_s ? (1u << e) : 0u;
              const uint32_t qi = pair / a.nprobe;
              while (hits) {
                const uint32_t slot = ud.slot0 + cb + (uint32_t)(__ffs(hits) - 1);
                hits &= hits - 1;
                const uint32_t pos = atomicAdd(a.qcnt + qi, 1u);
                if (pos < a.cap) a.bucket[(size_t)qi * a.cap + pos] = make_uint2(slot, rseq + slot);
              }
            }
          } else {
            float v = m;
#pragma unroll
            for (int i = 0; i < TOPR; ++i) {
              const float hi = fmaxf(top[i], v);
              v = fminf(top[i], v);
              top[i] = hi;
            }
          }
          return m;
        };

        // software pipeline over the 4 chunks of this warp's column half: the TMEM load of chunk
        // c+1 is in flight while chunk c is processed
        uint32_t ra[32], rbuf[32];
        tc_ld32_issue(taddr, ra);
        tc_ld32_wait(ra);
        tc_ld32_issue(taddr + 32, rbuf);
        const float m0 = process(ra, 0);
        tc_ld32_wait(rbuf);
        tc_ld32_issue(taddr + 64, ra);
        const float m1 = process(rbuf, 1);
        tc_ld32_wait(ra);
        tc_ld32_issue(taddr + 96, rbuf);
        const float m2 = process(ra, 2);
        tc_ld32_wait(rbuf);
        tc_fence_before();                                // this warp's part of the accumulator is read
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[buf]);
        const float m3 = process(rbuf, 3);
        if (MODE == 0 && a.cmax != nullptr)                   // one coalesced 512-byte store per warp and tile
          a.cmax[((size_t)(ud.tile0 + t) * 2 + half) * UNIT_ROWS + lrow] = make_float4(m0, m1, m2, m3);
        if (MODE == 0 && a.cmask != nullptr)
          a.cmask[((size_t)(ud.tile0 + t) * 2 + half) * UNIT_ROWS + lrow] = mword;
      }
      if (MODE == 0 && enabled) {
        float4* o = reinterpret_cast<float4*>(a.pairtop + ((size_t)pair * 2 + half) * TOPR);
#pragma unroll
        for (int i = 0; i < TOPR / 4; ++i) o[i] = make_float4(top[4 * i], top[4 * i + 1], top[4 * i + 2], top[4 * i + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// index side: slot layout → row-major TF32 rows + K-extension rows + norms
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float rna_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

// One warp per 32-slot group, lane = slot.
__global__ void slot_prep_kernel(const float* __restrict__ vecs, const uint64_t* __restrict__ slot_ids, uint64_t ngroups,
                                 uint32_t ld4, float* __restrict__ vtf, float* __restrict__ vext,
                                 float* __restrict__ vnorm, float* __restrict__ vres) {
  const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= ngroups) return;
  const float4* V4 = reinterpret_cast<const float4*>(vecs) + g * ld4 * 32 + lane;
  const uint64_t slot = g * 32 + lane;
  float4* o = reinterpret_cast<float4*>(vtf) + slot * ld4;
  float acc = 0.f, racc = 0.f;
  for (uint32_t c = 0; c < ld4; ++c) {
    const float4 v = __ldg(V4 + (size_t)c * 32);
    const float4 t = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
    o[c] = t;
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc);
    acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    const float ex = v.x - t.x, ey = v.y - t.y, ez = v.z - t.z, ew = v.w - t.w;   // exact (Sterbenz)
    racc = fmaf(ex, ex, racc); racc = fmaf(ey, ey, racc);
    racc = fmaf(ez, ez, racc); racc = fmaf(ew, ew, racc);
  }
  const bool valid = slot_ids[slot] != ~0ull;
  vnorm[slot] = valid ? acc : 0.f;
  vres[slot] = valid ? sqrtf(racc) : 0.f;
  float h = -__int_as_float(0x7f800000), m = 0.f, l = 0.f;
  if (valid) {
    const float v = -0.5f * acc;               // exact (power of two)
    h = rna_tf32(v);
    const float r1 = v - h;                    // exact: |r1| <= 2^-11 |v|
    m = rna_tf32(r1);
    l = rna_tf32(r1 - m);
  }
  float4* e = reinterpret_cast<float4*>(vext + slot * 8);
  e[0] = make_float4(h, m, 0.f, 0.f);
  e[1] = make_float4(l, 0.f, 0.f, 0.f);
}

// out2[0] = max(a), out2[1] = max(b) over n values >= 0; out2 must be zeroed.  Non-negative floats
// order like their bit patterns, so an integer atomicMax does it.
__global__ void max2_atomic_kernel(const float* __restrict__ a, const float* __restrict__ b, uint64_t n, float* out2) {
  float va = 0.f, vb = 0.f;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    va = fmaxf(va, a[i]);
    vb = fmaxf(vb, b[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    va = fmaxf(va, __shfl_xor_sync(0xffffffffu, va, o));
    vb = fmaxf(vb, __shfl_xor_sync(0xffffffffu, vb, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(reinterpret_cast<unsigned int*>(out2), __float_as_uint(va));
    atomicMax(reinterpret_cast<unsigned int*>(out2) + 1, __float_as_uint(vb));
  }
}

// ---------------------------------------------------------------------------------------------
// per call: units, gathered query rows, bounds, refinement
// ---------------------------------------------------------------------------------------------
// units per list = ceil(pairs probing it / 128), 0 for lists this rank does not hold
// totals[0] += units, totals[1] += 256-slot tiles over all units, totals[2] = max tiles of a probed list,
// totals[3] / [4] += units / tiles of the lists with more than one unit:
// everything the host needs for its sizing decisions comes back in ONE device -> host copy
__global__ void tc_unit_counts_kernel(const uint32_t* __restrict__ list_off, const uint64_t* __restrict__ grp_off,
                                      uint32_t nlists, uint32_t* __restrict__ counts, unsigned long long* __restrict__ totals) {
  const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l > nlists) return;
  uint32_t n = 0;
  if (l < nlists && grp_off[l + 1] != grp_off[l]) n = (list_off[l + 1] - list_off[l] + UNIT_ROWS - 1) / UNIT_ROWS;
  counts[l] = n;
  if (n) {
    const unsigned long long tiles = ((unsigned long long)(grp_off[l + 1] - grp_off[l]) * 32ull + BN - 1) / BN;
    atomicAdd(&totals[0], (unsigned long long)n);
    atomicAdd(&totals[1], (unsigned long long)n * tiles);
    atomicMax(&totals[2], tiles);
    if (n > 1) {
      atomicAdd(&totals[3], (unsigned long long)n);
      atomicAdd(&totals[4], (unsigned long long)n * tiles);
    }
  }
}

// Sort key of a unit rank: bit 31 = its list has a single unit (at most 128 probing pairs), low bits =
// 0x7fffffff - 32-slot groups of the list.  Ascending order = units of multi-unit lists first, longest
// list first within each class.
__host__ __device__ __forceinline__ uint32_t unit_key(uint32_t groups, bool single_unit) {
  return (single_unit ? 0x80000000u : 0u) | (0x7fffffffu - (groups < 0x7fffffffu ? groups : 0x7fffffffu));
}
__host__ __device__ __forceinline__ uint32_t unit_key_groups(uint32_t key) { return 0x7fffffffu - (key & 0x7fffffffu); }

// the centroid probe: ONE list probed by every query, so the units are known on the host
__global__ void tc_probe_units_kernel(const uint64_t* __restrict__ grp_off, uint32_t nunits, uint32_t* __restrict__ unit_off,
                                      uint32_t* __restrict__ keys_sorted, uint32_t* __restrict__ order) {
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u == 0) { unit_off[0] = 0; unit_off[1] = nunits; }
  if (u >= nunits) return;
  keys_sorted[u] = unit_key((uint32_t)(grp_off[1] - grp_off[0]), false);
  order[u] = u;
}

// Sort key of a unit: the units of lists probed by more than 128 pairs first, then longest list first
// (stable, so the units of one list stay adjacent and share the list through L2).  Dealing the sorted
// units round-robin to the persistent CTAs balances them.
__global__ void tc_unit_keys_kernel(const uint32_t* __restrict__ unit_off, const uint64_t* __restrict__ grp_off,
                                    uint32_t nlists, uint32_t nunits, uint32_t* __restrict__ keys,
                                    uint32_t* __restrict__ vals) {
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nunits) return;
  uint32_t lo = 0, hi = nlists;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (unit_off[mid] <= u) lo = mid; else hi = mid;
  }
  keys[u] = unit_key((uint32_t)(grp_off[lo + 1] - grp_off[lo]), unit_off[lo + 1] - unit_off[lo] < 2u);
  vals[u] = u;
}

struct GatherArgs {
  const uint32_t* pair_sorted; const uint32_t* list_off; const uint32_t* unit_off; uint32_t nlists;
  const uint64_t* grp_off; const uint32_t* lens; const uint32_t* seqbase;
  uint32_t nprobe, ld4, d;
  uint32_t u0;                     // first unit (rank in `order`) of this launch
  const uint32_t* order;           // unit ids, longest list first
  const uint32_t* tile_off;        // first chunk-maxima tile of every unit rank, or NULL
  const float* qtf;                // nq x ld rounded queries
  const float* qthr;               // nq thresholds on s (phase B) or NULL (phase A)
  uint32_t tau_probes;             // phase A: pairs with p < tau_probes take part
  int copy_rows;                   // 0: the gathered rows of this chunk are already in place
  float* A; UnitDesc* desc; float* rowthr; uint32_t* rowseq; uint32_t* rowpair;
  uint32_t* unit_slot0;            // per unit rank: first slot of its list
  float2* rowes; const float* qnorm; const float* qres; const float* vstat;   // per row (E, slop) of its query, or NULL
  unsigned long long* bytes;       // algorithmic scan bytes (counted when != NULL)
  unsigned long long* stream_bytes;   // [0] bytes of list tiles one pass requests, [1] the same counting every list once
};

// One CTA (8 warps) per unit: unit descriptor, the unit's query rows (TF32 copies) gathered into
// a contiguous 128 x ld tile for TMA, and the per-row scalars of the epilogue.
__global__ void __launch_bounds__(256) unit_gather_kernel(GatherArgs g) {
  const uint32_t u = g.order[g.u0 + blockIdx.x];
  uint32_t lo = 0, hi = g.nlists;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (g.unit_off[mid] <= u) lo = mid; else hi = mid;
  }
  const uint32_t l = lo;
  const uint32_t batch = u - g.unit_off[l];
  const uint32_t p0 = g.list_off[l] + batch * UNIT_ROWS;
  const uint32_t nb = min((uint32_t)UNIT_ROWS, g.list_off[l + 1] - p0);
  const uint32_t slot0 = (uint32_t)(g.grp_off[l] * 32);
  const float INF = __int_as_float(0x7f800000);
  if (threadIdx.x == 0) {
    UnitDesc ud;
    ud.slot0 = slot0;
    ud.nslots = (uint32_t)((g.grp_off[l + 1] - g.grp_off[l]) * 32);
    ud.tile0 = g.tile_off ? g.tile_off[g.u0 + blockIdx.x] : 0u;
    ud.pad = 0;
    g.desc[blockIdx.x] = ud;
    g.unit_slot0[g.u0 + blockIdx.x] = slot0;
    if (g.bytes) {
      atomicAdd(g.bytes, (unsigned long long)nb * g.lens[l] * g.d * 4ull);
      atomicAdd(g.stream_bytes, (unsigned long long)ud.nslots * (g.ld4 * 16ull + EXT_K * 4ull));
      if (batch == 0) atomicAdd(g.stream_bytes + 1, (unsigned long long)ud.nslots * (g.ld4 * 16ull + EXT_K * 4ull));
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t r = warp; r < (uint32_t)UNIT_ROWS; r += 8) {
    const size_t row = (size_t)blockIdx.x * UNIT_ROWS + r;        // in the gathered tile buffer of this launch
    const size_t grow = (size_t)(g.u0 + blockIdx.x) * UNIT_ROWS + r;  // in the per-row scalars (all units)
    uint32_t pair = NOPAIR, q = 0;
    if (r < nb) { pair = g.pair_sorted[p0 + r]; q = pair / g.nprobe; }
    // Padding rows are left as they are: a row of the GEMM only feeds its own accumulator lane,
    // and the epilogue never looks at a row whose threshold is +inf.
    if (g.copy_rows && pair != NOPAIR) {
      float4* dst = reinterpret_cast<float4*>(g.A) + row * g.ld4;
      const float4* src = reinterpret_cast<const float4*>(g.qtf) + (size_t)q * g.ld4;
      for (uint32_t c = lane; c < g.ld4; c += 32) dst[c] = __ldg(src + c);
    }
    if (lane == 0) {
      float th = INF;
      if (pair != NOPAIR) {
        if (g.qthr) th = g.qthr[q];
        else th = (pair - q * g.nprobe) < g.tau_probes ? -INF : INF;
      }
      g.rowthr[grow] = th;
      g.rowseq[grow] = pair != NOPAIR ? g.seqbase[pair] - slot0 : 0u;
      g.rowpair[grow] = pair;
      if (g.rowes) {
        float2 es = make_float2(INF, INF);
        if (pair != NOPAIR) {
          const float qn = g.qnorm[q], cnmax = g.vstat[0];
          es.x = tc_err_bound(qn, g.qres[q], cnmax, g.vstat[1], g.ld4 * 4);
          es.y = 1e-6f * (qn + cnmax) + 1e-30f;
        }
        g.rowes[grow] = es;
      }
    }
  }
}

struct TauArgs {
  uint64_t nq; uint32_t nprobe, tau_probes, K, ld, topr;
  const uint32_t* probe; const uint64_t* grp_off;
  const float* pairtop; const float* qnorm; const float* qres; const float* thr; const float* vstat;
  float* qthr; float* qbound; uint8_t* qflag;
};

// Sorted insertion of one value into a warp-distributed descending list of 32 floats (lane = rank).
__device__ __forceinline__ void top_insert_desc(float& val, float v, int lane) {
  const int pos = __popc(__ballot_sync(0xffffffffu, val >= v));
  const float up = __shfl_up_sync(0xffffffffu, val, 1);
  if (lane > pos) val = up;
  else if (lane == pos) val = v;
}

// One warp per query: s_K = K-th largest chunk maximum over the query's pairs, then the phase-B
// threshold on s.  A candidate can matter only if d_ref <= B = min(thr_q, |q|^2 - 2 s_K + E), and
// d_ref >= d_tf32 - E = |q|^2 - 2 s - E, i.e. only if  s >= (|q|^2 - E - B) / 2.
__global__ void __launch_bounds__(256) tau_kernel(TauArgs a) {
  const int lane = threadIdx.x & 31;
  const uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= a.nq) return;
  const float INF = __int_as_float(0x7f800000);
  float val = -INF, kth = -INF;
  const uint32_t np = min(a.nprobe, a.tau_probes);
  for (uint32_t p = 0; p < np; ++p) {
    const uint32_t pair = (uint32_t)(q * a.nprobe + p);
    const uint32_t l = a.probe[pair];
    if (a.grp_off[l + 1] == a.grp_off[l]) continue;       // not held here: no unit ran
    // the pair's two arrays (column halves) of topr descending values, 32 values per load
    for (uint32_t base = 0; base < 2 * a.topr; base += 32) {
      const float mine = a.pairtop[(size_t)pair * 2 * a.topr + base + lane];
      for (uint32_t a0 = 0; a0 < 32; a0 += a.topr) {       // the arrays inside this load (topr = 16: two)
        for (uint32_t j = 0; j < a.topr && j < 32; ++j) {
          const float v = __shfl_sync(0xffffffffu, mine, (int)(a0 + j));
          if (!(v > kth)) break;                            // each array is descending
          top_insert_desc(val, v, lane);
          kth = __shfl_sync(0xffffffffu, val, (int)a.K - 1);
        }
      }
    }
  }
  if (lane == 0) {
    const float qn = a.qnorm[q];
    const float cnmax = a.vstat[0], dcmax = a.vstat[1];
    const float E = tc_err_bound(qn, a.qres[q], cnmax, dcmax, a.ld);
    const bool hopeless = !(E < INF) || !(qn < INF);
    const float slop = 1e-6f * (qn + cnmax) + 1e-30f;
    const float by_thr = 0.5f * ((qn - E) - a.thr[q]);    // -inf when pruning is off; NaN thr: below
    const float by_tau = kth - E;                          // = (|q|^2 - E - (|q|^2 - 2 s_K + E)) / 2
    float th = fmaxf(by_thr, by_tau) - slop;
    if (!(a.thr[q] == a.thr[q])) th = INF;                 // NaN threshold: `dist <= thr` never holds
    a.qthr[q] = hopeless ? INF : th;
    // exact-side filter of the group refinement: the K-th smallest probed distance is <= this
    a.qbound[q] = (fmaf(-2.0f, kth, qn) + E) + slop;
    a.qflag[q] = hopeless ? 1 : 0;
  }
}

// List-sharded search: qbound (this rank's certified bound on the query's K-th distance) is
// min-reduced over the ranks; the flagging threshold then follows from the global bound B exactly
// like the pruning threshold does: a candidate matters only if d_ref <= B, and d_ref >= |q|^2 - 2 s - E
// (E of THIS rank's vectors), i.e. only if s >= (|q|^2 - E - B) / 2.
__global__ void bound_sanitize_kernel(float* __restrict__ qbound, uint64_t nq) {
  const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nq && !(qbound[q] == qbound[q])) qbound[q] = __int_as_float(0x7f800000);   // NaN never enters the reduction
}
__global__ void bound_tighten_kernel(const float* __restrict__ qbound, const float* __restrict__ qnorm,
                                     const float* __restrict__ qres, const float* __restrict__ vstat, uint32_t ld,
                                     uint64_t nq, float* __restrict__ qthr) {
  const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const float INF = __int_as_float(0x7f800000);
  const float th = qthr[q];
  if (!(th < INF)) return;                                 // hopeless query or NaN pruning threshold: stays closed
  const float qn = qnorm[q];
  const float cnmax = vstat[0], dcmax = vstat[1];
  const float E = tc_err_bound(qn, qres[q], cnmax, dcmax, ld);
  const float slop = 1e-6f * (qn + cnmax) + 1e-30f;
  const float by_global = 0.5f * ((qn - E) - qbound[q]) - slop;
  if (by_global > th) qthr[q] = by_global;
}

// Tiles of every unit rank (from its sort key).
__global__ void tc_unit_tiles_kernel(const uint32_t* __restrict__ keys_sorted, uint32_t nunits, uint32_t* __restrict__ ntiles) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > nunits) return;
  ntiles[r] = r < nunits ? (unit_key_groups(keys_sorted[r]) * 32u + BN - 1) / BN : 0u;
}

struct WorkItem { uint32_t pair, group; };   // group = 4-slot subgroup of the index (32-slot group * 8 + subgroup)

struct FlagArgs {
  const uint32_t* keys_sorted;     // unit ranks: unit_key() of the list
  const uint32_t* tile_off; const float4* cmax; const uint32_t* cmask;
  const uint32_t* rowpair; const uint32_t* unit_slot0;  // per row: pair; per unit rank: first slot of its list
  const float* qthr; uint32_t nprobe;
  WorkItem* work; uint32_t work_cap; uint32_t* nwork; uint8_t* qflag;
};

// One CTA per unit rank and run of FLAG_TILES tiles (the longest list would otherwise serialise
// hundreds of dependent loads in one CTA), thread = (column half, row): walks the chunk maxima and queues
// the marked 4-slot subgroups of every (pair, 32-slot group) whose maximum passes the pair's
// threshold on s.  Items that do not fit flag their query for the exact fallback.
constexpr uint32_t FLAG_TILES = 16;
__global__ void __launch_bounds__(256) chunk_flag_kernel(FlagArgs f) {
  // items are staged in shared memory and appended with one global atomic per CTA (a single
  // global counter bumped once per item serialises in L2)
  constexpr uint32_t SQ = 2048;
  __shared__ WorkItem sq[SQ];
  __shared__ uint32_t sn, sbase;
  const uint32_t r = blockIdx.x;
  {
    const uint32_t nt = (unit_key_groups(f.keys_sorted[r]) * 32u + BN - 1) / BN;
    if (blockIdx.y * FLAG_TILES >= nt) return;            // block-uniform
  }
  if (threadIdx.x == 0) sn = 0;
  __syncthreads();
  const uint32_t half = threadIdx.x >> 7, lrow = threadIdx.x & 127;
  const size_t row = (size_t)r * UNIT_ROWS + lrow;
  const uint32_t pair = f.rowpair[row];
  if (pair != NOPAIR) {
    const uint32_t q = pair / f.nprobe;
    const float th = f.qthr[q];
    const uint32_t groups = unit_key_groups(f.keys_sorted[r]);
    const uint32_t ntiles = (groups * 32u + BN - 1) / BN;
    const uint32_t g0 = f.unit_slot0[r] >> 5;
    const size_t e0 = ((size_t)f.tile_off[r] * 2 + half) * UNIT_ROWS + lrow;
    const float4* cm = f.cmax + e0;
    const uint32_t* mk = f.cmask + e0;
    const uint32_t t_end = min(ntiles, (blockIdx.y + 1) * FLAG_TILES);
#pragma unroll 4
    for (uint32_t t = blockIdx.y * FLAG_TILES; t < t_end; ++t) {
      const float4 m = cm[(size_t)t * 2 * UNIT_ROWS];
      const uint32_t mw = mk[(size_t)t * 2 * UNIT_ROWS];
      const float mv[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t grp = t * 8 + half * 4 + c;          // chunk = 32-slot group of the list
        if (grp < groups && mv[c] > th) {
          uint32_t sub = (mw >> (8 * c)) & 0xffu;           // never empty: the chunk maximum is in a marked subgroup
          while (sub) {
            WorkItem w;
            w.pair = pair; w.group = (g0 + grp) * 8u + (uint32_t)(__ffs(sub) - 1);
            sub &= sub - 1;
            const uint32_t sp = atomicAdd(&sn, 1u);
            if (sp < SQ) {
              sq[sp] = w;
            } else {                                          // staging full: straight to the global list
              const uint32_t pos = atomicAdd(f.nwork, 1u);
              if (pos < f.work_cap) f.work[pos] = w; else f.qflag[q] = 1;
            }
          }
        }
      }
    }
  }
  __syncthreads();
  const uint32_t n = min(sn, SQ);
  if (threadIdx.x == 0 && n) sbase = atomicAdd(f.nwork, n);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t pos = sbase + i;
    if (pos < f.work_cap) f.work[pos] = sq[i]; else f.qflag[sq[i].pair / f.nprobe] = 1;
  }
}

struct GroupArgs {
  ScanArgs s; const WorkItem* work; const uint32_t* nwork; uint32_t work_cap;
  const float* qbound; uint32_t cap; uint32_t* qcnt; uint4* ebucket;   // exact entries: key lo, key hi, slot
};

// Eight queued (pair, 4-slot subgroup) items per warp and step, four lanes each: the lane's vector
// against the pair's query, exact (the reference's sequential f32 sum, :172); vectors passing
// `dist <= threshold` (:176) and the certified bound on the K-th smallest distance go to the query's
// bucket with their (distance, encounter) key.
__global__ void __launch_bounds__(256) group_refine_kernel(GroupArgs g) {
  const ScanArgs& a = g.s;
  const int lane = threadIdx.x & 31;
  const uint32_t nw = min(*g.nwork, g.work_cap);
  const uint32_t ld4 = a.ld / 4;
  const float4* V4 = reinterpret_cast<const float4*>(a.vecs);
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t w0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 8u; w0 < nw; w0 += nwarps * 8u) {
    const uint32_t w = w0 + ((uint32_t)lane >> 2);
    if (w >= nw) continue;                                 // no warp-wide operation below
    const WorkItem it = g.work[w];
    const uint32_t q = it.pair / a.nprobe;
    const uint32_t slot = it.group * 4u + ((uint32_t)lane & 3u);
    const float4* Q4 = reinterpret_cast<const float4*>(a.Q) + (size_t)q * ld4;
    const float4* base = V4 + (size_t)(slot >> 5) * ld4 * 32 + (slot & 31u);
    // The sum only grows (or turns NaN), so a lane whose partial sum is already above the bound can
    // stop reading: it could not pass the tests below anyway.
    const float thr = a.thr[q], qb = g.qbound[q];
    const float bound = fminf(thr, qb);
    float acc = 0.0f;
    for (uint32_t c0 = 0; c0 < ld4; c0 += 4) {
      if (!(acc <= bound)) break;
#pragma unroll
      for (uint32_t u = 0; u < 4; ++u) {
        const uint32_t c = c0 + u;
        if (c < ld4) {
          const float4 v = __ldg(base + (size_t)c * 32);
          const float4 qv = __ldg(Q4 + c);
          acc = dist_step<SPF_METRIC_EUCLIDEAN>(acc, qv.x, v.x);
          acc = dist_step<SPF_METRIC_EUCLIDEAN>(acc, qv.y, v.y);
          acc = dist_step<SPF_METRIC_EUCLIDEAN>(acc, qv.z, v.z);
          acc = dist_step<SPF_METRIC_EUCLIDEAN>(acc, qv.w, v.w);
        }
      }
    }
    const bool valid = a.slot_ids[slot] != ~0ull;
    if (valid && acc <= thr && acc <= qb) {
      const uint32_t l = a.probe[it.pair];
      // encounter index: base of this probe + position in the list
      const uint32_t seq = a.seqbase[it.pair] + (slot - (uint32_t)(a.grp_off[l] * 32));
      const uint32_t pos = atomicAdd(g.qcnt + q, 1u);
      if (pos < g.cap) g.ebucket[(size_t)q * g.cap + pos] = make_uint4(seq, __float_as_uint(acc), slot, 0u);
    }
  }
}

struct RefineArgs {
  ScanArgs s; uint64_t nq; uint32_t cap;
  const uint32_t* qcnt; const uint2* bucket; uint8_t* qflag;
  const uint4* ebucket;            // EXACT: entries already hold the exact key (group refinement)
  unsigned long long* stats;       // [0] candidates emitted, [1] queries handed to the exact fallback
};

// One warp per query: exact distance of every emitted (query, slot) pair — the reference's
// sequential f32 sum — then `<= thr` and the K smallest (distance, encounter index) keys.
template <bool EXACT>
__global__ void __launch_bounds__(256) refine_kernel(RefineArgs r) {
  const ScanArgs& a = r.s;
  const int lane = threadIdx.x & 31;
  const uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= r.nq) return;
  const uint32_t n = r.qcnt[q];
  if (lane == 0) atomicAdd(r.stats, (unsigned long long)n);
  if (r.qflag[q] || n > r.cap) {                 // the exact query-major kernel owns this query
    if (lane == 0) { r.qflag[q] = 1; atomicAdd(r.stats + 1, 1ull); }
    return;
  }
  const uint32_t ld4 = a.ld / 4;
  const float4* V4 = reinterpret_cast<const float4*>(a.vecs);
  const float4* Q4 = reinterpret_cast<const float4*>(a.Q) + q * ld4;
  const float thr = a.thr[q];
  unsigned long long key = ~0ull, pay = ~0ull, kth = ~0ull;
  for (uint32_t b = 0; b < n; b += 32) {
    const uint32_t i = b + lane;
    const bool have = i < n;
    uint2 ent = make_uint2(0u, 0u);               // slot, encounter index
    float acc = 0.0f;
    if (EXACT) {
      if (have) {
        const uint4 e = r.ebucket[q * r.cap + i];
        ent = make_uint2(e.z, e.x);
        acc = __uint_as_float(e.y);
      }
    } else {
      if (have) ent = r.bucket[q * r.cap + i];
      const float4* base = V4 + ((size_t)(ent.x >> 5) * ld4) * 32 + (ent.x & 31u);
      if (have) {
#pragma unroll 4
        for (uint32_t c = 0; c < ld4; ++c) {
          const float4 v = __ldg(base + (size_t)c * 32);
          const float4 qv = __ldg(Q4 + c);
          acc = dist_step<SPF_METRIC_EUCLIDEAN>(acc, qv.x, v.x);
          acc = dist_step<SPF_METRIC_EUCLIDEAN>(acc, qv.y, v.y);
          acc = dist_step<SPF_METRIC_EUCLIDEAN>(acc, qv.z, v.z);
          acc = dist_step<SPF_METRIC_EUCLIDEAN>(acc, qv.w, v.w);
        }
      }
    }
    const unsigned long long ck = ((unsigned long long)__float_as_uint(acc) << 32) | (unsigned long long)ent.y;
    unsigned bal = __ballot_sync(0xffffffffu, have && acc <= thr && ck < kth);
    while (bal) {
      const int src = __ffs(bal) - 1;
      bal &= bal - 1;
      const unsigned long long k2 = __shfl_sync(0xffffffffu, ck, src);
      const unsigned long long p2 = __shfl_sync(0xffffffffu, (unsigned long long)ent.x, src);
      if (k2 < kth) {
        const int pos = __popc(__ballot_sync(0xffffffffu, key < k2));
        const unsigned long long uk = __shfl_up_sync(0xffffffffu, key, 1);
        const unsigned long long up = __shfl_up_sync(0xffffffffu, pay, 1);
        if (lane > pos) { key = uk; pay = up; }
        else if (lane == pos) { key = k2; pay = p2; }
        kth = __shfl_sync(0xffffffffu, key, (int)a.K - 1);
      }
    }
  }
  const bool ok = (uint32_t)lane < a.K && key != ~0ull;
  const uint32_t count = __popc(__ballot_sync(0xffffffffu, ok));
  if ((uint32_t)lane < a.K) {
    a.out_ids[q * a.K + lane] = ok ? a.slot_ids[pay] : ~0ull;
    a.out_dists[q * a.K + lane] = ok ? __uint_as_float((uint32_t)(key >> 32)) : __int_as_float(0x7f800000);
    a.out_keys[q * a.K + lane] = key;
    a.out_slots[q * a.K + lane] = ok ? pay : ~0ull;
  }
  if (lane == 0) a.out_counts[q] = count;
}

// ---------------------------------------------------------------------------------------------
// dense centroid probe (nprobe > 32): s of every query against every centroid, then selection
// ---------------------------------------------------------------------------------------------
// Per-unit / per-row scalars of a dense launch over the queries [0, nq) of one chunk: unit u holds the
// queries u*128 .. u*128+127, every unit scans the whole centroid list.
__global__ void dense_setup_kernel(uint32_t nunits, uint32_t nq, uint32_t nslots, UnitDesc* __restrict__ desc,
                                   float* __restrict__ rowthr, uint32_t* __restrict__ rowseq,
                                   uint32_t* __restrict__ rowpair) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nunits * UNIT_ROWS) return;
  if ((r % UNIT_ROWS) == 0) { UnitDesc ud; ud.slot0 = 0; ud.nslots = nslots; ud.tile0 = 0; ud.pad = 0; desc[r / UNIT_ROWS] = ud; }
  rowthr[r] = r < nq ? -__int_as_float(0x7f800000) : __int_as_float(0x7f800000);
  rowseq[r] = 0;
  rowpair[r] = r < nq ? r : NOPAIR;
}

struct DenseSelArgs {
  const float* S; uint32_t s_ld;                 // nq x s_ld approximate s = q'.c' - |c|^2/2
  const float* Q; const float* C; uint32_t ld;   // exact queries (chunk) and centroids, row-major
  const float* qnorm; const float* qres; const float* vstat;
  uint32_t nlists, nprobe; float prune_factor;
  const uint32_t* lens; uint32_t* probe; float* thr; uint32_t* seqbase; int* redo;
};

// One CTA per query.  With s_T the nprobe-th largest approximate s, nprobe centroids have
// d_ref <= |q|^2 - 2 s_T + E, so every centroid among the nprobe nearest has s >= s_T - E: those
// (nprobe plus a few) are evaluated exactly (the reference's sequential f32 sum), sorted by
// (distance bits, list id) and the first nprobe are the probe — kiddo's nearest_n (:164).  Queries
// without a certified bound, with a non-finite s, or with more candidates than 256 * OUT raise `redo`.
template <int ITEMS, int OUT>
__global__ void __launch_bounds__(256) probe_dense_select_kernel(DenseSelArgs a) {
  typedef cub::BlockRadixSort<unsigned long long, 256, OUT> Sort;
  typedef cub::BlockScan<uint32_t, 256> Scan;
  __shared__ union { typename Sort::TempStorage sort; typename Scan::TempStorage scan; } tmp;
  __shared__ uint32_t s_sel[256 * OUT];
  __shared__ uint32_t s_cnt[2][8];
  const uint64_t q = blockIdx.x;
  const float* row = a.S + q * a.s_ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float INF = __int_as_float(0x7f800000);
  // keys ascending with DEscending s: bits flipped into unsigned order, then inverted
  auto to_key = [](float s) {
    const uint32_t b = __float_as_uint(s);
    return ~((b & 0x80000000u) ? ~b : (b | 0x80000000u));
  };
  auto block_sum = [&](uint32_t c, int slot) {
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) s_cnt[slot][warp] = c;
    __syncthreads();
    uint32_t tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s_cnt[slot][w];
    return tot;
  };
  uint32_t dv[ITEMS];
  uint32_t bad = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint32_t j = i * 256 + threadIdx.x;
    const float s = j < a.nlists ? row[j] : -INF;
    bad |= (j < a.nlists && !(fabsf(s) < INF)) ? 1u : 0u;      // NaN or infinite
    dv[i] = j < a.nlists ? to_key(s) : ~0u;
  }
  const float qn = a.qnorm[q], cnmax = a.vstat[0];
  const float E = tc_err_bound(qn, a.qres[q], cnmax, a.vstat[1], a.ld);
  if (!(E < INF) || !(qn < INF)) bad = 1;
  if (block_sum(bad, 0) != 0) {                   // uniform: the exact probe owns such batches
    if (threadIdx.x == 0) *a.redo = 1;
    return;
  }
  // nprobe-th smallest key, most significant bit first (see probe_topn_kernel, search.cu)
  uint32_t prefix = 0, want = a.nprobe;
  for (int bit = 31; bit >= 0; --bit) {
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) c += ((dv[i] ^ prefix) >> bit) == 0u ? 1u : 0u;
    const uint32_t tot = block_sum(c, bit & 1);
    if (want > tot) { want -= tot; prefix |= 1u << bit; }
  }
  // s_T back from the key; candidates: s >= s_T - E (E in distance units is 2E on d = |q|^2 - 2 s)
  const uint32_t fb = ~prefix;
  const float s_t = __uint_as_float((fb & 0x80000000u) ? (fb & 0x7fffffffu) : ~fb);
  const float slop = 1e-6f * (qn + cnmax) + 1e-30f;
  const uint32_t cut = to_key((s_t - E) - slop);  // keys <= cut are candidates
  uint32_t mine = 0, base0 = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) mine += dv[i] <= cut ? 1u : 0u;
  __syncthreads();
  Scan(tmp.scan).ExclusiveSum(mine, base0);
  const uint32_t total = block_sum(mine, 0);
  if (total > 256u * OUT) {
    if (threadIdx.x == 0) *a.redo = 1;
    return;
  }
#pragma unroll
  for (int i = 0; i < ITEMS; ++i)
    if (dv[i] <= cut) s_sel[base0++] = i * 256 + threadIdx.x;
  __syncthreads();
  // exact distances of the candidates, (distance bits, list id) keys
  const uint32_t ld4 = a.ld / 4;
  const float4* Q4 = reinterpret_cast<const float4*>(a.Q) + q * ld4;
  unsigned long long sk[OUT];
#pragma unroll
  for (int i = 0; i < OUT; ++i) {
    const uint32_t p = threadIdx.x * OUT + i;       // blocked arrangement
    sk[i] = ~0ull;
    if (p < total) {
      const uint32_t j = s_sel[p];
      const float4* C4 = reinterpret_cast<const float4*>(a.C) + (size_t)j * ld4;
      float acc = 0.0f;
      for (uint32_t c = 0; c < ld4; ++c) {
        const float4 v = __ldg(C4 + c), qv = __ldg(Q4 + c);
        acc = dist_step<SPF_METRIC_EUCLIDEAN>(acc, qv.x, v.x);
        acc = dist_step<SPF_METRIC_EUCLIDEAN>(acc, qv.y, v.y);
        acc = dist_step<SPF_METRIC_EUCLIDEAN>(acc, qv.z, v.z);
        acc = dist_step<SPF_METRIC_EUCLIDEAN>(acc, qv.w, v.w);
      }
      sk[i] = ((unsigned long long)__float_as_uint(acc) << 12) | j;
    }
  }
  __syncthreads();
  Sort(tmp.sort).Sort(sk, 0, 44);
  __syncthreads();
  uint32_t len[OUT], base[OUT];
#pragma unroll
  for (int i = 0; i < OUT; ++i) {
    const uint32_t p = threadIdx.x * OUT + i;
    len[i] = (p < a.nprobe && sk[i] != ~0ull) ? a.lens[(uint32_t)(sk[i] & 0xfffull)] : 0u;
  }
  Scan(tmp.scan).ExclusiveSum(len, base);
#pragma unroll
  for (int i = 0; i < OUT; ++i) {
    const uint32_t p = threadIdx.x * OUT + i;
    if (p < a.nprobe) {
      a.probe[q * a.nprobe + p] = (uint32_t)(sk[i] & 0xfffull);
      a.seqbase[q * a.nprobe + p] = base[i];
    }
  }
  if (threadIdx.x == 0) {
    // :165  F::from(1.2) * (nearest.distance + F::epsilon())
    const float d0 = __uint_as_float((uint32_t)(sk[0] >> 12));
    a.thr[q] = __fmul_rn(a.prune_factor, __fadd_rn(d0, 1.1920929e-7f));
  }
}

}  // namespace

bool scan_tc_supported(const spf_ctx* c, uint32_t ld, uint64_t nslots, uint32_t K, uint64_t npairs) {
  return c->tma_encode != nullptr && ld % 4 == 0 && ld <= 1024u && K <= (uint32_t)TOPR_MAX &&
         nslots > 0 && nslots + BN < (1ull << 31) && npairs < (1ull << 32) - 1;
}

void scan_tc_release(ScanTcSide* side) {
  if (!side) return;
  if (side->vtf) cudaFree(side->vtf);
  if (side->vext) cudaFree(side->vext);
  if (side->vstat) cudaFree(side->vstat);
  *side = ScanTcSide();
}

int scan_tc_prepare(spf_ctx* c, const float* vecs, const uint64_t* slot_ids, uint64_t nslots, uint32_t ld,
                    ScanTcSide* side) {
  if (side->ready) return SPF_OK;
  cudaStream_t st = c->stream;
  scan_tc_release(side);
  if (cudaMalloc((void**)&side->vtf, nslots * ld * sizeof(float)) != cudaSuccess ||
      cudaMalloc((void**)&side->vext, nslots * 8 * sizeof(float)) != cudaSuccess ||
      cudaMalloc((void**)&side->vstat, 2 * sizeof(float)) != cudaSuccess) {
    scan_tc_release(side);
    cudaGetLastError();
    return fail(SPF_E_OOM, "tensor-scan side structures for %llu slots do not fit", (unsigned long long)nslots);
  }
  DevBuf<float> vnorm, vres;
  SPF_TRY(vnorm.alloc(st, nslots));
  SPF_TRY(vres.alloc(st, nslots));
  SPF_CUDA(cudaMemsetAsync(side->vstat, 0, 2 * sizeof(float), st));
  slot_prep_kernel<<<(unsigned)ceil_div(nslots, 256), 256, 0, st>>>(vecs, slot_ids, nslots / 32, ld / 4, side->vtf,
                                                                    side->vext, vnorm.p, vres.p);
  SPF_TRY(check_launch(c, "slot_prep_kernel"));
  max2_atomic_kernel<<<(unsigned)(c->sm_count * 4), 256, 0, st>>>(vnorm.p, vres.p, nslots, side->vstat);
  SPF_TRY(check_launch(c, "max2_atomic_kernel"));
  SPF_CUDA(cudaStreamSynchronize(st));
  side->nslots = nslots;
  side->ready = true;
  return SPF_OK;
}

int scan_tc_run(spf_ctx* c, const ScanTcCall& call) {
  cudaStream_t st = c->stream;
  const ScanArgs& s = call.s;
  const ScanTcSide& side = *call.side;
  const uint64_t nq = call.nq, npairs = nq * s.nprobe;
  const uint32_t ld = s.ld, nlists = call.nlists;
  // candidate entries per query; long rows have a wider error bound and therefore more candidates
  const uint32_t cap = (uint32_t)(c->params.scan_tc_bucket > 0 ? c->params.scan_tc_bucket : (s.ld <= 256 ? 256 : 1024));
  const uint32_t tau_probes = c->params.scan_tc_tau_probes > 0 ? (uint32_t)c->params.scan_tc_tau_probes : s.nprobe;
  const uint32_t topr = s.K <= 16 ? 16u : 32u;
  // timer / counter names of this call (the centroid probe runs through the same code)
  const bool pr = call.is_probe;
  const char* n_a = pr ? "probe_tc_a" : "scan_tc_a";
  const char* n_gather = pr ? "probe_tc_gather" : "scan_tc_gather";
  const char* n_tau = pr ? "probe_tc_tau" : "scan_tc_tau";
  const char* n_b = pr ? "probe_tc_b" : "scan_tc_b";
  const char* n_ref = pr ? "probe_tc_refine" : "scan_tc_refine";

  // rounded queries + norms
  DevBuf<float> qtf, qnorm, qres, qthr;
  SPF_TRY(qtf.alloc(st, (size_t)nq * ld));
  SPF_TRY(qnorm.alloc(st, nq));
  SPF_TRY(qres.alloc(st, nq));
  SPF_TRY(qthr.alloc(st, nq));
  SPF_TRY(launch_row_prep(c, s.Q, ld, nullptr, nq, qtf.p, qnorm.p, qres.p));

  // units
  DevBuf<uint32_t> ucnt, uoff;
  DevBuf<uint8_t> tmp;
  DevBuf<unsigned long long> utot;
  SPF_TRY(uoff.alloc(st, (size_t)nlists + 1));
  uint32_t nunits = 0;
  // units, tiles over all units, tiles of the longest probed list, units / tiles of the multi-unit lists
  unsigned long long h_tot[5] = {0, 0, 0, 0, 0};
  DevBuf<uint32_t> ukey, uval, ukey2, order;
  const bool one_list = call.is_probe && nlists == 1;
  if (one_list) {
    // the centroid probe: every query probes the one list, nothing to count, nothing to sort, no round trip
    nunits = (uint32_t)ceil_div(nq, (uint64_t)UNIT_ROWS);
    SPF_TRY(ukey2.alloc(st, nunits));
    SPF_TRY(order.alloc(st, nunits));
    tc_probe_units_kernel<<<(unsigned)ceil_div((uint64_t)nunits, 256), 256, 0, st>>>(s.grp_off, nunits, uoff.p, ukey2.p, order.p);
    SPF_TRY(check_launch(c, "tc_probe_units_kernel"));
  } else {
    SPF_TRY(ucnt.alloc(st, (size_t)nlists + 1));
    SPF_TRY(utot.alloc(st, 5));
    SPF_CUDA(cudaMemsetAsync(utot.p, 0, 5 * sizeof(unsigned long long), st));
    tc_unit_counts_kernel<<<(unsigned)ceil_div((uint64_t)nlists + 1, 256), 256, 0, st>>>(call.list_off, s.grp_off, nlists, ucnt.p, utot.p);
    SPF_TRY(check_launch(c, "tc_unit_counts_kernel"));
    size_t tmp_bytes = 0;
    SPF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ucnt.p, uoff.p, (int)(nlists + 1), st));
    SPF_TRY(tmp.alloc(st, tmp_bytes));
    SPF_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, ucnt.p, uoff.p, (int)(nlists + 1), st));
    c->launches += 1;
    SPF_CUDA(cudaMemcpyAsync(h_tot, utot.p, sizeof(h_tot), cudaMemcpyDeviceToHost, st));
    SPF_CUDA(cudaStreamSynchronize(st));
    if (h_tot[0] >= (1ull << 32)) return fail(SPF_E_INVALID, "scan_tc: too many units");
    nunits = (uint32_t)h_tot[0];
    if (nunits > 0) {
      SPF_TRY(ukey.alloc(st, nunits));
      SPF_TRY(uval.alloc(st, nunits));
      SPF_TRY(ukey2.alloc(st, nunits));
      SPF_TRY(order.alloc(st, nunits));
      tc_unit_keys_kernel<<<(unsigned)ceil_div(nunits, 256), 256, 0, st>>>(uoff.p, s.grp_off, nlists, nunits, ukey.p, uval.p);
      SPF_TRY(check_launch(c, "tc_unit_keys_kernel"));
      size_t sb = 0;
      SPF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sb, ukey.p, ukey2.p, uval.p, order.p, (int)nunits, 0, 32, st));
      DevBuf<uint8_t> stmp;
      SPF_TRY(stmp.alloc(st, sb));
      SPF_CUDA(cub::DeviceRadixSort::SortPairs(stmp.p, sb, ukey.p, ukey2.p, uval.p, order.p, (int)nunits, 0, 32, st));
      c->launches += 1;
    }
  }

  DevBuf<uint32_t> qcnt;
  DevBuf<uint2> bucket;
  DevBuf<uint4> ebucket;
  bool exact_entries = false;
  DevBuf<unsigned long long> stats;
  SPF_TRY(stats.alloc(st, 4));
  SPF_CUDA(cudaMemsetAsync(stats.p, 0, 4 * sizeof(unsigned long long), st));
  SPF_TRY(qcnt.alloc(st, nq));
  SPF_TRY(bucket.alloc(st, (size_t)nq * cap));
  SPF_CUDA(cudaMemsetAsync(qcnt.p, 0, nq * sizeof(uint32_t), st));
  SPF_CUDA(cudaMemsetAsync(call.qflag, 0, nq, st));

  // One GEMM pass or two?  When the chunk maxima of the bound pass fit in the budget they are kept
  // (4 KB per tile) and the candidates come from an exact re-evaluation of the few 32-slot groups
  // whose maximum passes ("group refinement"); otherwise — and for the centroid probe, where every
  // query has nprobe hits in one short list — a second GEMM pass emits the candidates.
  uint32_t total_tiles = 0, max_tiles = 0;
  DevBuf<uint32_t> utiles, tile_off;
  bool keep_cmax = false;
  // (a flagged group costs 32 exact vectors = 128 * ld bytes: beyond ld = 256 the second GEMM pass,
  // which pins down the single candidates, is cheaper than re-evaluating whole groups)
  if (nunits > 0 && !call.is_probe && c->params.scan_tc_cmax_mb > 0 && ld <= 256) {
    SPF_TRY(utiles.alloc(st, (size_t)nunits + 1));
    SPF_TRY(tile_off.alloc(st, (size_t)nunits + 1));
    tc_unit_tiles_kernel<<<(unsigned)ceil_div((uint64_t)nunits + 1, 256), 256, 0, st>>>(ukey2.p, nunits, utiles.p);
    SPF_TRY(check_launch(c, "tc_unit_tiles_kernel"));
    size_t sb = 0;
    SPF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, sb, utiles.p, tile_off.p, (int)(nunits + 1), st));
    DevBuf<uint8_t> stmp;
    SPF_TRY(stmp.alloc(st, sb));
    SPF_CUDA(cub::DeviceScan::ExclusiveSum(stmp.p, sb, utiles.p, tile_off.p, (int)(nunits + 1), st));
    c->launches += 1;
    // totals came back with the unit count: no second round trip
    if (h_tot[1] >= (1ull << 32)) return fail(SPF_E_INVALID, "scan_tc: too many tiles");
    total_tiles = (uint32_t)h_tot[1];
    max_tiles = (uint32_t)h_tot[2];
    keep_cmax = (uint64_t)total_tiles * 5120ull <= (uint64_t)c->params.scan_tc_cmax_mb << 20;
  }

  if (nunits > 0) {
    uint32_t max_units = (uint32_t)((1ull << 30) / ((uint64_t)UNIT_ROWS * ld * sizeof(float)));   // <= 1 GB of gathered rows
    if (max_units < 64u) max_units = 64u;
    const uint32_t chunk_units = nunits < max_units ? nunits : max_units;
    const bool single = chunk_units == nunits;
    DevBuf<float> A, rowthr, pairtop, qbound;
    DevBuf<uint32_t> rowseq, rowpair, unit_slot0;
    DevBuf<UnitDesc> desc;
    DevBuf<float4> cmax;
    DevBuf<uint32_t> cmask;
    DevBuf<float2> rowes;
    SPF_TRY(A.alloc(st, (size_t)chunk_units * UNIT_ROWS * ld));
    SPF_TRY(rowthr.alloc(st, (size_t)nunits * UNIT_ROWS));
    SPF_TRY(rowseq.alloc(st, (size_t)nunits * UNIT_ROWS));
    SPF_TRY(rowpair.alloc(st, (size_t)nunits * UNIT_ROWS));
    SPF_TRY(unit_slot0.alloc(st, nunits));
    SPF_TRY(desc.alloc(st, chunk_units));
    SPF_TRY(pairtop.alloc(st, npairs * 2 * topr));
    SPF_TRY(qbound.alloc(st, nq));
    if (keep_cmax) {
      SPF_TRY(cmax.alloc(st, (size_t)total_tiles * 2 * UNIT_ROWS));
      SPF_TRY(cmask.alloc(st, (size_t)total_tiles * 2 * UNIT_ROWS));
      SPF_TRY(rowes.alloc(st, (size_t)nunits * UNIT_ROWS));
    }

    CUtensorMap map_a, map_b, map_e;
    SPF_TRY(make_map_k128(c, &map_a, A.p, (uint64_t)chunk_units * UNIT_ROWS, ld, BM));
    SPF_TRY(make_map_k128(c, &map_b, side.vtf, side.nslots, ld, BN));
    SPF_TRY(make_map_ext(c, &map_e, side.vext, side.nslots, BN));
    SPF_CUDA(cudaFuncSetAttribute(scan_tc_kernel<0, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    SPF_CUDA(cudaFuncSetAttribute(scan_tc_kernel<0, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    SPF_CUDA(cudaFuncSetAttribute(scan_tc_kernel<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));

    GatherArgs g;
    g.pair_sorted = call.pair_sorted; g.list_off = call.list_off; g.unit_off = uoff.p; g.nlists = nlists;
    g.grp_off = s.grp_off; g.lens = s.lens; g.seqbase = s.seqbase;
    g.nprobe = s.nprobe; g.ld4 = ld / 4; g.d = s.d;
    g.qtf = qtf.p; g.tau_probes = keep_cmax ? s.nprobe : tau_probes; g.order = order.p;
    g.tile_off = keep_cmax ? tile_off.p : nullptr;
    g.A = A.p; g.desc = desc.p; g.rowthr = rowthr.p; g.rowseq = rowseq.p; g.rowpair = rowpair.p;
    g.unit_slot0 = unit_slot0.p;
    g.rowes = keep_cmax ? rowes.p : nullptr; g.qnorm = qnorm.p; g.qres = qres.p; g.vstat = side.vstat;
    g.stream_bytes = stats.p + 2;

    ScanTcArgs k;
    k.kb = (ld + BK - 1) / BK; k.nprobe = s.nprobe; k.cap = cap;
    k.desc = desc.p; k.rowthr = rowthr.p; k.rowseq = rowseq.p; k.rowpair = rowpair.p;
    k.pairtop = pairtop.p; k.qcnt = qcnt.p; k.bucket = bucket.p;
    k.cmax = keep_cmax ? cmax.p : nullptr;
    k.cmask = keep_cmax ? cmask.p : nullptr; k.rowes = rowes.p; k.topk = s.K;
    k.dense = nullptr; k.dense_ld = 0; k.ubase = 0;

    for (uint32_t u0 = 0; u0 < nunits; u0 += chunk_units) {
      const uint32_t nu = nunits - u0 < chunk_units ? nunits - u0 : chunk_units;
      {
        KernelTimer t(c, n_gather);
        g.u0 = u0; g.qthr = nullptr; g.copy_rows = 1; g.bytes = s.bytes;
        unit_gather_kernel<<<nu, 256, 0, st>>>(g);
        SPF_TRY(check_launch(c, "unit_gather_kernel"));
      }
      KernelTimer t(c, n_a);                              // the bound pass alone (the roofline's kernel time)
      k.nunits = nu; k.u0 = u0; k.ubase = 0;
      auto launch_a = [&](unsigned grid, cudaStream_t on) {
        if (topr == 16) scan_tc_kernel<0, 16><<<grid, NUM_THREADS, SMEM_TOTAL, on>>>(map_a, map_b, map_e, k);
        else scan_tc_kernel<0, 32><<<grid, NUM_THREADS, SMEM_TOTAL, on>>>(map_a, map_b, map_e, k);
        return check_launch(c, "scan_tc_kernel<A>");
      };
      // Split launch.  The units of lists probed by more than 128 pairs re-read their list from L2 and are
      // bound by the tensor pipe; the single-unit lists stream from HBM and leave the tensor pipe half idle.
      // One after the other (sorted order) the two phases add up; as two persistent kernels on disjoint
      // SMs they overlap, and each list's units still run side by side.  Measured on the 10 k-query batch
      // (438 multi-unit units holding 42 % of the 51.5 k tiles): 1.10 ms unsplit, 1.42 / 1.05 / 0.94 / 0.97 /
      // 1.04 / 1.27 ms with 40 / 61 / 70 / 76 / 84 / 100 SMs on the multi-unit side — a little more than its
      // share of the tiles, because those units pay a per-unit prologue and the other side is HBM-bound
      // anyway.  With 30 k queries (82 % of the tiles in multi-unit lists) nothing is left to overlap:
      // 1.69 ms either way, so the split is used only while the HBM-bound side holds 40 % of the tiles.
      const uint32_t hub_units = (uint32_t)h_tot[3];
      const uint32_t sms = (uint32_t)c->sm_count;
      unsigned grid_hub = 0;
      if (single && !one_list && hub_units >= 1 && nunits - hub_units >= 1 && h_tot[1] > 0 && sms >= 16) {
        const double share = (double)h_tot[4] / (double)h_tot[1];
        if (c->params.scan_tc_split > 1) {                  // tests / experiments: fixed SM count, no size conditions
          grid_hub = (unsigned)c->params.scan_tc_split < sms - 1 ? (unsigned)c->params.scan_tc_split : sms - 1;
        } else if (c->params.scan_tc_split == 1 && hub_units >= 8 && nunits - hub_units >= 2 * sms &&
                   share >= 0.05 && share <= 0.6) {
          grid_hub = (unsigned)(1.2 * share * sms + 0.5);
          if (grid_hub < 4) grid_hub = 4;
          if (grid_hub > sms - 8) grid_hub = sms - 8;
        }
        if (grid_hub > hub_units) grid_hub = hub_units;
      }
      if (grid_hub == 0) {
        SPF_TRY(launch_a(nu < sms ? nu : sms, st));
      } else {
        if (!c->aux_stream) SPF_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
        if (!c->aux_ev[0]) {
          SPF_CUDA(cudaEventCreateWithFlags(&c->aux_ev[0], cudaEventDisableTiming));
          SPF_CUDA(cudaEventCreateWithFlags(&c->aux_ev[1], cudaEventDisableTiming));
        }
        SPF_CUDA(cudaEventRecord(c->aux_ev[0], st));                     // gathered rows and row scalars are ready
        SPF_CUDA(cudaStreamWaitEvent(c->aux_stream, c->aux_ev[0], 0));
        k.ubase = 0; k.nunits = hub_units;
        SPF_TRY(launch_a(grid_hub, c->aux_stream));
        SPF_CUDA(cudaEventRecord(c->aux_ev[1], c->aux_stream));
        k.ubase = hub_units; k.nunits = nunits - hub_units;
        const int rc_main = launch_a(k.nunits < sms - grid_hub ? k.nunits : sms - grid_hub, st);
        SPF_CUDA(cudaStreamWaitEvent(st, c->aux_ev[1], 0));              // both halves done before tau (and before any free)
        SPF_TRY(rc_main);
        k.ubase = 0; k.nunits = nu;
        if (c->profiling) c->kernel_ms[pr ? "probe_tc_split" : "scan_tc_split"] = (float)grid_hub;
      }
    }
    {
      KernelTimer t(c, n_tau);
      TauArgs ta;
      ta.nq = nq; ta.nprobe = s.nprobe; ta.tau_probes = g.tau_probes; ta.K = s.K; ta.ld = ld; ta.topr = topr;
      ta.probe = s.probe; ta.grp_off = s.grp_off; ta.pairtop = pairtop.p; ta.qnorm = qnorm.p; ta.qres = qres.p;
      ta.thr = s.thr; ta.vstat = side.vstat; ta.qthr = qthr.p; ta.qbound = qbound.p; ta.qflag = call.qflag;
      tau_kernel<<<(unsigned)ceil_div(nq * 32, 256), 256, 0, st>>>(ta);
      SPF_TRY(check_launch(c, "tau_kernel"));
    }
    if (call.comm && !call.is_probe) {
      KernelTimer t(c, "scan_tc_bound_exchange");
      bound_sanitize_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>(qbound.p, nq);
      SPF_TRY(check_launch(c, "bound_sanitize_kernel"));
      SPF_TRY(comm_allreduce_min_f32(c, call.comm, qbound.p, nq));
      if (call.bound_exchanged) *call.bound_exchanged = true;
      bound_tighten_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>(qbound.p, qnorm.p, qres.p, side.vstat, ld, nq, qthr.p);
      SPF_TRY(check_launch(c, "bound_tighten_kernel"));
    }
    if (keep_cmax) {
      // group refinement: queue the (pair, group) items whose chunk maximum passes, evaluate them exactly
      KernelTimer t(c, n_b);
      const uint64_t want = nq * 128ull > (1ull << 20) ? nq * 128ull : (1ull << 20);
      const uint32_t work_cap = (uint32_t)(want < (1ull << 30) ? want : (1ull << 30));
      DevBuf<WorkItem> work;
      DevBuf<uint32_t> nwork;
      SPF_TRY(work.alloc(st, work_cap));
      SPF_TRY(nwork.alloc(st, 1));
      SPF_TRY(ebucket.alloc(st, (size_t)nq * cap));
      SPF_CUDA(cudaMemsetAsync(nwork.p, 0, sizeof(uint32_t), st));
      FlagArgs f;
      f.keys_sorted = ukey2.p; f.tile_off = tile_off.p; f.cmax = cmax.p; f.cmask = cmask.p; f.rowpair = rowpair.p;
      f.unit_slot0 = unit_slot0.p; f.qthr = qthr.p; f.nprobe = s.nprobe;
      f.work = work.p; f.work_cap = work_cap; f.nwork = nwork.p; f.qflag = call.qflag;
      {
        KernelTimer tf(c, "scan_tc_flag");
        chunk_flag_kernel<<<dim3(nunits, (max_tiles + FLAG_TILES - 1) / FLAG_TILES), 256, 0, st>>>(f);
        SPF_TRY(check_launch(c, "chunk_flag_kernel"));
      }
      GroupArgs ga;
      ga.s = s; ga.work = work.p; ga.nwork = nwork.p; ga.work_cap = work_cap; ga.qbound = qbound.p; ga.cap = cap;
      ga.qcnt = qcnt.p; ga.ebucket = ebucket.p;
      group_refine_kernel<<<(unsigned)(c->sm_count * 8), 256, 0, st>>>(ga);
      SPF_TRY(check_launch(c, "group_refine_kernel"));
      exact_entries = true;
      if (c->profiling) {
        uint32_t h = 0;
        SPF_CUDA(cudaMemcpyAsync(&h, nwork.p, sizeof(h), cudaMemcpyDeviceToHost, st));
        SPF_CUDA(cudaStreamSynchronize(st));
        c->kernel_ms["scan_tc_groups"] = (float)h;
      }
    } else {
      KernelTimer t(c, n_b);
      for (uint32_t u0 = 0; u0 < nunits; u0 += chunk_units) {
        const uint32_t nu = nunits - u0 < chunk_units ? nunits - u0 : chunk_units;
        g.u0 = u0; g.qthr = qthr.p; g.copy_rows = single ? 0 : 1; g.bytes = nullptr;
        unit_gather_kernel<<<nu, 256, 0, st>>>(g);
        SPF_TRY(check_launch(c, "unit_gather_kernel"));
        k.nunits = nu; k.u0 = u0;
        const unsigned grid = nu < (uint32_t)c->sm_count ? nu : (unsigned)c->sm_count;
        scan_tc_kernel<1, 16><<<grid, NUM_THREADS, SMEM_TOTAL, st>>>(map_a, map_b, map_e, k);
        SPF_TRY(check_launch(c, "scan_tc_kernel<B>"));
      }
    }
  }
  {
    KernelTimer t(c, n_ref);
    RefineArgs r;
    r.s = s; r.nq = nq; r.cap = cap; r.qcnt = qcnt.p; r.bucket = bucket.p; r.qflag = call.qflag;
    r.ebucket = ebucket.p; r.stats = stats.p;
    if (exact_entries) refine_kernel<true><<<(unsigned)ceil_div(nq * 32, 256), 256, 0, st>>>(r);
    else refine_kernel<false><<<(unsigned)ceil_div(nq * 32, 256), 256, 0, st>>>(r);
    SPF_TRY(check_launch(c, "refine_kernel"));
  }
  if (c->profiling) {   // counters for bench / profiling runs, reported through spf_ctx_kernel_ms
    unsigned long long h[4] = {0, 0, 0, 0};
    SPF_CUDA(cudaMemcpyAsync(h, stats.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    SPF_CUDA(cudaStreamSynchronize(st));
    c->kernel_ms[pr ? "probe_tc_candidates" : "scan_tc_candidates"] = (float)h[0];
    c->kernel_ms[pr ? "probe_tc_flagged" : "scan_tc_flagged"] = (float)h[1];
    c->kernel_ms[pr ? "probe_tc_units" : "scan_tc_units"] = (float)nunits;
    c->kernel_ms[pr ? "probe_tc_hub_units" : "scan_tc_hub_units"] = (float)h_tot[3];
    c->kernel_ms[pr ? "probe_tc_hub_tiles" : "scan_tc_hub_tiles"] = (float)h_tot[4];
    c->kernel_ms[pr ? "probe_tc_tiles" : "scan_tc_tiles"] = (float)h_tot[1];
    c->kernel_ms[pr ? "probe_tc_stream_mb" : "scan_tc_stream_mb"] = (float)((double)h[2] / 1e6);
    c->kernel_ms[pr ? "probe_tc_unique_mb" : "scan_tc_unique_mb"] = (float)((double)h[3] / 1e6);
  }
  return SPF_OK;
}

// Dense tensor-core probe for 32 < nprobe <= 1024 and nlists <= 4096 (see probe_dense_select_kernel).
// `side` / `centroids`: the centroid list's TF32 side structures and the exact row-major centroids.
// Sets *redo (device) when some query must go through the exact probe instead.
int probe_tc_dense(spf_ctx* c, const ScanTcSide& side, const float* centroids, const float* Q, uint64_t nq, uint32_t ld,
                   uint32_t nlists, uint32_t nprobe, float prune_factor, const uint32_t* lens, uint32_t* probe,
                   float* thr, uint32_t* seqbase, int* d_redo) {
  cudaStream_t st = c->stream;
  const uint32_t cslots = (uint32_t)side.nslots;
  const uint64_t chunk = nq < 32768 ? nq : 32768;          // 512 MB of s per chunk at 4096 lists
  const uint32_t cunits = (uint32_t)ceil_div(chunk, UNIT_ROWS);
  DevBuf<float> qtf, qnorm, qres, S, rowthr;
  DevBuf<uint32_t> rowseq, rowpair;
  DevBuf<UnitDesc> desc;
  SPF_TRY(qtf.alloc(st, (size_t)nq * ld));
  SPF_TRY(qnorm.alloc(st, nq));
  SPF_TRY(qres.alloc(st, nq));
  SPF_TRY(S.alloc(st, (size_t)chunk * cslots));
  SPF_TRY(rowthr.alloc(st, (size_t)cunits * UNIT_ROWS));
  SPF_TRY(rowseq.alloc(st, (size_t)cunits * UNIT_ROWS));
  SPF_TRY(rowpair.alloc(st, (size_t)cunits * UNIT_ROWS));
  SPF_TRY(desc.alloc(st, cunits));
  SPF_TRY(launch_row_prep(c, Q, ld, nullptr, nq, qtf.p, qnorm.p, qres.p));
  CUtensorMap map_a, map_b, map_e;
  SPF_TRY(make_map_k128(c, &map_b, side.vtf, side.nslots, ld, BN));
  SPF_TRY(make_map_ext(c, &map_e, side.vext, side.nslots, BN));
  SPF_CUDA(cudaFuncSetAttribute(scan_tc_kernel<2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
  for (uint64_t q0 = 0; q0 < nq; q0 += chunk) {
    const uint32_t nc = (uint32_t)((nq - q0) < chunk ? (nq - q0) : chunk);
    const uint32_t nu = (uint32_t)ceil_div(nc, UNIT_ROWS);
    SPF_TRY(make_map_k128(c, &map_a, qtf.p + q0 * ld, nc, ld, BM));   // the chunk's rounded queries, no gather needed
    dense_setup_kernel<<<(unsigned)ceil_div((uint64_t)nu * UNIT_ROWS, 256), 256, 0, st>>>(nu, nc, cslots, desc.p, rowthr.p,
                                                                                          rowseq.p, rowpair.p);
    SPF_TRY(check_launch(c, "dense_setup_kernel"));
    ScanTcArgs k;
    k.nunits = nu; k.kb = (ld + BK - 1) / BK; k.nprobe = 1; k.cap = 0; k.u0 = 0; k.ubase = 0;
    k.desc = desc.p; k.cmax = nullptr; k.cmask = nullptr; k.rowes = nullptr; k.topk = 0;
    k.rowthr = rowthr.p; k.rowseq = rowseq.p; k.rowpair = rowpair.p;
    k.pairtop = nullptr; k.qcnt = nullptr; k.bucket = nullptr;
    k.dense = S.p; k.dense_ld = cslots;
    {
      KernelTimer t(c, "probe_tc_dense");
      const unsigned grid = nu < (uint32_t)c->sm_count ? nu : (unsigned)c->sm_count;
      scan_tc_kernel<2, 16><<<grid, NUM_THREADS, SMEM_TOTAL, st>>>(map_a, map_b, map_e, k);
      SPF_TRY(check_launch(c, "scan_tc_kernel<dense>"));
    }
    DenseSelArgs a;
    a.S = S.p; a.s_ld = cslots; a.Q = Q + q0 * ld; a.C = centroids; a.ld = ld;
    a.qnorm = qnorm.p + q0; a.qres = qres.p + q0; a.vstat = side.vstat;
    a.nlists = nlists; a.nprobe = nprobe; a.prune_factor = prune_factor;
    a.lens = lens; a.probe = probe + q0 * nprobe; a.thr = thr + q0; a.seqbase = seqbase + q0 * nprobe; a.redo = d_redo;
    {
      KernelTimer t(c, "probe_tc_select");
      if (nprobe <= 192) probe_dense_select_kernel<16, 1><<<nc, 256, 0, st>>>(a);
      else if (nprobe <= 448) probe_dense_select_kernel<16, 2><<<nc, 256, 0, st>>>(a);
      else probe_dense_select_kernel<16, 5><<<nc, 256, 0, st>>>(a);
      SPF_TRY(check_launch(c, "probe_dense_select_kernel"));
    }
  }
  return SPF_OK;
}

}  // namespace spf
