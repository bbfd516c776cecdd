// assign_tc.cu — tcgen05 / TMEM / TMA candidate GEMM for squared-Euclidean assignment (sm_100a).
//
// Replaces the Euclidean case of the n x k distance loop in assign_points_to_clusters,
// src/clustering/hierarchical.rs:302-326 (reference).  d(x,c) = |x|^2 - 2 x.c + |c|^2 is a dense
// contraction: X.C^T runs on the 5th-gen tensor cores as ONE TF32 pass with fp32 accumulation
// in TMEM; the m x k matrix is never written.  |c|^2 rides along as an extra K = 8 block
// (-|c|^2/2 split into three TF32 terms against ones), so the accumulator holds
// s = x.c - |c|^2/2 and d = |x|^2 - 2 s.  The epilogue keeps, per point, a running minimum of d
// (maximum of s) and emits only the centroids that can still matter:
//     d_tf32 < f * (runmin_tf32 + E) + E,      E = tc_err_bound(|x|^2, max|c|^2, ld)
// which is a superset of {argmin candidates} U {j : d_ref(j) < f * dmin_ref} because E bounds
// |d_tf32 - d_ref| and the running minimum only decreases.  resolve.cu then decides every
// comparison on exact direct-form values, so the final result is bit-identical to the
// reference while the O(n k d) work runs at tensor-core rate (1 pass instead of the 3 a
// 3xTF32 split would need).
//
// Candidates leave the kernel as "group records" (kernels.cuh): when any of four consecutive
// columns passes the test the thread stores all four t values with one 128-bit store, which keeps
// the per-hit instruction cost low; resolve filters the bystanders.
//
// Kernel anatomy (persistent, one CTA per SM, 320 threads):
//   warp 0     TMA producer: the 128 x ld point tile (A, stationary for a whole row block) and a
//              4-stage ring of 256 x 32 centroid tiles (B), both SWIZZLE_128B, K-major
//   warp 1     TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=256, K=8 per MMA)
//   warps 2-9  epilogue (two warps per SM sub-partition so one hides the other's latencies):
//              warp w reads TMEM lane quarter w%4 and column half (w-2)/4 of the double-buffered
//              2 x 256 column accumulator: all 128 columns of the thread's point go to registers
//              (4 x tcgen05.ld 32x32b.x32), the accumulator is released, then per 32-column chunk:
//              running minimum (shared between the two halves through shared memory), one-FFMA
//              threshold, candidate emission with predicated 256-bit stores.
// Round-2 measurements (1M x 128, k = 4096, profiles/r02_experiment_notes.md): 16 epilogue warps
// (96 registers) 2.10 ms, pipelined loads + two 128-bit stores 1.85 ms, 256-bit stores 1.80 ms,
// load-all-first + 256-bit stores 1.76 ms, + per-chunk capacity vote 1.74 ms; the same instruction
// stream with no record stores 1.66 ms, MMA + TMA alone 1.48 ms.  Also tried and dropped: records as
// two parallel arrays (values + indices, 1.79 ms: two store transactions per record cost more than
// the operand-register wait of the single 256-bit store), two alternating index register quads
// (1.77 ms: the kernel sits at the 168-register limit of 10 warps per SM).
// An optional per-point seed (a certified upper bound of the point's minimum distance, e.g. the
// running minimum of the k-means++ rounds or the distance to the previous iteration's centroid)
// tightens the candidate test from the first column on; resolve validates it a posteriori.
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
// K extension: one extra K = 8 block per tile carries -|c|^2/2 (split into three TF32 terms), so
// the accumulator holds s = x.c - |c|^2/2 and the epilogue needs neither |c|^2 nor an FMA per
// element (d = |x|^2 - 2 s).  Rows of 8 floats = 32 bytes, SWIZZLE_32B.
constexpr int NEXT = 2;                                      // ring depth of the extension tiles
constexpr int AEXT_BYTES = BM * EXT_K * 4;                   // 4 KB, constant for the whole kernel
constexpr int BEXT_BYTES = BN * EXT_K * 4;                   // 8 KB per tile
constexpr int SMEM_AEXT_OFF = SMEM_A + SMEM_B;
constexpr int SMEM_BEXT_OFF = SMEM_AEXT_OFF + AEXT_BYTES;
constexpr int SMEM_BAR_OFF = SMEM_BEXT_OFF + NEXT * BEXT_BYTES;
constexpr int SMEM_PUB_OFF = SMEM_BAR_OFF + 256;             // published (row block, running min) [2][128]
constexpr int SMEM_TOTAL = SMEM_PUB_OFF + 2 * BM * 8 + 1024; // + alignment slack

struct TcArgs {
  uint32_t m, k, ld, kb;            // kb = ceil(ld / 32) K blocks
  uint32_t eld;                     // row length entering the certified error bound
  uint32_t ntiles;                  // ceil(k / 256)
  uint32_t nrowblocks;              // ceil(m / 128)
  float factor;
  const float* xnorm; const float* xres; const float* cstat;
  const float* seed;                // optional: per point an upper bound of its minimum distance
  CandRec* rec; RowInfo* info; int cap;
};

// PIPE 0: the thread's 128 accumulator columns are loaded first, the buffer is released, then the
// arithmetic runs (the load latency of every tile is exposed once).  PIPE 1: two halves of 64 columns
// in a software pipeline that runs ACROSS tiles — the second half of tile i is in flight while the
// first is processed, the buffer is released in the middle of the tile, and the first half of tile
// i + 1 is in flight while the second half of tile i is processed: no tcgen05.ld latency is exposed and
// the register footprint (128 value registers) is unchanged.
template <int PIPE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
assign_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_e, TcArgs a) {
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
    // a centroid stage is written by both CTAs of the pair (multicast halves) and must be released
    // by both MMA issuers before either producer may overwrite it
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 2); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], NUM_EPI_WARPS); }
    for (int i = 0; i < NEXT; ++i) { mbar_init(&e_full[i], 1); mbar_init(&e_empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x >= 64) reinterpret_cast<unsigned long long*>(smem + SMEM_PUB_OFF)[threadIdx.x - 64] = ~0ull;
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
  cluster_sync_all();                      // the peer's barriers are initialised before anything remote arrives
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t crank = cluster_ctarank();
  // ld <= 128: the point tile (<= 4 K blocks) stays in shared memory for the whole row block.
  // Longer rows: the point K block is streamed with the centroid K block, one ring stage for both
  // (the 64 KB of the stationary tile are then four 16 KB stages).
  const bool stream_a = a.kb > (uint32_t)KB_MAX;
  static_assert(KB_MAX == NSTAGE, "the stationary A tile and the streamed A stages share one region");

  if (warp == 0) {
    // =============================== TMA producer ===========================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, it = 0, ecount = 0;
      for (uint32_t rb = blockIdx.x; rb < a.nrowblocks; rb += gridDim.x, ++it) {
        for (uint32_t t = 0; t < a.ntiles; ++t, ++ecount) {
          for (uint32_t kb = 0; kb < a.kb; ++kb) {
            if (!stream_a && t == 0) {
              // the row block's point K block, requested as soon as the previous row block's last tile is
              // done with THIS block (not with all four), so the centroid ring does not drain in between
              mbar_wait(&a_empty[kb], (it & 1) ^ 1);
              mbar_expect_tx(&a_full[kb], A_KB_BYTES);
              tma_load_2d(smem_a + kb * A_KB_BYTES, &map_a, &a_full[kb], (int)(kb * BK), (int)(rb * BM));
            }
            mbar_wait(&b_empty[stage], phase ^ 1);          // both CTAs are done with this stage
            if (stream_a) {                                 // own point K block in the same ring stage (not multicast)
              mbar_expect_tx(&b_full[stage], A_KB_BYTES + B_STAGE_BYTES);
              tma_load_2d(smem_a + stage * A_KB_BYTES, &map_a, &b_full[stage], (int)(kb * BK), (int)(rb * BM));
            } else {
              mbar_expect_tx(&b_full[stage], B_STAGE_BYTES);  // own half + the peer's multicast half
            }
            tma_load_2d_mc(smem_b + stage * B_STAGE_BYTES + crank * (B_STAGE_BYTES / 2), &map_b, &b_full[stage],
                           (int)(kb * BK), (int)(t * BN + crank * (BN / 2)), (uint16_t)0x3);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
          const uint32_t es = ecount % NEXT, eu = ecount / NEXT;
          mbar_wait(&e_empty[es], (eu & 1) ^ 1);
          mbar_expect_tx(&e_full[es], BEXT_BYTES);
          tma_load_2d(smem_bext + es * BEXT_BYTES, &map_e, &e_full[es], 0, (int)(t * BN));
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ==============================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, it = 0, tcount = 0;
      for (uint32_t rb = blockIdx.x; rb < a.nrowblocks; rb += gridDim.x, ++it) {
        for (uint32_t t = 0; t < a.ntiles; ++t, ++tcount) {
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
            tc_commit_mc(&b_empty[stage], (uint16_t)0x3);   // frees the B stage in both CTAs once these MMAs retire
            if (!stream_a && t + 1 == a.ntiles) tc_commit(&a_empty[kb]);   // last use of this A block
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
          {   // K extension: accumulator += -|c|^2/2
            const uint32_t es = tcount % NEXT, eu = tcount / NEXT;
            mbar_wait(&e_full[es], eu & 1);
            tc_fence_after();
            tc_mma_tf32(tmem_d, make_smem_desc32(smem_u32(smem_aext)),
                        make_smem_desc32(smem_u32(smem_bext + es * BEXT_BYTES)), IDESC_TF32, 1u);
            tc_commit(&e_empty[es]);
          }
          tc_commit(&t_full[buf]);                        // accumulator complete → epilogue
        }
      }
    }
  } else {
    // =============================== epilogue warps ==========================================
    const uint32_t quarter = warp & 3;                    // TMEM lane quarter this warp may access
    const uint32_t half = (uint32_t)(warp - 2) >> 2;      // column half of every accumulator
    const uint32_t lrow = quarter * 32 + lane;
    const float INF = __int_as_float(0x7f800000);
    const float cnmax = a.cstat[0], dcmax = a.cstat[1];
    const float f1 = fmaxf(a.factor, 1.0f);
    const uint32_t segcap = (uint32_t)a.cap >> 1;
    const uint32_t segbytes = segcap * (uint32_t)sizeof(CandRec);
    const uint32_t pub_mine = smem_u32(smem + SMEM_PUB_OFF) + (half * BM + lrow) * 8u;
    const uint32_t pub_other = smem_u32(smem + SMEM_PUB_OFF) + ((half ^ 1u) * BM + lrow) * 8u;
    // the three don't-care words of a record (never read by resolve): any registers, so that the
    // record is one 256-bit store
    const uint32_t z1 = lrow, z2 = warp, z3 = threadIdx.x;
    uint32_t tcount = 0;
    // PIPE 1: tiles this CTA walks in total, and the first half of the first tile
    const uint32_t total_tiles =
        blockIdx.x < a.nrowblocks ? ((a.nrowblocks - blockIdx.x + gridDim.x - 1) / gridDim.x) * a.ntiles : 0u;
    const uint32_t tlane = tmem_base + ((quarter * 32u) << 16) + half * (BN / 2);
    uint32_t r0[32], r1[32], r2[32], r3[32];
    if (PIPE == 1 && total_tiles > 0) {
      mbar_wait(&t_full[0], 0);
      tc_fence_after();
      tc_ld32_issue(tlane, r0);
      tc_ld32_issue(tlane + 32, r1);
      tc_ld32_wait(r0);
      tc_ld32_wait(r1);
    }
    for (uint32_t rb = blockIdx.x; rb < a.nrowblocks; rb += gridDim.x) {
      const uint32_t row = rb * BM + lrow;
      const bool row_ok = row < a.m;
      const float xn = row_ok ? a.xnorm[row] : 0.0f;
      const float E = tc_err_bound(xn, row_ok ? a.xres[row] : 0.0f, cnmax, dcmax, a.eld);
      // non-finite norms or bounds: no certified test exists, the brute-force kernels own the row
      const bool hopeless = !(E < INF) || !(xn < INF);
      // candidate test  d < f (dmin_run + E) + E  with d = |x|^2 - 2 s, dmin_run = |x|^2 - 2 smax:
      //   s > f smax + ( |x|^2 - E - f (|x|^2 + E) ) / 2
      // one FFMA per chunk; `slop` covers the roundings of this form (a few ulps of |s|, |x|^2).
      const float slop = 1e-6f * (xn + cnmax) + 1e-30f;
      const float thrK = row_ok ? 0.5f * ((xn - E) - f1 * (xn + E)) - slop : INF;
      // seed: some centroid has d_ref <= seed, hence d_tf32 <= seed + E and the final maximum of s
      // is at least tc_seed_bound().  resolve checks that the observed maximum really reaches it
      // and hands the row to the dense fallback otherwise.
      float sseed = -INF;
      if (a.seed != nullptr && row_ok) {
        const float sd = a.seed[row];
        if (sd < INF) sseed = tc_seed_bound(xn, sd, E, cnmax);
      }
      // this thread's segment of the row's records: write pointer and end
      const unsigned long long seg = (unsigned long long)(a.rec + ((size_t)row * a.cap + (size_t)half * segcap));
      const unsigned long long wend = seg + segbytes;
      uint32_t wlo = (uint32_t)seg, whi = (uint32_t)(seg >> 32);   // write pointer
      uint32_t overflow = 0;                              // records that did not fit
      float smax = -INF;                                  // running maximum of s = x.c - |c|^2/2 over this half
      float other = sseed;                                // bound from the seed / the partner half (one chunk old)
      for (uint32_t t = 0; t < a.ntiles; ++t, ++tcount) {
        const uint32_t buf = tcount & 1, use = tcount >> 1;
        if (PIPE == 0) {
          mbar_wait(&t_full[buf], use & 1);
          tc_fence_after();
        }
        const uint32_t taddr = tlane + buf * BN;
        const uint32_t gtile = (t * BN + half * (BN / 2)) >> 2;
        // a tile emits at most 32 records per thread: when every thread of the warp has room for
        // them, the whole tile takes the straight-line path without any further capacity test
        const bool fast = __all_sync(0xffffffffu, (((unsigned long long)whi << 32) | wlo) + 32ull * sizeof(CandRec) <= wend &&
                                                      wlo <= 0xffffffffu - 32u * (uint32_t)sizeof(CandRec));

        // one 32-column chunk of s: running maximum (= running minimum of d = |x|^2 - 2 s), candidate
        // emission
        auto process = [&](uint32_t (&rr)[32], int c) {
          float q[8];
#pragma unroll
          for (int g = 0; g < 8; ++g)
            q[g] = fmaxf(fmaxf(__uint_as_float(rr[g * 4 + 0]), __uint_as_float(rr[g * 4 + 1])),
                         fmaxf(__uint_as_float(rr[g * 4 + 2]), __uint_as_float(rr[g * 4 + 3])));
          smax = fmaxf(smax, fmaxf(fmaxf(fmaxf(q[0], q[1]), fmaxf(q[2], q[3])),
                                   fmaxf(fmaxf(q[4], q[5]), fmaxf(q[6], q[7]))));
          // Exchange running maxima with the partner half of the same row through shared memory
          // (tagged with the row block: any value published for this row block is a valid lower
          // bound of the final maximum, so a stale one only makes the candidate set a little
          // larger).  The partner's value is read one chunk late (`other` was loaded while the
          // previous chunk was processed) to keep the shared-memory round trip off the critical path.
          const float sshare = fmaxf(smax, other);
          {
            uint32_t pv_lo, pv_hi;
            asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(pv_lo), "=r"(pv_hi) : "r"(pub_other) : "memory");
            other = fmaxf(other, pv_hi == rb ? __uint_as_float(pv_lo) : -INF);
            asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(pub_mine), "r"(__float_as_uint(smax)), "r"(rb) : "memory");
          }
          const float thr_s = fmaf(f1, sshare, thrK);
          uint32_t gi = gtile + c * 8;                     // group index of the record being tested
          // room for this chunk's 8 records in every lane of the warp: known for the whole tile
          // (`fast`), else voted per chunk — the branchy slow path only runs when some lane is
          // within 8 records of its segment's end (1.78 -> 1.74 ms)
          const bool fastc = fast || __all_sync(0xffffffffu, (((unsigned long long)whi << 32) | wlo) + 8ull * sizeof(CandRec) <= wend &&
                                                                 wlo <= 0xffffffffu - 8u * (uint32_t)sizeof(CandRec));
          if (fastc) {
            // straight-line predicated record stores, no vote and no branch: one 256-bit store per
            // record — half the store transactions of two 128-bit stores (measured: 1.81 -> 1.76 ms)
            // and the whole 32-byte sector is written, so DRAM never has to read-fill it.  The last
            // three words are don't-care registers.  `fast` guarantees that the low pointer word
            // cannot wrap inside this tile, so the pointer advances with one predicated 32-bit add.
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              asm volatile(
                  "{\n\t.reg .pred p;\n\t.reg .b64 a;\n\t"
                  "setp.gt.f32 p, %3, %4;\n\t"
                  "mov.b64 a, {%0, %1};\n\t"
                  "@p st.global.v8.b32 [a], {%5, %6, %7, %8, %2, %9, %10, %11};\n\t"
                  "@p add.u32 %0, %0, 32;\n\t"
                  "add.u32 %2, %2, 1;\n\t}"
                  : "+r"(wlo), "+r"(whi), "+r"(gi)
                  : "f"(q[g]), "f"(thr_s), "r"(rr[g * 4 + 0]), "r"(rr[g * 4 + 1]), "r"(rr[g * 4 + 2]),
                    "r"(rr[g * 4 + 3]), "r"(z1), "r"(z2), "r"(z3)
                  : "memory");
            }
          } else {
            // slow path: some thread of the warp is close to the end of its segment
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (q[g] > thr_s) {
                unsigned long long wp = ((unsigned long long)whi << 32) | wlo;
                if (wp < wend) {
                  CandRec* w = reinterpret_cast<CandRec*>(wp);
                  w->t = make_float4(__uint_as_float(rr[g * 4 + 0]), __uint_as_float(rr[g * 4 + 1]),
                                     __uint_as_float(rr[g * 4 + 2]), __uint_as_float(rr[g * 4 + 3]));
                  *reinterpret_cast<uint4*>(&w->g) = make_uint4(gi + g, 0u, 0u, 0u);
                  wp += sizeof(CandRec);
                  wlo = (uint32_t)wp; whi = (uint32_t)(wp >> 32);
                } else {
                  ++overflow;
                }
              }
            }
          }
        };

        // The thread's whole slice of the accumulator (128 columns) goes to registers first and the
        // TMEM buffer is released before any arithmetic: with only two accumulator buffers the
        // next-but-one MMA may start as soon as the epilogue has *read* this one, so the tensor
        // pipe waits for the epilogue's throughput only, never for its latency chain.
        if (PIPE == 0) {
          tc_ld32_issue(taddr, r0);
          tc_ld32_issue(taddr + 32, r1);
          tc_ld32_issue(taddr + 64, r2);
          tc_ld32_issue(taddr + 96, r3);
          tc_ld32_wait(r0);
          tc_ld32_wait(r1);
          tc_ld32_wait(r2);
          tc_ld32_wait(r3);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[buf]);
          process(r0, 0);
          process(r1, 1);
          process(r2, 2);
          process(r3, 3);
        } else {
          // columns 0..63 of this tile are in r0 / r1 (loaded while the previous tile's second half was processed)
          tc_ld32_issue(taddr + 64, r2);
          tc_ld32_issue(taddr + 96, r3);
          process(r0, 0);
          process(r1, 1);
          tc_ld32_wait(r2);
          tc_ld32_wait(r3);
          tc_fence_before();                              // the whole slice is in registers: release the buffer
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[buf]);
          const bool more = tcount + 1 < total_tiles;     // uniform over the CTA's epilogue warps
          if (more) {
            const uint32_t nt = tcount + 1;
            mbar_wait(&t_full[nt & 1], (nt >> 1) & 1);
            tc_fence_after();
            tc_ld32_issue(tlane + (nt & 1) * BN, r0);
            tc_ld32_issue(tlane + (nt & 1) * BN + 32, r1);
          }
          process(r2, 2);
          process(r3, 3);
          if (more) {
            tc_ld32_wait(r0);
            tc_ld32_wait(r1);
          }
        }
      }
      if (row_ok) {
        uint2* const info2 = reinterpret_cast<uint2*>(a.info + row) + half;
        *info2 = make_uint2(hopeless ? segcap + 1u : (uint32_t)(((((unsigned long long)whi << 32) | wlo) - seg) / sizeof(CandRec)) + overflow, __float_as_uint(smax));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // no CTA leaves while its peer may still write into it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

}  // namespace

bool assign_tc_supported(const spf_ctx* c, uint64_t m, uint32_t k, uint32_t ld) {
  return c->tma_encode != nullptr && ld % 4 == 0 && ld <= 1024u &&
         (int64_t)k >= c->params.tc_min_k && (int64_t)m >= c->params.tc_min_m;
}

int launch_assign_tc(spf_ctx* c, const float* Ptf, uint64_t m, const float* Ctf, uint32_t k, uint32_t ld,
                     const float* xnorm, const float* xres, const float* cext_pad, const float* d_cstat,
                     const float* seed, float factor, const CandBuf& cand, uint32_t eld) {
  CUtensorMap map_a, map_b, map_e;
  SPF_TRY(make_map_k128(c, &map_a, Ptf, m, ld, BM));
  SPF_TRY(make_map_k128(c, &map_b, Ctf, k, ld, BN / 2));   // each CTA of a pair fetches half a tile
  // K-extension rows: round_up(k, 256) x 8 floats
  SPF_TRY(make_map_ext(c, &map_e, cext_pad, (uint64_t)((k + BN - 1) / BN) * BN, BN));
  TcArgs a;
  a.m = (uint32_t)m; a.k = k; a.ld = ld; a.kb = (ld + BK - 1) / BK;
  a.eld = eld ? eld : ld;
  a.ntiles = (k + BN - 1) / BN;
  a.nrowblocks = (uint32_t)(ceil_div(m, 2 * BM) * 2);   // even: the CTAs of a pair walk the tiles in lockstep
  a.factor = factor;
  a.xnorm = xnorm; a.xres = xres; a.cstat = d_cstat; a.seed = seed;
  a.rec = cand.rec; a.info = cand.info; a.cap = cand.cap;
  unsigned grid = a.nrowblocks < (uint32_t)c->sm_count ? a.nrowblocks : (unsigned)c->sm_count & ~1u;
  if (c->params.tc_pipe) {
    SPF_CUDA(cudaFuncSetAttribute(assign_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    assign_tc_kernel<1><<<grid, NUM_THREADS, SMEM_TOTAL, c->stream>>>(map_a, map_b, map_e, a);
  } else {
    SPF_CUDA(cudaFuncSetAttribute(assign_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    assign_tc_kernel<0><<<grid, NUM_THREADS, SMEM_TOTAL, c->stream>>>(map_a, map_b, map_e, a);
  }
  return check_launch(c, "assign_tc_kernel");
}

}  // namespace spf
