// ops.cu — centroid update, k-means++ rounds, bisect seed and per-pair distances.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include <string.h>

#include "comm.cuh"
#include "kernels.cuh"
#include "pairdist.cuh"
#include "tc_ptx.cuh"

using namespace spf;

struct spf_kmpp {
  spf_dataset* ds = nullptr;
  int metric = 0;
  uint64_t rounds = 0;       // centroids folded into mind so far
  uint64_t newest = 0;       // row of the newest centroid (not yet folded)
  bool pending = true;       // newest still has to be folded
  float* mind = nullptr;     // n running minimum distances
  double* block_sums = nullptr;
  int* bad = nullptr;
  uint64_t nblocks = 0;
  // device result slots: [0]=status [1]=chosen ; sums
  uint64_t* res = nullptr;
  float* d_sum = nullptr;
  double* d_total = nullptr;
  float last_sum = 0.f;
  double last_total = 0.0;
  float* d_vec = nullptr;    // sharded sessions: the newest centroid as an explicit vector (ld floats)
  // batched rounds (spf_kmpp_rounds): draws, picked rows and {stop, done} stay on the device
  double* d_u01 = nullptr;
  uint64_t* d_chosen = nullptr;
  int* d_ctl = nullptr;      // [0] stop flag, [1] rounds completed
  uint32_t batch_cap = 0;
  // device-resident sharded rounds (spf_kmpp_rounds_sharded): exchange buffers, sized for `sh_world` ranks
  bool vec_pending = false;  // d_vec holds a centroid that is not folded yet
  int sh_world = 0;
  float* sh_sums = nullptr;      // world local f32 sums, rank order
  double* sh_tinfo = nullptr;    // world x {local f64 weight total, 1 = all local weights valid}
  int* sh_owner = nullptr;       // rank that picks this round, -1: no weighted pick possible
  double* sh_target = nullptr;   // u * total - totals of the lower ranks
  uint8_t* sh_cand = nullptr;    // (world + 1) x {u64 global row, ld floats}; slot `world` is this rank's
};

namespace spf {
namespace {

constexpr int WBLOCK = 1024;   // weights per block in the k-means++ pick

// ---- compute_mean (src/clustering/utils.rs:5-15): row-by-row f32 sum in member order, then a
// true division by m.  One CTA per cluster, one thread per 4 consecutive dimensions.  The sum of
// a dimension is a strictly sequential chain over the members, so the only parallelism inside a
// cluster is memory-level: every thread keeps a ring of MEAN_RING rows of its own 16-byte column
// in flight with cp.async (it consumes only what it copied itself, so no barrier is needed) and
// the chain never waits for a gather.
template <int RING>
__global__ void cluster_mean_kernel(const float* __restrict__ X, uint32_t ld4, const uint64_t* __restrict__ offsets,
                                    const uint64_t* __restrict__ rows, float* __restrict__ means, int divide) {
  extern __shared__ __align__(16) unsigned char mean_raw[];
  float4* buf = reinterpret_cast<float4*>(mean_raw);       // [RING][blockDim.x]
  const uint32_t c = blockIdx.x;
  const uint64_t b = offsets[c], e = offsets[c + 1];
  const float4* X4 = reinterpret_cast<const float4*>(X);
  for (uint32_t col0 = 0; col0 < ld4; col0 += blockDim.x) {
    const uint32_t col = col0 + threadIdx.x;
    const bool ok = col < ld4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // member indices are fetched 32 at a time (lane l of every warp holds rows[blk*32 + l], narrowed
    // to 32 bits: n < 2^32) one block ahead of their use, so the gather address never waits for
    // an index load
    const int lane = threadIdx.x & 31;
    auto load_idx = [&](uint64_t blk) {
      const uint64_t t = b + blk * 32 + lane;
      return t < e ? (uint32_t)rows[t] : 0u;
    };
    uint32_t idx_cur = load_idx(0), idx_nxt = load_idx(1);
    uint64_t idx_blk = 0;                                     // block idx_cur belongs to
    // rows are copied and consumed in groups of MEAN_G (one commit / wait per group, loads of a
    // group issued back to back): the sum of a dimension is a serial chain over the cluster, so
    // instructions per row are what bounds the largest cluster
    constexpr int MEAN_G = 8;
    static_assert(RING % MEAN_G == 0 && 32 % MEAN_G == 0, "ring and index blocks hold whole groups");
    const uint64_t nrows = e - b;
    auto issue = [&](uint64_t rel) {                          // rows rel .. rel + MEAN_G - 1 (relative), rel % MEAN_G == 0
      if ((rel >> 5) != idx_blk) {                            // warp-uniform: advance to the next index block
        idx_cur = idx_nxt;
        idx_blk = rel >> 5;
        idx_nxt = load_idx(idx_blk + 1);
      }
      const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(&buf[(size_t)(rel % RING) * blockDim.x + threadIdx.x]);
      if (rel + MEAN_G <= nrows) {                            // full group: no per-row guard
#pragma unroll
        for (int g = 0; g < MEAN_G; ++g) {
          const uint32_t row = __shfl_sync(0xffffffffu, idx_cur, (int)((rel & 31) + g));
          if (ok) {
            const float4* src = X4 + (size_t)row * ld4 + col;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + g * blockDim.x * 16u), "l"(src) : "memory");
          }
        }
      } else {
#pragma unroll
        for (int g = 0; g < MEAN_G; ++g) {
          const uint32_t row = __shfl_sync(0xffffffffu, idx_cur, (int)((rel & 31) + g));
          if (rel + g < nrows && ok) {
            const float4* src = X4 + (size_t)row * ld4 + col;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + g * blockDim.x * 16u), "l"(src) : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");   // one group per MEAN_G rows
    };
    for (int i = 0; i < RING; i += MEAN_G) issue((uint64_t)i);
    for (uint64_t rel = 0; rel < nrows; rel += MEAN_G) {
      // RING / MEAN_G groups are in flight; the oldest has landed once at most RING / MEAN_G - 1 are pending
      asm volatile("cp.async.wait_group %0;" ::"n"(RING / MEAN_G - 1) : "memory");
      if (ok) {
        const float4* slot = &buf[(size_t)(rel % RING) * blockDim.x + threadIdx.x];
        if (rel + MEAN_G <= nrows) {
          float4 v[MEAN_G];
#pragma unroll
          for (int g = 0; g < MEAN_G; ++g) v[g] = slot[(size_t)g * blockDim.x];
#pragma unroll
          for (int g = 0; g < MEAN_G; ++g) {
            acc.x = __fadd_rn(acc.x, v[g].x); acc.y = __fadd_rn(acc.y, v[g].y);
            acc.z = __fadd_rn(acc.z, v[g].z); acc.w = __fadd_rn(acc.w, v[g].w);
          }
        } else {
          for (int g = 0; rel + g < nrows; ++g) {
            const float4 v = slot[(size_t)g * blockDim.x];
            acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y);
            acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
          }
        }
      }
      issue(rel + RING);                                      // reuses the slots just consumed
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (ok) {
      if (divide && e > b) {
        const float fm = (float)(e - b);
        acc.x = __fdiv_rn(acc.x, fm); acc.y = __fdiv_rn(acc.y, fm);
        acc.z = __fdiv_rn(acc.z, fm); acc.w = __fdiv_rn(acc.w, fm);
      }
      reinterpret_cast<float4*>(means)[(size_t)c * ld4 + col] = acc;
    }
  }
}

// ---- compute_mean for skewed cluster sizes: warp-specialised producer / consumer ------------
// High-dimensional data makes hub clusters (1M x 128 N(0,1), k = 4096: the largest cluster holds
// 90 000 of 5.7 M members), and the ordered sum of a cluster is a serial chain, so the time of the
// whole update is the time of the largest cluster.  One CTA of four warps per cluster:
//   producers  (4 warps, or 12 for hub clusters): stages of member rows, copied with cp.async (one
//              16-byte piece per lane and 128-bit column) into a ring in shared memory; the stage's
//              mbarrier completes when every producer lane's copies have landed
//              (cp.async.mbarrier.arrive.noinc); member indices are fetched three stages ahead.  A
//              warp sustains only ~10 gathered rows per microsecond (its copies in flight / memory
//              latency), so a hub cluster gets 12 producers and a ring of up to 192 KB
//   last warp  consumer: waits for a stage, adds its rows in member order — one LDS.128 + four
//              un-fused FADD per row and 128-bit column, i.e. the chain itself — releases the stage
// The chain runs at a few cycles per row instead of waiting for a gather round trip per group of
// rows, and small clusters still fill the machine (<= 64 KB of ring per CTA).  Same operations in the
// same order as cluster_mean_kernel.  (A TMA variant, one cp.async.bulk per 512-byte row, was
// measured slower: 83 ns per row against 56 for the single-warp kernel.)
constexpr int CSB_COLS = 8;            // 128-bit columns per lane: rows up to 32 * 8 * 4 = 1024 floats

// nprod producer warps + one consumer warp; only clusters with min_n <= size < max_n are processed
// (two launches: a light configuration for the many ordinary clusters, a deep one for the hubs).
// blockIdx.y selects a slice of w4 = ld4 / gridDim.y 128-bit columns: the ordered sum is a chain per
// DIMENSION, so a hub cluster is split over several CTAs by columns (each gathers only its 16 * w4
// bytes of every member row) without touching the order of the additions.
// Producer-side wait with a suspend-time hint: the warp sleeps in the barrier unit until the phase
// completes (or ~1 us passes) instead of polling.  ncu of the polling form (profiles/r02_ncu_update_v1.txt):
// 41 % of the hub launch's warp instructions were the 12 producer warps' poll loops, issued on the
// consumer warp's schedulers.  Bounded like tc::mbar_wait.
__device__ __forceinline__ void mbar_wait_suspended(uint64_t* bar, uint32_t parity) {
  if (tc::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(tc::smem_u32(bar)), "r"(parity), "r"(1000u)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}

// fast != 0: producers wait suspended and rows of exactly 32 128-bit columns (d = 125..128) are copied by a
// three-instruction loop body (the generic column loop compiled to ~55 instructions per row).
__global__ void __launch_bounds__(512)
cluster_sum_ws_kernel(const float* __restrict__ X, uint32_t ld4, const uint64_t* __restrict__ offsets,
                      const uint64_t* __restrict__ rows, float* __restrict__ out, int divide, uint32_t stage_rows,
                      uint32_t nstage, uint32_t nprod, uint64_t min_n, uint64_t max_n, int fast) {
  extern __shared__ __align__(128) unsigned char csb_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(csb_raw);          // [nstage]
  uint64_t* empty = full + nstage;                                 // [nstage]
  float4* ring = reinterpret_cast<float4*>(csb_raw + 1024);        // [nstage][stage_rows][w4]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t c = blockIdx.x;
  const uint64_t b = offsets[c], e = offsets[c + 1];
  const uint64_t n = e - b;
  if (n < min_n || n >= max_n) return;                            // the other launch owns this cluster
  const uint32_t w4 = ld4 / gridDim.y, col0 = blockIdx.y * w4;
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < nstage; ++s) { tc::mbar_init(&full[s], 32); tc::mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // 32-bit bookkeeping throughout (a cluster has < 2^32 members) and no division in the loops: the
  // stage index / phase pair is advanced incrementally (64-bit `it % nstage` cost more than the copies)
  const uint32_t n32 = (uint32_t)n;
  const uint32_t nit = (n32 + stage_rows - 1) / stage_rows;
  const float4* X4 = reinterpret_cast<const float4*>(X) + col0;
  const uint64_t* crow = rows + b;
  if ((uint32_t)warp < nprod) {
    // ---------------- producers: stage `it` is filled by warp it % nprod; lane l < stage_rows owns row l of
    // the stage.  The member indices are fetched three of the warp's stages ahead of their use, so the
    // dependent index load never sits in front of the copies. ------------------------------------------
    auto load_idx = [&](uint32_t it) -> uint32_t {          // dataset rows are < 2^32
      const uint32_t t = it * stage_rows + (uint32_t)lane;
      return ((uint32_t)lane < stage_rows && it < nit && t < n32) ? (uint32_t)crow[t] : 0u;
    };
    uint32_t idx0 = load_idx((uint32_t)warp), idx1 = load_idx((uint32_t)warp + nprod), idx2 = load_idx((uint32_t)warp + 2 * nprod);
    uint32_t s = (uint32_t)warp % nstage, ph = ((uint32_t)warp / nstage) & 1u;
    const bool narrow = w4 < 32 && (w4 & (w4 - 1)) == 0;   // several rows per copy instruction
    const uint32_t wsh = narrow ? (uint32_t)__ffs((int)w4) - 1u : 0u;
    for (uint32_t it = (uint32_t)warp; it < nit; it += nprod) {
      const uint32_t idx = idx0;
      idx0 = idx1;
      idx1 = idx2;
      idx2 = load_idx(it + 3 * nprod);
      const uint32_t r0 = it * stage_rows;
      const uint32_t cnt = (n32 - r0) < stage_rows ? (n32 - r0) : stage_rows;
      if (fast) mbar_wait_suspended(&empty[s], ph ^ 1);       // the consumer is done with this stage
      else tc::mbar_wait(&empty[s], ph ^ 1);
      const uint32_t st = tc::smem_u32(ring + (size_t)s * stage_rows * w4);
      if (fast && w4 == 32) {                                 // one 512-byte row per instruction: lane = column
        const float4* srcl = X4 + lane;
        const uint32_t dl = st + (uint32_t)lane * 16u;
#pragma unroll 4
        for (uint32_t r = 0; r < cnt; ++r) {
          const uint32_t row = __shfl_sync(0xffffffffu, idx, (int)r);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dl + r * 512u), "l"(srcl + (size_t)row * ld4) : "memory");
        }
      } else if (narrow) {
        const uint32_t pieces = cnt << wsh;
        for (uint32_t p0 = 0; p0 < pieces; p0 += 32) {
          const uint32_t p = p0 + (uint32_t)lane;
          const uint32_t r = p >> wsh, col = p & (w4 - 1u);
          const uint32_t row = __shfl_sync(0xffffffffu, idx, (int)(r & 31u));
          if (p < pieces)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(st + p * 16u), "l"(X4 + (size_t)row * ld4 + col) : "memory");
        }
      } else {
        for (uint32_t r = 0; r < cnt; ++r) {
          const uint32_t row = __shfl_sync(0xffffffffu, idx, (int)r);
          const float4* src = X4 + (size_t)row * ld4;
          for (uint32_t col = (uint32_t)lane; col < w4; col += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(st + (r * w4 + col) * 16u), "l"(src + col) : "memory");
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(&full[s])) : "memory");
      s += nprod;
      while (s >= nstage) { s -= nstage; ph ^= 1u; }
    }
  } else {
    // ------------------------------ consumer -------------------------------------------------
    float4 acc[CSB_COLS];
#pragma unroll
    for (int j = 0; j < CSB_COLS; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t s = 0, ph = 0;
    for (uint32_t it = 0; it < nit; ++it) {
      const uint32_t r0 = it * stage_rows;
      const uint32_t cnt = (n32 - r0) < stage_rows ? (n32 - r0) : stage_rows;
      tc::mbar_wait(&full[s], ph);
      const float4* st = ring + (size_t)s * stage_rows * w4 + lane;
      if (w4 <= 32) {                                          // one column per lane: groups of 8 rows, loads first
        if ((uint32_t)lane < w4) {
          // software pipeline over groups of 8 rows, two register sets in ping-pong: the next group's
          // LDS.128 are in flight while the current group's four FADD chains run
          uint32_t r = 0;
          float4 v[8], w[8];
          bool have_v = false;
          if (cnt >= 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = st[u * w4];
            have_v = true;
          }
          while (have_v) {
            const bool have_w = r + 16 <= cnt;
            if (have_w) {
#pragma unroll
              for (int u = 0; u < 8; ++u) w[u] = st[(r + 8 + u) * w4];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              acc[0].x = __fadd_rn(acc[0].x, v[u].x); acc[0].y = __fadd_rn(acc[0].y, v[u].y);
              acc[0].z = __fadd_rn(acc[0].z, v[u].z); acc[0].w = __fadd_rn(acc[0].w, v[u].w);
            }
            r += 8;
            if (!have_w) break;
            have_v = r + 16 <= cnt;
            if (have_v) {
#pragma unroll
              for (int u = 0; u < 8; ++u) v[u] = st[(r + 8 + u) * w4];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              acc[0].x = __fadd_rn(acc[0].x, w[u].x); acc[0].y = __fadd_rn(acc[0].y, w[u].y);
              acc[0].z = __fadd_rn(acc[0].z, w[u].z); acc[0].w = __fadd_rn(acc[0].w, w[u].w);
            }
            r += 8;
          }
          for (; r < cnt; ++r) {
            const float4 t = st[r * w4];
            acc[0].x = __fadd_rn(acc[0].x, t.x); acc[0].y = __fadd_rn(acc[0].y, t.y);
            acc[0].z = __fadd_rn(acc[0].z, t.z); acc[0].w = __fadd_rn(acc[0].w, t.w);
          }
        }
      } else {
        for (uint32_t r = 0; r < cnt; ++r) {
#pragma unroll
          for (int j = 0; j < CSB_COLS; ++j) {
            if ((uint32_t)lane + 32u * j < w4) {
              const float4 v = st[r * w4 + 32u * j];
              acc[j].x = __fadd_rn(acc[j].x, v.x); acc[j].y = __fadd_rn(acc[j].y, v.y);
              acc[j].z = __fadd_rn(acc[j].z, v.z); acc[j].w = __fadd_rn(acc[j].w, v.w);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&empty[s]);
      if (++s == nstage) { s = 0; ph ^= 1u; }
    }
#pragma unroll
    for (int j = 0; j < CSB_COLS; ++j) {
      const uint32_t col = (uint32_t)lane + 32u * j;
      if (col < w4) {
        float4 a4 = acc[j];
        if (divide && n > 0) {
          const float fm = (float)n;
          a4.x = __fdiv_rn(a4.x, fm); a4.y = __fdiv_rn(a4.y, fm); a4.z = __fdiv_rn(a4.z, fm); a4.w = __fdiv_rn(a4.w, fm);
        }
        reinterpret_cast<float4*>(out)[(size_t)c * ld4 + col0 + col] = a4;
      }
    }
  }
}

// per-cluster ordered row sums (divide != 0: the mean)
int launch_cluster_mean(spf_ctx* c, const float* X, uint32_t ld, const uint64_t* d_offsets, const uint64_t* d_rows,
                        uint32_t k, float* means, int divide) {
  cudaStream_t st = c->stream;

  const uint32_t ld4 = ld / 4;
  if (ld4 <= 32 * CSB_COLS) {
    SPF_CUDA(cudaFuncSetAttribute(cluster_sum_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024));
    const uint64_t hub = c->params.sum_hub > 0 ? (uint64_t)c->params.sum_hub : 8192;   // clusters at least this large get the deep configuration
    // Every ring slot must always be filled by the same producer warp (the parity waits allow a
    // producer to be one phase ahead of the consumer, not two), so the number of producers divides
    // the number of stages.
    {   // hub clusters: 12 producers, the same stages, up to 192 KB of ring (one CTA per SM), optionally split
        // by columns (sum_slices); the CTAs of ordinary clusters return at once.  Runs on the context's auxiliary stream next to the launch
        // below (disjoint clusters), so the hubs' serial chains overlap the ordinary clusters.
      // (measured: slicing does NOT pay — random 128-byte pieces reach 1.2 TB/s of HBM against 2.7 TB/s for
      // whole 384..512-byte rows, and the gather, not the chain, bounds the launch; it stays a parameter)
      uint32_t nslice = 1;
      if (c->params.sum_slices > 0) nslice = (uint32_t)c->params.sum_slices;
      if (nslice < 1 || ld4 % nslice != 0) nslice = 1;
      const uint32_t w4 = ld4 / nslice;
      uint32_t stage_rows = (16u * 1024u) / (w4 * 16u);
      if (stage_rows > 32) stage_rows = 32;
      if (stage_rows < 1) stage_rows = 1;
      const size_t stage_bytes = (size_t)stage_rows * w4 * 16;
      uint32_t nstage = (uint32_t)((192u * 1024u) / stage_bytes);
      if (nstage > 12) nstage = 12;
      if (nstage < 2) nstage = 2;
      const uint32_t nprod = nstage;                                   // one fixed slot per producer warp
      const size_t smem = 1024 + (size_t)nstage * stage_bytes;
      if (!c->aux_stream) SPF_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
      if (!c->aux_ev[0]) {
        SPF_CUDA(cudaEventCreateWithFlags(&c->aux_ev[0], cudaEventDisableTiming));
        SPF_CUDA(cudaEventCreateWithFlags(&c->aux_ev[1], cudaEventDisableTiming));
      }
      SPF_CUDA(cudaEventRecord(c->aux_ev[0], st));                     // inputs are ready on the main stream
      SPF_CUDA(cudaStreamWaitEvent(c->aux_stream, c->aux_ev[0], 0));
      cluster_sum_ws_kernel<<<dim3(k, nslice), (nprod + 1) * 32, smem, c->aux_stream>>>(X, ld4, d_offsets, d_rows, means, divide,
                                                                                      stage_rows, nstage, nprod, hub, ~0ull, c->params.sum_fast);
      SPF_TRY(check_launch(c, "cluster_sum_ws_kernel"));
      SPF_CUDA(cudaEventRecord(c->aux_ev[1], c->aux_stream));
    }
    {   // ordinary clusters: 4 producers, 8 stages of up to 16 rows and <= 8 KB (64 KB of ring, 3 CTAs per SM)
      uint32_t stage_rows = (8u * 1024u) / (ld4 * 16u);
      if (stage_rows > 32) stage_rows = 32;
      if (stage_rows < 1) stage_rows = 1;
      const size_t stage_bytes = (size_t)stage_rows * ld4 * 16;
      const uint32_t nstage = 8, nprod = 4;
      const size_t smem = 1024 + (size_t)nstage * stage_bytes;       // <= 129 KB for rows of 1024 floats
      cluster_sum_ws_kernel<<<k, (nprod + 1) * 32, smem, st>>>(X, ld4, d_offsets, d_rows, means, divide, stage_rows, nstage, nprod,
                                                              0ull, hub, c->params.sum_fast);
      SPF_TRY(check_launch(c, "cluster_sum_ws_kernel"));
    }
    SPF_CUDA(cudaStreamWaitEvent(st, c->aux_ev[1], 0));                // both halves done before anything downstream
    return SPF_OK;
  }
  unsigned threads = round_up(ld / 4, 32);
  if (threads > 1024) threads = 1024;
  const size_t row_bytes = (size_t)threads * 16;
  if (row_bytes * 128 <= 64 * 1024) {
    const size_t smem = row_bytes * 128;
    SPF_CUDA(cudaFuncSetAttribute(cluster_mean_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cluster_mean_kernel<128><<<k, threads, smem, st>>>(X, ld / 4, d_offsets, d_rows, means, divide);
  } else if (row_bytes * 32 <= 64 * 1024) {
    const size_t smem = row_bytes * 32;
    SPF_CUDA(cudaFuncSetAttribute(cluster_mean_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cluster_mean_kernel<32><<<k, threads, smem, st>>>(X, ld / 4, d_offsets, d_rows, means, divide);
  } else {
    const size_t smem = row_bytes * 8;     // <= 128 KB (rows of up to 1024 float4 per pass)
    SPF_CUDA(cudaFuncSetAttribute(cluster_mean_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cluster_mean_kernel<8><<<k, threads, smem, st>>>(X, ld / 4, d_offsets, d_rows, means, divide);
  }
  return check_launch(c, "cluster_mean_kernel");
}

__global__ void expand_cluster_ids_kernel(const uint64_t* __restrict__ offsets, uint32_t* __restrict__ cid) {
  const uint32_t c = blockIdx.x;
  for (uint64_t t = offsets[c] + threadIdx.x; t < offsets[c + 1]; t += blockDim.x) cid[t] = c;
}

// ---- medoid (hierarchical.rs:155-171): argmin over members of d(row, mean), strict <, leftmost
// wins, identity (0, +inf).  key = dist bits << 32 | position inside the member list.
// DIRECT: only the member rows are staged (warp_row_dist), the mean is read in place.
template <int METRIC, bool DIRECT>
__global__ void __launch_bounds__(PD_THREADS)
medoid_kernel(const float* __restrict__ X, uint32_t ld, const uint64_t* __restrict__ rows,
              const uint32_t* __restrict__ cid, const uint64_t* __restrict__ offsets,
              const float* __restrict__ means, uint64_t total, unsigned long long* __restrict__ keys) {
  __shared__ typename std::conditional<DIRECT, RowDistSmem, PairDistSmem>::type sm[PD_THREADS / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t nwarps = (uint64_t)gridDim.x * (PD_THREADS / 32);
  for (uint64_t base = ((uint64_t)blockIdx.x * (PD_THREADS / 32) + warp) * 32; base < total; base += nwarps * 32) {
    const uint64_t t = base + lane;
    const bool valid = t < total;
    const float* pa = nullptr;
    const float* pb = nullptr;
    uint32_t c = 0;
    if (valid) {
      c = cid[t];
      pa = X + (size_t)rows[t] * ld;
      pb = means + (size_t)c * ld;
    }
    float dv;
    if constexpr (DIRECT) dv = warp_row_dist<METRIC>(pa, pb, ld, sm[warp]);
    else dv = warp_pair_dist<METRIC>(pa, pb, ld, sm[warp]);
    if (valid && dv < __int_as_float(0x7f800000)) {
      const unsigned long long key = ((unsigned long long)__float_as_uint(dv) << 32) | (uint32_t)(t - offsets[c]);
      atomicMin(&keys[c], key);
    }
  }
}

__global__ void medoid_finalize_kernel(const unsigned long long* __restrict__ keys, const uint64_t* __restrict__ offsets,
                                       const uint64_t* __restrict__ rows, const uint64_t* __restrict__ old_rows,
                                       uint32_t k, uint64_t* __restrict__ out) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= k) return;
  const uint64_t b = offsets[c], e = offsets[c + 1];
  if (e == b) out[c] = old_rows[c];                                   // :146-149
  else if (keys[c] == ~0ull) out[c] = 0;                              // identity (0, +inf)
  else out[c] = rows[b + (keys[c] & 0xffffffffull)];
}

// Sharded build: per cluster the best local member for a given mean, as (distance, dataset row);
// (+inf, UINT64_MAX) when the shard holds no member of the cluster or no distance is < +inf.
__global__ void medoid_candidates_kernel(const unsigned long long* __restrict__ keys,
                                         const uint64_t* __restrict__ offsets, const uint64_t* __restrict__ rows,
                                         uint32_t k, float* __restrict__ dist, uint64_t* __restrict__ row) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= k) return;
  if (keys[c] == ~0ull) {
    dist[c] = __int_as_float(0x7f800000);
    row[c] = ~0ull;
  } else {
    dist[c] = __uint_as_float((uint32_t)(keys[c] >> 32));
    row[c] = rows[offsets[c] + (keys[c] & 0xffffffffull)];
  }
}

// ---- farthest point (hierarchical.rs:112-126): argmax over members != c1, strict >, identity
// (0, 0.0): only distances > 0 compete, the earliest maximum wins.
template <int METRIC>
__global__ void __launch_bounds__(PD_THREADS)
farthest_kernel(const float* __restrict__ X, uint32_t ld, const uint64_t* __restrict__ members, uint64_t m,
                const float* __restrict__ c1vec, uint64_t c1, unsigned long long* __restrict__ key) {
  __shared__ RowDistSmem sm[PD_THREADS / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t nwarps = (uint64_t)gridDim.x * (PD_THREADS / 32);
  for (uint64_t base = ((uint64_t)blockIdx.x * (PD_THREADS / 32) + warp) * 32; base < m; base += nwarps * 32) {
    const uint64_t t = base + lane;
    const bool valid = t < m && members[t] != c1;
    const float* pa = valid ? c1vec : nullptr;      // c1 as a dataset row or an explicit vector (c1 = UINT64_MAX)
    const float* pb = valid ? X + (size_t)members[t] * ld : nullptr;
    // the member row is staged, c1 is read in place by every lane (a broadcast); d(x, c1) has the bits of
    // d(c1, x) for all three metrics: the difference only changes sign before it is squared / made absolute
    const float dv = warp_row_dist<METRIC>(pb, pa, ld, sm[warp]);
    if (valid && dv > 0.0f) {
      const unsigned long long kk = ((unsigned long long)__float_as_uint(dv) << 32) | (0xffffffffu - (uint32_t)t);
      atomicMax(key, kk);
    }
  }
}

// ---- k-means++ ----------------------------------------------------------------------------
// mind[i] = min(mind[i], d(x_i, newest))  (hierarchical.rs:260-276 restated as a running min)
template <int METRIC>
__global__ void __launch_bounds__(PD_THREADS)
kmpp_update_kernel(const float* __restrict__ X, uint32_t ld, uint64_t n, const float* __restrict__ cvec, int first,
                   float* __restrict__ mind, const uint64_t* __restrict__ d_row, const int* __restrict__ stop) {
  __shared__ RowDistSmem sm[PD_THREADS / 32];
  if (stop && *stop) return;
  if (d_row) cvec = X + (size_t)d_row[0] * ld;            // batched rounds: the row the previous pick wrote
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t nwarps = (uint64_t)gridDim.x * (PD_THREADS / 32);
  for (uint64_t base = ((uint64_t)blockIdx.x * (PD_THREADS / 32) + warp) * 32; base < n; base += nwarps * 32) {
    const uint64_t i = base + lane;
    const bool valid = i < n;
    const float* pa = valid ? X + (size_t)i * ld : nullptr;
    const float* pb = valid ? cvec : nullptr;       // the newest centroid (a dataset row or an explicit vector), read in place
    const float dv = warp_row_dist<METRIC>(pa, pb, ld, sm[warp]);
    if (valid && (first || dv < mind[i])) mind[i] = dv;
  }
}

// The same update for ld <= 448 at HBM rate: a CTA stages 128 rows in shared memory with coalesced
// 16-byte cp.async copies (row pitch ld + 4 floats: conflict-free 128-bit reads per quarter warp),
// then thread r walks row r in dimension order — the reference's sequential f32 chain — against the
// centroid held in shared memory.
constexpr int KU_ROWS = 128;
template <int METRIC>
__global__ void __launch_bounds__(KU_ROWS)
kmpp_update_tiled_kernel(const float* __restrict__ X, uint32_t ld, uint64_t n, const float* __restrict__ cvec, int first,
                         float* __restrict__ mind, const uint64_t* __restrict__ d_row, const int* __restrict__ stop) {
  extern __shared__ __align__(16) float ku_smem[];
  if (stop && *stop) return;
  if (d_row) cvec = X + (size_t)d_row[0] * ld;            // batched rounds: the row the previous pick wrote
  const uint32_t ld4 = ld / 4, pitch4 = ld4 + 1;           // in float4 units
  float4* s_c = reinterpret_cast<float4*>(ku_smem);         // the centroid, ld4 float4
  float4* s_x = s_c + ld4;                                  // KU_ROWS x pitch4
  for (uint32_t c = threadIdx.x; c < ld4; c += KU_ROWS) s_c[c] = reinterpret_cast<const float4*>(cvec)[c];
  const uint64_t ntiles = (n + KU_ROWS - 1) / KU_ROWS;
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const uint64_t row0 = t * KU_ROWS;
    const uint32_t rows = (uint32_t)((n - row0) < (uint64_t)KU_ROWS ? (n - row0) : (uint64_t)KU_ROWS);
    __syncthreads();                                        // the previous tile is consumed (and s_c is written)
    const float4* src = reinterpret_cast<const float4*>(X) + row0 * ld4;
    for (uint32_t w = threadIdx.x; w < rows * ld4; w += KU_ROWS) {
      const uint32_t r = w / ld4, c = w - r * ld4;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_x + r * pitch4 + c);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + w) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < rows) {
      const float4* xr = s_x + threadIdx.x * pitch4;
      float acc = 0.0f;
#pragma unroll 4
      for (uint32_t c = 0; c < ld4; ++c) {
        const float4 a = xr[c], b = s_c[c];
        acc = dist_step<METRIC>(acc, a.x, b.x);
        acc = dist_step<METRIC>(acc, a.y, b.y);
        acc = dist_step<METRIC>(acc, a.z, b.z);
        acc = dist_step<METRIC>(acc, a.w, b.w);
      }
      const uint64_t i = row0 + threadIdx.x;
      if (first || acc < mind[i]) mind[i] = acc;
    }
  }
}

unsigned pd_grid(spf_ctx* c, uint64_t count);

// Launches the tiled kernel when the row fits its shared-memory tile, the generic one otherwise.
template <int METRIC>
int launch_kmpp_update(spf_ctx* c, const float* X, uint32_t ld, uint64_t n, const float* cvec, int first, float* mind,
                       const uint64_t* d_row = nullptr, const int* stop = nullptr) {
  const size_t smem = ((size_t)(ld / 4) + (size_t)KU_ROWS * (ld / 4 + 1)) * sizeof(float4);
  if (smem <= 72 * 1024) {                                  // three CTAs per SM
    SPF_CUDA(cudaFuncSetAttribute(kmpp_update_tiled_kernel<METRIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint64_t ntiles = (n + KU_ROWS - 1) / KU_ROWS;
    const uint64_t cap = (uint64_t)c->sm_count * 3 * 4;
    kmpp_update_tiled_kernel<METRIC><<<(unsigned)(ntiles < cap ? ntiles : cap), KU_ROWS, smem, c->stream>>>(X, ld, n, cvec, first, mind, d_row, stop);
  } else {
    kmpp_update_kernel<METRIC><<<pd_grid(c, n), PD_THREADS, 0, c->stream>>>(X, ld, n, cvec, first, mind, d_row, stop);
  }
  return check_launch(c, "kmpp_update_kernel");
}

// :278 `distances.iter().fold(F::zero(), |acc, &x| acc + x)` — a strictly sequential f32 sum.
// Only the add chain is serial: the other threads of the block stage the next tile in shared
// memory (double buffered) while thread 0 folds the current one, so the cost is the 4-cycle FADD
// latency per element, not a global-memory round trip per 16 elements.
constexpr int SEQ_TILE = 4096;       // floats per tile
constexpr int SEQ_THREADS = 256;
__global__ void __launch_bounds__(SEQ_THREADS) seq_sum_kernel(const float* __restrict__ v, uint64_t n, float* __restrict__ out,
                                                              const int* __restrict__ stop = nullptr) {
  __shared__ __align__(16) float buf[2][SEQ_TILE];
  if (stop && *stop) return;
  const uint64_t ntiles = (n + SEQ_TILE - 1) / SEQ_TILE;
  auto stage = [&](uint64_t t, int b) {            // threads 1.. load tile t (the fold only reads the valid part)
    const uint64_t base = t * SEQ_TILE;
    for (uint32_t i = threadIdx.x - 1; i < (uint32_t)SEQ_TILE; i += SEQ_THREADS - 1)
      buf[b][i] = base + i < n ? v[base + i] : 0.0f;
  };
  if (threadIdx.x != 0 && ntiles) stage(0, 0);
  __syncthreads();
  float acc = 0.0f;
  for (uint64_t t = 0; t < ntiles; ++t) {
    const int b = (int)(t & 1);
    if (threadIdx.x != 0) {
      if (t + 1 < ntiles) stage(t + 1, b ^ 1);
    } else {
      const uint64_t left = n - t * SEQ_TILE;
      const uint32_t cnt = left < (uint64_t)SEQ_TILE ? (uint32_t)left : (uint32_t)SEQ_TILE;
      const float4* b4 = reinterpret_cast<const float4*>(buf[b]);
      uint32_t i = 0;
#pragma unroll 4
      for (; i + 4 <= cnt; i += 4) {
        const float4 x = b4[i / 4];
        acc = __fadd_rn(acc, x.x); acc = __fadd_rn(acc, x.y); acc = __fadd_rn(acc, x.z); acc = __fadd_rn(acc, x.w);
      }
      for (; i < cnt; ++i) acc = __fadd_rn(acc, buf[b][i]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = acc;
}

// The same sequential f32 sum, bit for bit, without a serial chain over the elements.
//
// For finite x >= 0 the running sum acc only grows.  While acc stays inside one binade
// [2^e, 2^(e+1)) it is an integer multiple A of u = ulp(acc), and fl(acc + x) = u * (A + k) where k
// is x / u rounded to an integer (round-half-to-even on A + floor(x/u)).  k depends on acc only
// through the parity of A, so a run of elements is a two-state transducer (parity in -> sum of k,
// parity out); transducers compose associatively, which turns the fold into a scan.  The element
// whose addition reaches the next binade is added with the hardware FADD and the scan restarts
// behind it with the new ulp: the sum crosses ~25 binades, so a 1 M element fold costs ~150 block
// iterations instead of 1 M dependent adds.  Anything else (negative or non-finite input) drops to
// the serial fold from the current position, and so does a sum that is already infinite.
constexpr int FS_THREADS = 1024;
constexpr int FS_RUN = 16;                              // consecutive elements per thread
constexpr int FS_WINDOW = FS_THREADS * FS_RUN;
constexpr int FS_SMEM = (FS_WINDOW + FS_WINDOW / 32) * 4;   // dynamic shared memory of the staging buffer

struct FsSumm { unsigned long long k0, k1; uint32_t po; };   // po: bit p = parity out for parity in p

__device__ __forceinline__ FsSumm fs_compose(const FsSumm& a, const FsSumm& b) {   // a first, then b
  FsSumm r;
  const uint32_t m0 = a.po & 1u, m1 = (a.po >> 1) & 1u;
  r.k0 = a.k0 + (m0 ? b.k1 : b.k0);
  r.k1 = a.k1 + (m1 ? b.k1 : b.k0);
  r.po = ((b.po >> m0) & 1u) | (((b.po >> m1) & 1u) << 1);
  return r;
}

__global__ void __launch_bounds__(FS_THREADS) seq_sum_scan_kernel(const float* __restrict__ v, uint64_t n, float* __restrict__ out,
                                                                  const int* __restrict__ stop = nullptr) {
  extern __shared__ float fs_buf[];                     // FS_WINDOW + FS_WINDOW / 32 staged values
  __shared__ FsSumm s_lane[FS_THREADS];                 // per thread: its exclusive prefix inside the warp
  __shared__ FsSumm s_warp[FS_THREADS / 32];
  __shared__ unsigned long long s_wk[FS_THREADS / 32];  // prefix of the warps for the actual parity
  __shared__ uint32_t s_wp[FS_THREADS / 32];
  __shared__ unsigned long long s_cross, s_ktotal;
  __shared__ unsigned long long s_pos;
  __shared__ uint32_t s_acc;                            // bits of the running sum
  __shared__ int s_bad;
  const uint32_t LIMIT = 1u << 24;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (stop && *stop) return;
  if (threadIdx.x == 0) { s_pos = 0; s_acc = 0; s_bad = 0; }
  __syncthreads();
  while (true) {
    const uint64_t pos = s_pos;
    const uint32_t accb = s_acc;
    if (pos >= n || s_bad || (accb >> 23) == 255u) break;
    // the binade of acc: A = acc / u as an integer, ulp exponent from max(E, 1)
    const uint32_t Ea = accb >> 23;
    const uint32_t Eeff = Ea ? Ea : 1u;
    const uint32_t A0 = Ea ? ((accb & 0x7fffffu) | 0x800000u) : accb;
    // ---- my run: floor(x/u), rounding direction / tie flag per element
    uint32_t F[FS_RUN];
    uint32_t gt = 0, tie = 0;
    bool bad = false;
    // the window goes through shared memory: coalesced global loads, then every thread reads its
    // run (one pad word per 32 keeps the 16-float runs on distinct banks)
#pragma unroll
    for (int r = 0; r < FS_RUN; ++r) {
      const uint32_t w = (uint32_t)r * FS_THREADS + threadIdx.x;
      const uint64_t i = pos + w;
      fs_buf[w + (w >> 5)] = i < n ? v[i] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < FS_RUN; ++r) {
      const uint32_t w = threadIdx.x * FS_RUN + (uint32_t)r;
      const uint32_t xb = __float_as_uint(fs_buf[w + (w >> 5)]);
      const uint32_t Ex = (xb >> 23) & 0xffu;
      if ((xb >> 31) && (xb << 1)) bad = true;          // negative (−0 counts as 0)
      if (Ex == 255u) bad = true;                       // inf / NaN
      const uint32_t mx = Ex ? ((xb & 0x7fffffu) | 0x800000u) : (xb & 0x7fffffu);
      const uint32_t Exeff = Ex ? Ex : 1u;
      uint32_t f = 0;
      if (Exeff > Eeff) {
        f = LIMIT;                                      // x alone reaches the next binade
      } else if (Exeff == Eeff) {
        f = mx;
      } else {
        const uint32_t sh = Eeff - Exeff;
        if (sh <= 25u) {
          f = mx >> sh;
          const uint32_t rem = mx & ((1u << sh) - 1u), half = 1u << (sh - 1u);
          gt |= (rem > half ? 1u : 0u) << r;
          tie |= (rem == half ? 1u : 0u) << r;
        }
      }
      F[r] = f;
    }
    if (bad) s_bad = 1;
    FsSumm me;
    if (tie == 0) {                                     // no half-way case in the run: k does not depend on the parity
      unsigned long long K = (unsigned long long)__popc(gt);
#pragma unroll
      for (int r = 0; r < FS_RUN; ++r) K += F[r];
      me.k0 = K; me.k1 = K;
      me.po = (uint32_t)(K & 1ull) | ((uint32_t)((K + 1ull) & 1ull) << 1);
    } else {
      unsigned long long k[2];
      uint32_t po = 0;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        unsigned long long K = 0;
        uint32_t par = (uint32_t)p;
#pragma unroll
        for (int r = 0; r < FS_RUN; ++r) {
          const uint32_t kk = F[r] + (((tie >> r) & 1u) ? ((par + F[r]) & 1u) : ((gt >> r) & 1u));
          K += kk;
          par = (par + kk) & 1u;
        }
        k[p] = K;
        po |= par << p;
      }
      me.k0 = k[0]; me.k1 = k[1]; me.po = po;
    }
    if (threadIdx.x == 0) s_cross = ~0ull;
    __syncthreads();
    if (s_bad) break;
    // ---- exclusive prefixes: a shuffle scan of the transducers inside each warp, then over the warps
    {
      FsSumm inc = me;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        FsSumm l;
        l.k0 = __shfl_up_sync(0xffffffffu, inc.k0, d);
        l.k1 = __shfl_up_sync(0xffffffffu, inc.k1, d);
        l.po = __shfl_up_sync(0xffffffffu, inc.po, d);
        if (lane >= d) inc = fs_compose(l, inc);
      }
      FsSumm ex;                                           // exclusive = inclusive of the lane before
      ex.k0 = __shfl_up_sync(0xffffffffu, inc.k0, 1);
      ex.k1 = __shfl_up_sync(0xffffffffu, inc.k1, 1);
      ex.po = __shfl_up_sync(0xffffffffu, inc.po, 1);
      if (lane == 0) { ex.k0 = 0; ex.k1 = 0; ex.po = 2u; }  // identity: parity out = parity in
      s_lane[threadIdx.x] = ex;
      if (lane == 31) s_warp[warp] = inc;
    }
    __syncthreads();
    if (warp == 0) {                                        // the 32 warp totals: one more shuffle scan
      static_assert(FS_THREADS / 32 == 32, "one lane per warp total");
      FsSumm inc = s_warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        FsSumm l;
        l.k0 = __shfl_up_sync(0xffffffffu, inc.k0, d);
        l.k1 = __shfl_up_sync(0xffffffffu, inc.k1, d);
        l.po = __shfl_up_sync(0xffffffffu, inc.po, d);
        if (lane >= d) inc = fs_compose(l, inc);
      }
      FsSumm ex;
      ex.k0 = __shfl_up_sync(0xffffffffu, inc.k0, 1);
      ex.k1 = __shfl_up_sync(0xffffffffu, inc.k1, 1);
      ex.po = __shfl_up_sync(0xffffffffu, inc.po, 1);
      if (lane == 0) { ex.k0 = 0; ex.k1 = 0; ex.po = 2u; }
      const uint32_t p0 = A0 & 1u;                          // evaluated for the parity the window starts with
      s_wk[lane] = p0 ? ex.k1 : ex.k0;
      s_wp[lane] = (ex.po >> p0) & 1u;
      if (lane == 31) s_ktotal = p0 ? inc.k1 : inc.k0;
    }
    __syncthreads();
    // ---- second walk with the actual A: the first element whose addition reaches the next binade
    {
      const uint32_t wp = s_wp[warp];
      const FsSumm pre = s_lane[threadIdx.x];
      unsigned long long A = (unsigned long long)A0 + s_wk[warp] + (wp ? pre.k1 : pre.k0);
      // k >= 0, so the run can only reach the next binade if its end does
      const unsigned long long mine = ((uint32_t)A & 1u) ? me.k1 : me.k0;
      if (A < LIMIT && A + mine >= LIMIT) {
#pragma unroll
      for (int r = 0; r < FS_RUN; ++r) {
        const uint32_t kk = F[r] + (((tie >> r) & 1u) ? (((uint32_t)A + F[r]) & 1u) : ((gt >> r) & 1u));
        if (A + kk >= LIMIT) {
          if (A < LIMIT)                                  // prefixes behind an earlier crossing are meaningless
            atomicMin(&s_cross, ((unsigned long long)(threadIdx.x * FS_RUN + r) << 32) | A);
          break;
        }
        A += kk;
      }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long cr = s_cross;
      auto to_bits = [&](uint32_t A) { return A < 0x800000u ? A : ((Eeff << 23) | (A & 0x7fffffu)); };
      if (cr == ~0ull) {                                  // the whole window stays in this binade
        s_acc = to_bits((uint32_t)(A0 + s_ktotal));
        s_pos = pos + FS_WINDOW;
      } else {
        const uint32_t ci = (uint32_t)(cr >> 32);
        const float before = __uint_as_float(to_bits((uint32_t)(cr & 0xffffffffull)));
        s_acc = __float_as_uint(__fadd_rn(before, fs_buf[ci + (ci >> 5)]));
        s_pos = pos + ci + 1;
      }
    }
    __syncthreads();
  }
  __syncthreads();
  // serial tail (empty in the normal case): negative / non-finite input or an infinite sum
  if (threadIdx.x == 0) {
    float acc = __uint_as_float(s_acc);
    for (uint64_t i = s_pos; i < n; ++i) acc = __fadd_rn(acc, v[i]);
    out[0] = acc;
  }
}

int launch_seq_sum_scan(spf_ctx* c, const float* v, uint64_t n, float* out) {
  SPF_CUDA(cudaFuncSetAttribute(seq_sum_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM));
  seq_sum_scan_kernel<<<1, FS_THREADS, FS_SMEM, c->stream>>>(v, n, out, nullptr);
  return SPF_OK;
}

// ---- the same scan over a thread-block cluster ---------------------------------------------------
// One cluster of CS_MAXC (16, else 8) CTAs owns the fold: a window is C x 8192 elements, thread t of
// CTA r holds 16 consecutive elements in registers.  An iteration = decode the runs for the current
// binade, transducer scan inside the warp (shuffles), over the warps (shared memory) and over the
// CTAs (every warp reads the C CTA summaries through distributed shared memory after ONE cluster
// barrier), so every CTA knows the sum at the end of the window and which CTA — if any — holds the
// first element that reaches the next binade.  Only that CTA walks its runs a second time and
// publishes (new sum, restart position) behind a second cluster barrier.  A restart keeps the
// window: the elements stay in registers and the ones in front of the restart position are masked
// to +0 (the identity), so a crossing costs no memory traffic.  1 M elements: 8 windows + ~20
// crossings = ~28 iterations of ~2 us instead of ~150 single-CTA iterations (0.54 ms -> see
// profiles/r02_experiment_notes.md).  Same bits as the serial chain for every input; the serial
// tail (negative / non-finite input, infinite sum) runs on CTA 0 from the current position.
constexpr int CS_THREADS = 512;
constexpr int CS_RUN = 16;
constexpr int CS_SLICE = CS_THREADS * CS_RUN;
constexpr int CS_MAXC = 16;
constexpr int CS_WARPS = CS_THREADS / 32;
constexpr int CS_PREFIX = 4096;      // elements folded by the plain add chain before the scan starts
constexpr uint32_t CS_SAT = 1u << 26;

// A run's transducer in 8 bytes: k[p] = sum of the rounded quotients when the run is entered with
// parity p, saturated at 2^26 (anything >= 2^24 means "reached the next binade"; sums in front of
// the first crossing stay exact).  The parity behind the run is (p + k[p]) & 1, so it needs no field.
struct CsT { uint32_t k0, k1; };
__device__ __forceinline__ CsT cs_compose(const CsT& a, const CsT& b) {   // a first, then b
  CsT r;
  r.k0 = min(a.k0 + ((a.k0 & 1u) ? b.k1 : b.k0), CS_SAT);
  r.k1 = min(a.k1 + ((a.k1 & 1u) ? b.k0 : b.k1), CS_SAT);
  return r;
}
__device__ __forceinline__ CsT cs_shfl_up(const CsT& a, int d) {
  CsT l;
  l.k0 = __shfl_up_sync(0xffffffffu, a.k0, d);
  l.k1 = __shfl_up_sync(0xffffffffu, a.k1, d);
  return l;
}
struct CsRes { unsigned long long start; uint32_t acc, pad; };

__device__ __forceinline__ uint32_t cs_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cs_cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cs_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cs_map(const void* p, uint32_t rank) {   // shared::cluster address of p in CTA `rank`
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
// 16 bytes into a peer CTA's shared memory; the peer's mbarrier counts them (complete_tx): the data is
// visible to whoever observes the phase flip with acquire.cluster — no cluster barrier, no MEMBAR.GPU
__device__ __forceinline__ void cs_send16(uint32_t remote_dst, uint4 v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               :: "r"(remote_dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void cs_bar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cs_bar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  }
}

__global__ void __launch_bounds__(CS_THREADS)
seq_sum_cluster_kernel(const float* __restrict__ v, uint64_t n, float* __restrict__ out, const int* __restrict__ stop,
                       unsigned long long* __restrict__ dbg) {
  __shared__ __align__(16) float s_pre[CS_PREFIX];
  __shared__ CsT s_warp[CS_WARPS];
  __shared__ __align__(16) uint4 s_in[2][CS_MAXC];        // {k0, k1, bad, -} of every CTA, by iteration parity
  __shared__ __align__(16) uint4 s_resin[2];              // {start lo, start hi, acc, -} from the crossing CTA
  __shared__ __align__(8) uint64_t s_bar[2], s_rbar[2];   // complete_tx barriers of the two exchanges
  __shared__ unsigned long long s_cross;
  __shared__ uint32_t s_acc0;
  if (stop && *stop) return;                              // uniform over the cluster (batched k-means++ rounds)
  const uint32_t LIMIT = 1u << 24;
  const uint32_t C = cs_cluster_size(), rank = cs_cluster_rank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t W = (uint64_t)C * CS_SLICE;
  if (threadIdx.x == 0) {
    tc::mbar_init(&s_bar[0], 1); tc::mbar_init(&s_bar[1], 1); tc::mbar_init(&s_rbar[0], 1); tc::mbar_init(&s_rbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cs_cluster_sync();                                      // every CTA's barriers exist before anybody sends
  // ---- the first CS_PREFIX elements by the plain add chain, in every CTA alike (no exchange): the
  // sum doubles about every time the position does, so half of the ~20 binade crossings of a fold
  // fall into this prefix and would each cost a scan iteration
  const uint32_t npre = n < (uint64_t)CS_PREFIX ? (uint32_t)n : (uint32_t)CS_PREFIX;
  for (uint32_t i = threadIdx.x; i < (uint32_t)CS_PREFIX; i += CS_THREADS) s_pre[i] = i < npre ? __ldcg(v + i) : 0.0f;
  __syncthreads();
  if (threadIdx.x == 0) {
    float acc = 0.0f;
    const float4* b4 = reinterpret_cast<const float4*>(s_pre);
    uint32_t i = 0;
#pragma unroll 4
    for (; i + 4 <= npre; i += 4) {
      const float4 x = b4[i / 4];
      acc = __fadd_rn(acc, x.x); acc = __fadd_rn(acc, x.y); acc = __fadd_rn(acc, x.z); acc = __fadd_rn(acc, x.w);
    }
    for (; i < npre; ++i) acc = __fadd_rn(acc, s_pre[i]);
    s_acc0 = __float_as_uint(acc);
  }
  __syncthreads();
  uint64_t wbase = 0, start = npre;
  uint32_t accb = s_acc0, it = 0, xit = 0;                // exchanges so far: summaries / crossing results
  bool fallback = (accb >> 23) == 255u && (accb & 0x7fffffu);   // NaN in the prefix: the chain carries on serially
  if (accb >> 31) fallback = true;                        // a negative sum so far: not the scan's domain
  uint32_t cur[CS_RUN];
  auto load_window = [&]() {
    const uint64_t i0 = wbase + (uint64_t)rank * CS_SLICE + (uint64_t)threadIdx.x * CS_RUN;
    if (i0 + CS_RUN <= n && (reinterpret_cast<uintptr_t>(v) & 15u) == 0) {
#pragma unroll
      for (int q = 0; q < CS_RUN / 4; ++q) {
        const uint4 x = __ldcg(reinterpret_cast<const uint4*>(v + i0) + q);
        cur[4 * q] = x.x; cur[4 * q + 1] = x.y; cur[4 * q + 2] = x.z; cur[4 * q + 3] = x.w;
      }
    } else {
#pragma unroll
      for (int r = 0; r < CS_RUN; ++r) cur[r] = i0 + r < n ? __float_as_uint(__ldcg(v + i0 + r)) : 0u;
    }
  };
  if (!fallback && start < n) load_window();
  if (dbg && rank == 0 && threadIdx.x == 0) { dbg[0] = C; dbg[2] = clock64(); }
  while (!fallback && start < n && (accb >> 23) != 255u) {
    if (dbg && rank == 0 && threadIdx.x == 0 && it < 60u) { dbg[3 + 2 * it] = clock64(); dbg[4 + 2 * it] = start; }
    const uint32_t Ea = accb >> 23;
    const uint32_t Eeff = Ea ? Ea : 1u;
    const uint32_t A0 = Ea ? ((accb & 0x7fffffu) | 0x800000u) : accb;
    const uint64_t i0 = wbase + (uint64_t)rank * CS_SLICE + (uint64_t)threadIdx.x * CS_RUN;
    // elements in front of `start` are already in the sum: masked to +0, the identity
    const uint32_t live = start <= i0 ? 0xffffu : (start >= i0 + CS_RUN ? 0u : (0xffffu << (uint32_t)(start - i0)) & 0xffffu);
    // ---- my run for this binade
    uint32_t F[CS_RUN];
    uint32_t gt = 0, tie = 0;
    bool bad = false;
#pragma unroll
    for (int r = 0; r < CS_RUN; ++r) {
      uint32_t xb = (live >> r) & 1u ? cur[r] : 0u;
      if (xb == 0x80000000u) xb = 0u;                     // -0 counts as 0
      if (xb > 0x7f7fffffu) bad = true;                   // negative, inf or NaN
      const uint32_t Ex = xb >> 23;
      const uint32_t mx = Ex ? ((xb & 0x7fffffu) | 0x800000u) : xb;
      const uint32_t Exeff = Ex ? Ex : 1u;
      uint32_t f = 0;
      if (Exeff > Eeff) {
        f = LIMIT;                                        // x alone reaches the next binade
      } else {
        const uint32_t sh = Eeff - Exeff;
        if (sh <= 25u) {
          f = mx >> sh;
          const uint32_t rem = mx & ((1u << sh) - 1u), half = (1u << sh) >> 1;   // sh = 0: rem = 0, half = 0 -> no flag
          gt |= (rem > half ? 1u : 0u) << r;
          tie |= ((rem == half && sh != 0u) ? 1u : 0u) << r;
        }
      }
      F[r] = f;
    }
    CsT me;
    if (tie == 0) {                                       // no half-way case in the run: k does not depend on the parity
      uint32_t K = (uint32_t)__popc(gt);
#pragma unroll
      for (int r = 0; r < CS_RUN; ++r) K += F[r];
      me.k0 = me.k1 = min(K, CS_SAT);
    } else {
      uint32_t k[2];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        uint32_t K = 0, par = (uint32_t)p;
#pragma unroll
        for (int r = 0; r < CS_RUN; ++r) {
          const uint32_t kk = F[r] + (((tie >> r) & 1u) ? ((par + F[r]) & 1u) : ((gt >> r) & 1u));
          K += kk;
          par = (par + kk) & 1u;
        }
        k[p] = min(K, CS_SAT);
      }
      me.k0 = k[0]; me.k1 = k[1];
    }
    const bool anybad = __syncthreads_or(bad ? 1 : 0) != 0;   // also orders the previous iteration's shared reads
    // ---- inside the warp
    CsT inc = me;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const CsT l = cs_shfl_up(inc, d);
      if (lane >= d) inc = cs_compose(l, inc);
    }
    CsT ex = cs_shfl_up(inc, 1);
    if (lane == 0) { ex.k0 = 0; ex.k1 = 0; }
    if (lane == 31) s_warp[warp] = inc;
    if (threadIdx.x == 0) s_cross = ~0ull;
    __syncthreads();
    // ---- over the warps of the CTA (every warp repeats the 16-entry scan: no second barrier)
    CsT winc = s_warp[lane & (CS_WARPS - 1)];
#pragma unroll
    for (int d = 1; d < CS_WARPS; d <<= 1) {
      const CsT l = cs_shfl_up(winc, d);
      if (lane >= d) winc = cs_compose(l, winc);
    }
    CsT wex;                                              // exclusive prefix of my warp inside the CTA
    {
      const int src = warp ? warp - 1 : 0;
      wex.k0 = __shfl_sync(0xffffffffu, winc.k0, src);
      wex.k1 = __shfl_sync(0xffffffffu, winc.k1, src);
      if (warp == 0) { wex.k0 = 0; wex.k1 = 0; }
    }
    // ---- over the CTAs: warp 0 sends this CTA's summary to every CTA (lane r -> CTA r, 16 bytes counted by
    // the receiver's barrier); every warp then reads the C summaries from its own shared memory
    if (warp == 0) {
      const uint32_t tk0 = __shfl_sync(0xffffffffu, winc.k0, CS_WARPS - 1), tk1 = __shfl_sync(0xffffffffu, winc.k1, CS_WARPS - 1);
      if (lane == 0) cs_bar_expect(&s_bar[it & 1u], C * 16u);
      if ((uint32_t)lane < C)
        cs_send16(cs_map(&s_in[it & 1u][rank], (uint32_t)lane), make_uint4(tk0, tk1, anybad ? 1u : 0u, 0u),
                  cs_map(&s_bar[it & 1u], (uint32_t)lane));
    }
    cs_bar_wait(&s_bar[it & 1u], (it >> 1) & 1u);
    CsT cinc; cinc.k0 = 0; cinc.k1 = 0;
    uint32_t cbad = 0;
    if ((uint32_t)lane < C) {
      const uint4 t = s_in[it & 1u][lane];
      cinc.k0 = t.x; cinc.k1 = t.y; cbad = t.z;
    }
#pragma unroll
    for (int d = 1; d < CS_MAXC; d <<= 1) {
      const CsT l = cs_shfl_up(cinc, d);
      if (lane >= d) cinc = cs_compose(l, cinc);
    }
    if (__any_sync(0xffffffffu, cbad != 0)) { fallback = true; break; }
    const uint32_t p0 = A0 & 1u;
    const uint32_t kinc = p0 ? cinc.k1 : cinc.k0;         // inclusive over the CTAs 0..lane, for the actual parity
    const uint32_t crossers = __ballot_sync(0xffffffffu, (uint32_t)lane < C && A0 + kinc >= LIMIT);
    const uint32_t ktotal = __shfl_sync(0xffffffffu, kinc, (int)C - 1);
    const uint32_t kpre = __shfl_sync(0xffffffffu, kinc, rank ? (int)rank - 1 : 0);
    ++it;
    if (crossers == 0) {                                  // the whole window stays in this binade
      const uint32_t A = A0 + ktotal;
      accb = A < 0x800000u ? A : ((Eeff << 23) | (A & 0x7fffffu));
      wbase += W;
      start = wbase;
      if (start < n) load_window();
      continue;
    }
    const uint32_t rc = (uint32_t)__ffs((int)crossers) - 1u;
    if (rc == rank) {                                     // the first crossing is in my CTA: second walk
      uint32_t A = A0 + (rank ? kpre : 0u);
      A += (A & 1u) ? wex.k1 : wex.k0;
      A = min(A, CS_SAT);
      A += (A & 1u) ? ex.k1 : ex.k0;
      A = min(A, CS_SAT);
      const uint32_t mine = (A & 1u) ? me.k1 : me.k0;
      unsigned long long mykey = ~0ull;                   // (index in the CTA, A in front of the element)
      uint32_t myx = 0;
      if (A < LIMIT && A + mine >= LIMIT) {
        bool open = true;
#pragma unroll
        for (int r = 0; r < CS_RUN; ++r) {
          const uint32_t kk = F[r] + (((tie >> r) & 1u) ? ((A + F[r]) & 1u) : ((gt >> r) & 1u));
          if (open && A + kk >= LIMIT) {
            mykey = ((unsigned long long)(threadIdx.x * CS_RUN + r) << 32) | A;
            myx = cur[r];                                 // live: a masked element is +0 and cannot cross
            open = false;
          }
          if (open) A += kk;
        }
        if (mykey != ~0ull) atomicMin(&s_cross, mykey);
      }
      __syncthreads();
      if (mykey != ~0ull && s_cross == mykey) {           // the owner of the crossing element adds it in hardware
        const uint32_t Ab = (uint32_t)(mykey & 0xffffffffull);
        const float before = __uint_as_float(Ab < 0x800000u ? Ab : ((Eeff << 23) | (Ab & 0x7fffffu)));
        const unsigned long long st = wbase + (uint64_t)rank * CS_SLICE + (uint32_t)(mykey >> 32) + 1;
        const uint4 t = make_uint4((uint32_t)st, (uint32_t)(st >> 32), __float_as_uint(__fadd_rn(before, __uint_as_float(myx))), 0u);
        for (uint32_t r = 0; r < C; ++r)                  // (new sum, restart position) to every CTA
          cs_send16(cs_map(&s_resin[xit & 1u], r), t, cs_map(&s_rbar[xit & 1u], r));
      }
    }
    if (threadIdx.x == 0) cs_bar_expect(&s_rbar[xit & 1u], 16u);
    cs_bar_wait(&s_rbar[xit & 1u], (xit >> 1) & 1u);
    {
      const uint4 t = s_resin[xit & 1u];
      start = ((unsigned long long)t.y << 32) | t.x;
      accb = t.z;
    }
    ++xit;
    if (start >= wbase + W) {
      wbase += W;
      if (start < n) load_window();
    }
  }
  cs_cluster_sync();                                      // nobody leaves while a peer may still read its shared memory
  if (dbg && rank == 0 && threadIdx.x == 0) { dbg[1] = it; if (it < 60u) dbg[3 + 2 * it] = clock64(); }
  if (rank == 0 && threadIdx.x == 0) {
    float acc = __uint_as_float(accb);
    if (fallback || (accb >> 23) == 255u)
      for (uint64_t i = start; i < n; ++i) acc = __fadd_rn(acc, v[i]);
    out[0] = acc;
  }
}

// cluster size the device can co-schedule (16 needs the non-portable opt-in), 0 = no cluster launch
int seq_sum_cluster_size() {
  static int cached = -1;
  if (cached >= 0) return cached;
  int best = 0;
  const char* force = getenv("SPF_SEQSUM_CLUSTER");        // experiments: force the cluster size (8 or 16)
  for (int C : {CS_MAXC, 8}) {
    if (force && atoi(force) != C) continue;
    if (C > 8 && cudaFuncSetAttribute(seq_sum_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)C);
    cfg.blockDim = dim3(CS_THREADS);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int ncl = 0;
    if (cudaOccupancyMaxActiveClusters(&ncl, seq_sum_cluster_kernel, &cfg) == cudaSuccess && ncl >= 1) { best = C; break; }
    cudaGetLastError();
  }
  cached = best;
  return best;
}

int launch_seq_sum_cluster(spf_ctx* c, const float* v, uint64_t n, float* out, const int* stop, unsigned long long* dbg = nullptr) {
  const int C = seq_sum_cluster_size();
  if (C == 0) return fail(SPF_E_CUDA, "thread-block clusters are not available on this device");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)C);
  cfg.blockDim = dim3(CS_THREADS);
  cfg.stream = c->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  SPF_CUDA(cudaLaunchKernelEx(&cfg, seq_sum_cluster_kernel, v, n, out, stop, dbg));
  return SPF_OK;
}

// the k-means++ rounds' sum: the cluster scan for long vectors, the single-CTA scan otherwise
int launch_kmpp_sum(spf_ctx* c, const float* v, uint64_t n, float* out, const int* stop) {
  if (n > (uint64_t)FS_WINDOW && seq_sum_cluster_size() > 0) return launch_seq_sum_cluster(c, v, n, out, stop);
  SPF_CUDA(cudaFuncSetAttribute(seq_sum_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM));
  seq_sum_scan_kernel<<<1, FS_THREADS, FS_SMEM, c->stream>>>(v, n, out, stop);
  return SPF_OK;
}

// Deterministic tree sum (fast mode: picks equal the reference's only up to near-ties).
__global__ void tree_sum_kernel(const float* __restrict__ v, uint64_t n, float* __restrict__ out, const int* __restrict__ stop = nullptr) {
  __shared__ double sm[1024];
  if (stop && *stop) return;
  double acc = 0.0;
  for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) acc += (double)v[i];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)sm[0];
}

__device__ __forceinline__ float kmpp_weight(float d, float denom) {
  return __fdiv_rn(__fmul_rn(d, d), denom);             // :281 (d * d) / max(sum, 1e-10)
}

// f64 block sums of the weights (rand 0.9 WeightedIndex accumulates f64 weights) + validity.
__global__ void __launch_bounds__(256)
kmpp_block_sums_kernel(const float* __restrict__ mind, uint64_t n, const float* __restrict__ d_sum,
                       double* __restrict__ block_sums, int* __restrict__ bad, const int* __restrict__ stop = nullptr) {
  __shared__ double sm[256];
  if (stop && *stop) return;
  const float denom = fmaxf(d_sum[0], 1e-10f);
  const uint64_t b0 = (uint64_t)blockIdx.x * WBLOCK;
  double acc = 0.0;
  bool isbad = false;
#pragma unroll
  for (int e = 0; e < WBLOCK / 256; ++e) {            // thread owns 4 consecutive weights
    const uint64_t i = b0 + (uint64_t)threadIdx.x * (WBLOCK / 256) + e;
    if (i < n) {
      const double w = (double)kmpp_weight(mind[i], denom);
      if (!(w >= 0.0)) isbad = true;
      acc += w;
    }
  }
  sm[threadIdx.x] = acc;
  if (isbad) *bad = 1;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) block_sums[blockIdx.x] = sm[0];
}

// cumulative_weights.partition_point(|w| w <= u) with u = u01 * total (rand 0.9 WeightedIndex).
// One CTA of 1024 threads, no serial walk over the vector: thread t adds its contiguous chunk of block
// sums in order, the chunk sums are scanned over the CTA (f64 shuffle scan in the warp, then over the
// 32 warp totals), the first chunk whose inclusive prefix exceeds u is walked by one thread (a chunk
// is one block sum at 1 M rows, ~100 at 100 M), and inside the block found the 1024 weights are
// scanned the same way.  The cumulative sums are therefore f64 sums in a fixed tree order — equal to
// the reference's sequential f64 prefix up to its last bits (documented near-tie deviation).
constexpr int PICK_THREADS = 1024;
static_assert(PICK_THREADS == WBLOCK, "one thread per weight of the block that holds the crossing");
// Device-side control of batched rounds (spf_kmpp_rounds): the draw of round `round` is read from
// u01[], the picked row goes to chosen[] (and to res[1], where the next round's update reads it), the
// first round whose weighted pick is impossible raises *stop and every later kernel of the batch
// returns at once.  All pointers null: single-round mode, the draw is the kernel argument.
struct KmppBatch {
  const double* u01 = nullptr;
  uint64_t* chosen = nullptr;
  int* stop = nullptr;
  uint32_t* done = nullptr;
  uint32_t round = 0;
  // sharded rounds: only the owning rank picks, with the target the owner kernel computed; a failed
  // pick is reported through the candidate exchange, not through *stop (the other ranks must agree)
  const int* owner = nullptr;
  const double* d_target = nullptr;
  int my_rank = 0;
};

// inclusive f64 scan over the 1024 threads of the CTA; *total = the last thread's value, *excl = the
// value of the thread in front (0 for the first)
__device__ __forceinline__ double pick_scan(double x, double* s_w, double* total, double* excl) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  const double up1 = __shfl_up_sync(0xffffffffu, x, 1);
  __syncthreads();                                        // the previous scan's warp totals are consumed
  if (lane == 31) s_w[warp] = x;
  __syncthreads();
  double w = s_w[lane];                                   // 32 warps: every warp scans the totals itself
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double y = __shfl_up_sync(0xffffffffu, w, d);
    if (lane >= d) w += y;
  }
  *total = __shfl_sync(0xffffffffu, w, 31);
  const double before = __shfl_sync(0xffffffffu, w, warp ? warp - 1 : 0);
  *excl = lane ? (warp ? before + up1 : up1) : (warp ? before : 0.0);
  return warp ? before + x : x;
}

__global__ void __launch_bounds__(PICK_THREADS)
kmpp_pick_kernel(const float* __restrict__ mind, uint64_t n, const float* __restrict__ d_sum,
                 const double* __restrict__ block_sums, uint64_t nblocks,
                 const int* __restrict__ bad, double u01, double target, uint64_t* __restrict__ res,
                 double* __restrict__ d_total, KmppBatch bt = KmppBatch()) {
  __shared__ double s_w[32];
  __shared__ unsigned long long s_first;
  __shared__ double s_cum;
  __shared__ unsigned long long s_b;
  if (bt.stop && *bt.stop) return;
  if (bt.owner) {
    if (*bt.owner != bt.my_rank) return;
    target = *bt.d_target;
  } else if (bt.u01) {
    u01 = bt.u01[bt.round];
  }
  // ---- chunk sums and their prefix
  const uint64_t chunk = (nblocks + PICK_THREADS - 1) / PICK_THREADS;
  const uint64_t c0 = (uint64_t)threadIdx.x * chunk;
  const uint64_t c1 = c0 + chunk < nblocks ? c0 + chunk : nblocks;
  double mine = 0.0;
  for (uint64_t i = c0; i < c1; ++i) mine += block_sums[i];
  if (threadIdx.x == 0) s_first = ~0ull;
  double total, excl;
  const double incl = pick_scan(mine, s_w, &total, &excl);
  if (threadIdx.x == 0) d_total[0] = total;
  if (*bad || total == 0.0 || !isfinite(total)) {         // uniform: every thread sees the same total
    if (threadIdx.x == 0) {
      res[0] = 1; res[1] = 0;
      if (bt.stop && !bt.owner) *bt.stop = 1;
    }
    return;
  }
  const double u = target >= 0.0 ? target : u01 * total;   // sharded pick: target already relative to this shard
  // ---- the chunk, then the block, where the cumulative sum crosses u (the last block takes what is left)
  if (c0 < nblocks && !(incl <= u)) atomicMin(&s_first, (unsigned long long)threadIdx.x);
  __syncthreads();
  const uint64_t nchunks = (nblocks + chunk - 1) / chunk;
  const uint64_t tc = s_first == ~0ull ? nchunks - 1 : s_first;
  if (threadIdx.x == tc) {
    double cum = excl;                                    // everything in front of my chunk
    uint64_t bi = c0;
    for (; bi < nblocks; ++bi) {                          // runs past the chunk only if rounding moved the crossing
      const double b = block_sums[bi];
      if (bi + 1 >= nblocks || !(cum + b <= u)) break;
      cum += b;
    }
    s_cum = cum;
    s_b = bi;
    s_first = ~0ull;
  }
  __syncthreads();
  // ---- the element inside that block
  const uint64_t lo = (uint64_t)s_b * WBLOCK;
  const double cum0 = s_cum;
  const float denom = fmaxf(d_sum[0], 1e-10f);
  const uint64_t i = lo + threadIdx.x;
  const double w = i < n ? (double)kmpp_weight(mind[i], denom) : 0.0;
  double tot2, ex2;
  const double e = cum0 + pick_scan(w, s_w, &tot2, &ex2);
  if (i + 1 < n && !(e <= u)) atomicMin(&s_first, (unsigned long long)i);
  __syncthreads();
  if (threadIdx.x != 0) return;
  uint64_t idx = n - 1;
  if (s_first != ~0ull) {
    idx = s_first;
  } else {                 // rounding pushed the crossing into a later block: continue sequentially
    double cum = cum0 + tot2;
    for (uint64_t j = lo + WBLOCK; j + 1 < n; ++j) {
      cum += (double)kmpp_weight(mind[j], denom);
      if (!(cum <= u)) { idx = j; break; }
    }
  }
  res[0] = 0;
  res[1] = idx;
  if (bt.chosen) { bt.chosen[bt.round] = idx; *bt.done = bt.round + 1u; }
}

// ---- device-resident row-sharded rounds (same arithmetic as the host-staged sequence
// fold_vector / weight_total / pick_local of sharded.py): small single-CTA kernels between the
// three all-gathers of a round.
__global__ void kmpp_combine_sums_kernel(const float* __restrict__ sums, int world, float* __restrict__ gsum, const int* __restrict__ stop) {
  if (*stop) return;
  float s = 0.0f;
  for (int r = 0; r < world; ++r) s = __fadd_rn(s, sums[r]);          // :278 over the ranks, in rank order
  gsum[0] = s;
}

// this shard's f64 weight total (the pick kernel's chunk sums + scan, bit for bit) and validity
__global__ void __launch_bounds__(PICK_THREADS)
kmpp_total_kernel(const double* __restrict__ block_sums, uint64_t nblocks, const int* __restrict__ bad,
                  double* __restrict__ my_tinfo, const int* __restrict__ stop) {
  __shared__ double s_w[32];
  if (*stop) return;
  const uint64_t chunk = (nblocks + PICK_THREADS - 1) / PICK_THREADS;
  const uint64_t c0 = (uint64_t)threadIdx.x * chunk;
  const uint64_t c1 = c0 + chunk < nblocks ? c0 + chunk : nblocks;
  double mine = 0.0;
  for (uint64_t i = c0; i < c1; ++i) mine += block_sums[i];
  double total, excl;
  pick_scan(mine, s_w, &total, &excl);
  if (threadIdx.x == 0) { my_tinfo[0] = total; my_tinfo[1] = *bad ? 0.0 : 1.0; }
}

__global__ void kmpp_owner_kernel(const double* __restrict__ tinfo, int world, const double* __restrict__ u01, uint32_t round,
                                  int* __restrict__ owner, double* __restrict__ target, int* __restrict__ stop) {
  if (*stop) return;
  double total = 0.0;
  bool ok = true;
  for (int r = 0; r < world; ++r) { total += tinfo[2 * r]; ok = ok && tinfo[2 * r + 1] == 1.0; }
  if (!ok || !(total > 0.0) || !isfinite(total)) { owner[0] = -1; *stop = 1; return; }   // the Err arm of WeightedIndex::new
  const double u = u01[round] * total;
  double prefix = 0.0;
  int own = world - 1;
  for (int r = 0; r < world - 1; ++r) {
    if (u < prefix + tinfo[2 * r]) { own = r; break; }
    prefix += tinfo[2 * r];
  }
  owner[0] = own;
  target[0] = u - prefix;
}

// this rank's candidate of the round: {global row, vector} on the owner, {~0, -} elsewhere
__global__ void kmpp_pack_kernel(const float* __restrict__ X, uint32_t ld, const uint64_t* __restrict__ res,
                                 const int* __restrict__ owner, int my_rank, uint64_t row_base,
                                 uint8_t* __restrict__ my_cand, const int* __restrict__ stop) {
  if (*stop) return;
  const bool mine = owner[0] == my_rank && res[0] == 0;
  if (threadIdx.x == 0) reinterpret_cast<unsigned long long*>(my_cand)[0] = mine ? row_base + res[1] : ~0ull;
  float* v = reinterpret_cast<float*>(my_cand + 16);
  if (mine)
    for (uint32_t j = threadIdx.x; j < ld; j += blockDim.x) v[j] = X[(size_t)res[1] * ld + j];
}

__global__ void kmpp_take_kernel(const uint8_t* __restrict__ cand, size_t cand_bytes, const int* __restrict__ owner,
                                 float* __restrict__ d_vec, uint32_t ld, uint64_t* __restrict__ chosen, uint32_t round,
                                 uint32_t* __restrict__ done, int* __restrict__ stop) {
  if (*stop) return;                                      // uniform: read before anybody writes it
  const uint8_t* slot = cand + (size_t)owner[0] * cand_bytes;
  const unsigned long long row = reinterpret_cast<const unsigned long long*>(slot)[0];
  __syncthreads();
  if (row == ~0ull) {                                     // the owner could not pick inside its shard
    if (threadIdx.x == 0) *stop = 1;
    return;
  }
  const float* v = reinterpret_cast<const float*>(slot + 16);
  for (uint32_t j = threadIdx.x; j < ld; j += blockDim.x) d_vec[j] = v[j];
  if (threadIdx.x == 0) { chosen[round] = row; *done = round + 1u; }
}

template <typename F>
int dispatch_metric(int metric, F&& f) {
  switch (metric) {
    case SPF_METRIC_EUCLIDEAN: return f(std::integral_constant<int, SPF_METRIC_EUCLIDEAN>());
    case SPF_METRIC_MANHATTAN: return f(std::integral_constant<int, SPF_METRIC_MANHATTAN>());
    case SPF_METRIC_CHEBYSHEV: return f(std::integral_constant<int, SPF_METRIC_CHEBYSHEV>());
  }
  return fail(SPF_E_INVALID, "unknown metric %d", metric);
}

unsigned pd_grid(spf_ctx* c, uint64_t count) {
  uint64_t blocks = ceil_div(count, PD_THREADS);
  const uint64_t cap = (uint64_t)c->sm_count * 12;   // 6 CTAs of 2 warps are resident per SM (35 KB of staging each)
  return (unsigned)(blocks > cap ? cap : (blocks ? blocks : 1));
}

// the medoid pass of all three entry points (params.medoid_direct picks the staging, results are the same bits)
int launch_medoid_kernel(spf_ctx* c, int metric, const float* X, uint32_t ld, const uint64_t* d_rows, const uint32_t* cid,
                         const uint64_t* d_offsets, const float* means, uint64_t total, unsigned long long* keys) {
  cudaStream_t st = c->stream;
  return dispatch_metric(metric, [&](auto M) {
    if (c->params.medoid_direct)
      medoid_kernel<decltype(M)::value, true><<<pd_grid(c, total), PD_THREADS, 0, st>>>(X, ld, d_rows, cid, d_offsets,
                                                                                          means, total, keys);
    else
      medoid_kernel<decltype(M)::value, false><<<pd_grid(c, total), PD_THREADS, 0, st>>>(X, ld, d_rows, cid, d_offsets,
                                                                                           means, total, keys);
    return check_launch(c, "medoid_kernel");
  });
}

}  // namespace

// ---- launchers shared with kmeans.cu (the device-resident row-sharded iteration) ---------------
// per-cluster f32 sums of the member rows in member order (divide != 0: the mean, utils.rs:13-14)
int launch_cluster_sums(spf_ctx* c, const float* X, uint32_t ld, const uint64_t* d_offsets, const uint64_t* d_rows,
                        uint32_t k, float* out, int divide) {
  return launch_cluster_mean(c, X, ld, d_offsets, d_rows, k, out, divide);
}

// keys[c] = min over the members of cluster c of (distance to means[c] bits << 32 | position in
// the member list), ~0 when the cluster is empty or no distance is < +inf (hierarchical.rs:155-171)
int launch_medoid_keys(spf_ctx* c, int metric, const float* X, uint32_t ld, const uint64_t* d_rows, uint64_t total,
                       const uint64_t* d_offsets, uint32_t k, const float* means, unsigned long long* keys) {
  cudaStream_t st = c->stream;
  SPF_CUDA(cudaMemsetAsync(keys, 0xff, (size_t)k * sizeof(unsigned long long), st));
  if (total == 0) return SPF_OK;
  DevBuf<uint32_t> cid;
  SPF_TRY(cid.alloc(st, total));
  expand_cluster_ids_kernel<<<k, 256, 0, st>>>(d_offsets, cid.p);
  SPF_TRY(check_launch(c, "expand_cluster_ids_kernel"));
  return launch_medoid_kernel(c, metric, X, ld, d_rows, cid.p, d_offsets, means, total, keys);
}

namespace {

// shared tail of the two spf_update_medoids entry points; d_offsets/d_rows are device arrays
int update_medoids_dev(spf_dataset* ds, int metric, const uint64_t* d_offsets, const uint64_t* d_rows,
                       uint64_t total, uint32_t k, const uint64_t* old_rows, uint64_t* new_rows,
                       float* means_out) {
  spf_ctx* c = ds->ctx;
  cudaStream_t st = c->stream;
  const uint32_t ld = ds->ld;
  DevBuf<float> means;
  DevBuf<uint32_t> cid;
  DevBuf<unsigned long long> keys;
  DevBuf<uint64_t> d_old, d_new;
  SPF_TRY(means.alloc(st, (size_t)k * ld));
  SPF_TRY(cid.alloc(st, total));
  SPF_TRY(keys.alloc(st, k));
  SPF_TRY(d_old.alloc(st, k));
  SPF_TRY(d_new.alloc(st, k));
  SPF_CUDA(cudaMemcpyAsync(d_old.p, old_rows, (size_t)k * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  SPF_CUDA(cudaMemsetAsync(keys.p, 0xff, (size_t)k * sizeof(unsigned long long), st));
  {
    KernelTimer t(c, "cluster_mean");
    SPF_TRY(launch_cluster_mean(c, ds->x, ld, d_offsets, d_rows, k, means.p, 1));
  }
  expand_cluster_ids_kernel<<<k, 256, 0, st>>>(d_offsets, cid.p);
  SPF_TRY(check_launch(c, "expand_cluster_ids_kernel"));
  if (total) {
    KernelTimer t(c, "medoid");
    SPF_TRY(launch_medoid_kernel(c, metric, ds->x, ld, d_rows, cid.p, d_offsets, means.p, total, keys.p));
  }
  medoid_finalize_kernel<<<(k + 255) / 256, 256, 0, st>>>(keys.p, d_offsets, d_rows, d_old.p, k, d_new.p);
  SPF_TRY(check_launch(c, "medoid_finalize_kernel"));
  SPF_CUDA(cudaMemcpyAsync(new_rows, d_new.p, (size_t)k * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  if (means_out)
    SPF_CUDA(cudaMemcpy2DAsync(means_out, (size_t)ds->d * sizeof(float), means.p, (size_t)ld * sizeof(float),
                               (size_t)ds->d * sizeof(float), k, cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  return SPF_OK;
}

}  // namespace
}  // namespace spf

extern "C" {

int spf_update_medoids(spf_dataset* ds, int metric, const uint64_t* offsets, const uint64_t* members,
                       uint32_t k, const uint64_t* old_rows, uint64_t* new_rows, float* means_out) {
  return spf::guarded([&]() -> int {
  if (!ds || !offsets || !old_rows || !new_rows) return fail(SPF_E_INVALID, "spf_update_medoids: NULL argument");
  if (metric < 0 || metric > 2) return fail(SPF_E_INVALID, "unknown metric %d", metric);
  if (k == 0) return SPF_OK;
  const uint64_t total = offsets[k];
  if (total && !members) return fail(SPF_E_INVALID, "members is NULL");
  for (uint32_t j = 0; j < k; ++j)
    if (offsets[j] > offsets[j + 1]) return fail(SPF_E_INVALID, "offsets must be non-decreasing");
  for (uint64_t t = 0; t < total; ++t)
    if (members[t] >= ds->n) return fail(SPF_E_INVALID, "member row %llu >= n", (unsigned long long)members[t]);
  spf_ctx* c = ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  DevBuf<uint64_t> d_off, d_rows;
  SPF_TRY(d_off.alloc(st, (size_t)k + 1));
  SPF_TRY(d_rows.alloc(st, total));
  SPF_CUDA(cudaMemcpyAsync(d_off.p, offsets, ((size_t)k + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  if (total) SPF_CUDA(cudaMemcpyAsync(d_rows.p, members, total * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  return update_medoids_dev(ds, metric, d_off.p, d_rows.p, total, k, old_rows, new_rows, means_out);
  });
}

int spf_update_medoids_from(spf_dataset* ds, int metric, const spf_assign_result* r,
                            const uint64_t* old_rows, uint64_t* new_rows, float* means_out) {
  return spf::guarded([&]() -> int {
  if (!ds || !r || !old_rows || !new_rows) return fail(SPF_E_INVALID, "spf_update_medoids_from: NULL argument");
  if (metric < 0 || metric > 2) return fail(SPF_E_INVALID, "unknown metric %d", metric);
  if (!r->has_csr) return fail(SPF_E_STATE, "the assign result has no CSR (SPF_ASSIGN_NO_CSR)");
  if (r->ctx != ds->ctx) return fail(SPF_E_INVALID, "result and dataset belong to different contexts");
  spf_ctx* c = ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  DevBuf<uint64_t> d_rows;
  SPF_TRY(d_rows.alloc(c->stream, r->total));
  SPF_TRY(assign_members_as_rows(r, d_rows.p));
  return update_medoids_dev(ds, metric, r->offsets, d_rows.p, r->total, r->k, old_rows, new_rows, means_out);
  });
}

// ---- sharded build: the two halves of update_centroids, so that the mean can be
// formed from the partial sums of all row shards and the medoid from the per-shard candidates ----
int spf_cluster_sums(spf_dataset* ds, const spf_assign_result* r, float* sums, uint64_t* counts) {
  return spf::guarded([&]() -> int {
  if (!ds || !r || !sums || !counts) return fail(SPF_E_INVALID, "spf_cluster_sums: NULL argument");
  if (!r->has_csr) return fail(SPF_E_STATE, "the assign result has no CSR (SPF_ASSIGN_NO_CSR)");
  if (r->ctx != ds->ctx) return fail(SPF_E_INVALID, "result and dataset belong to different contexts");
  spf_ctx* c = ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const uint32_t k = r->k, ld = ds->ld;
  DevBuf<uint64_t> d_rows;
  DevBuf<float> acc;
  SPF_TRY(d_rows.alloc(st, r->total));
  SPF_TRY(assign_members_as_rows(r, d_rows.p));
  SPF_TRY(acc.alloc(st, (size_t)k * ld));
  SPF_TRY(launch_cluster_mean(c, ds->x, ld, r->offsets, d_rows.p, k, acc.p, 0));
  std::vector<uint64_t> off((size_t)k + 1);
  SPF_CUDA(cudaMemcpy2DAsync(sums, (size_t)ds->d * sizeof(float), acc.p, (size_t)ld * sizeof(float),
                             (size_t)ds->d * sizeof(float), k, cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaMemcpyAsync(off.data(), r->offsets, ((size_t)k + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  for (uint32_t j = 0; j < k; ++j) counts[j] = off[j + 1] - off[j];
  return SPF_OK;
  });
}

int spf_medoid_candidates(spf_dataset* ds, int metric, const spf_assign_result* r, const float* means,
                          float* dist, uint64_t* row) {
  return spf::guarded([&]() -> int {
  if (!ds || !r || !means || !dist || !row) return fail(SPF_E_INVALID, "spf_medoid_candidates: NULL argument");
  if (metric < 0 || metric > 2) return fail(SPF_E_INVALID, "unknown metric %d", metric);
  if (!r->has_csr) return fail(SPF_E_STATE, "the assign result has no CSR (SPF_ASSIGN_NO_CSR)");
  if (r->ctx != ds->ctx) return fail(SPF_E_INVALID, "result and dataset belong to different contexts");
  spf_ctx* c = ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const uint32_t k = r->k, ld = ds->ld;
  DevBuf<uint64_t> d_rows, d_row;
  DevBuf<float> d_means, d_dist;
  DevBuf<uint32_t> cid;
  DevBuf<unsigned long long> keys;
  SPF_TRY(d_rows.alloc(st, r->total));
  SPF_TRY(assign_members_as_rows(r, d_rows.p));
  SPF_TRY(d_means.alloc(st, (size_t)k * ld));
  SPF_TRY(cid.alloc(st, r->total));
  SPF_TRY(keys.alloc(st, k));
  SPF_TRY(d_dist.alloc(st, k));
  SPF_TRY(d_row.alloc(st, k));
  if (ld != ds->d) SPF_CUDA(cudaMemsetAsync(d_means.p, 0, (size_t)k * ld * sizeof(float), st));
  SPF_CUDA(cudaMemcpy2DAsync(d_means.p, (size_t)ld * sizeof(float), means, (size_t)ds->d * sizeof(float),
                             (size_t)ds->d * sizeof(float), k, cudaMemcpyHostToDevice, st));
  SPF_CUDA(cudaMemsetAsync(keys.p, 0xff, (size_t)k * sizeof(unsigned long long), st));
  expand_cluster_ids_kernel<<<k, 256, 0, st>>>(r->offsets, cid.p);
  SPF_TRY(check_launch(c, "expand_cluster_ids_kernel"));
  if (r->total) {
    SPF_TRY(launch_medoid_kernel(c, metric, ds->x, ld, d_rows.p, cid.p, r->offsets, d_means.p, r->total, keys.p));
  }
  medoid_candidates_kernel<<<(k + 255) / 256, 256, 0, st>>>(keys.p, r->offsets, d_rows.p, k, d_dist.p, d_row.p);
  SPF_TRY(check_launch(c, "medoid_candidates_kernel"));
  SPF_CUDA(cudaMemcpyAsync(dist, d_dist.p, (size_t)k * sizeof(float), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaMemcpyAsync(row, d_row.p, (size_t)k * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  return SPF_OK;
  });
}

// Shared by spf_farthest (c1 is a dataset row) and spf_farthest_from (c1 is an explicit vector, the
// row-sharded bisect): packed (distance, earliest position) maximum over the members != skip_row.
static int farthest_impl(spf_dataset* ds, int metric, const float* h_c1vec, uint64_t c1_row, const uint64_t* members,
                         uint64_t m, unsigned long long* out_key) {
  if (metric < 0 || metric > 2) return fail(SPF_E_INVALID, "unknown metric %d", metric);
  if (m >= (1ull << 32)) return fail(SPF_E_INVALID, "m must be < 2^32");
  for (uint64_t t = 0; t < m; ++t)
    if (members[t] >= ds->n) return fail(SPF_E_INVALID, "member row %llu >= n", (unsigned long long)members[t]);
  spf_ctx* c = ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  *out_key = 0;
  if (m == 0) return SPF_OK;
  DevBuf<uint64_t> d_mem;
  DevBuf<unsigned long long> d_key;
  DevBuf<float> d_vec;
  SPF_TRY(d_mem.alloc(st, m));
  SPF_TRY(d_key.alloc(st, 1));
  const float* c1vec = nullptr;
  if (h_c1vec) {
    SPF_TRY(d_vec.alloc(st, ds->ld));
    SPF_CUDA(cudaMemsetAsync(d_vec.p, 0, ds->ld * sizeof(float), st));
    SPF_CUDA(cudaMemcpyAsync(d_vec.p, h_c1vec, ds->d * sizeof(float), cudaMemcpyHostToDevice, st));
    c1vec = d_vec.p;
  } else {
    c1vec = ds->x + (size_t)c1_row * ds->ld;
  }
  SPF_CUDA(cudaMemcpyAsync(d_mem.p, members, m * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  SPF_CUDA(cudaMemsetAsync(d_key.p, 0, sizeof(unsigned long long), st));
  {
    KernelTimer t(c, "farthest");
    SPF_TRY(dispatch_metric(metric, [&](auto M) {
      farthest_kernel<decltype(M)::value><<<pd_grid(c, m), PD_THREADS, 0, st>>>(ds->x, ds->ld, d_mem.p, m, c1vec, c1_row, d_key.p);
      return check_launch(c, "farthest_kernel");
    }));
  }
  SPF_CUDA(cudaMemcpyAsync(out_key, d_key.p, sizeof(*out_key), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  return SPF_OK;
}

int spf_farthest(spf_dataset* ds, int metric, uint64_t c1_row, const uint64_t* members, uint64_t m,
                 uint64_t* out_row) {
  if (!ds || !out_row || (m && !members)) return fail(SPF_E_INVALID, "spf_farthest: NULL argument");
  if (c1_row >= ds->n) return fail(SPF_E_INVALID, "c1_row >= n");
  unsigned long long key = 0;
  SPF_TRY(farthest_impl(ds, metric, nullptr, c1_row, members, m, &key));
  *out_row = key == 0 ? 0 : members[0xffffffffu - (uint32_t)(key & 0xffffffffull)];
  return SPF_OK;
}

int spf_farthest_from(spf_dataset* ds, int metric, const float* c1_vector, uint64_t skip_row, const uint64_t* members,
                      uint64_t m, float* out_dist, uint64_t* out_row) {
  if (!ds || !c1_vector || !out_dist || !out_row || (m && !members))
    return fail(SPF_E_INVALID, "spf_farthest_from: NULL argument");
  unsigned long long key = 0;
  SPF_TRY(farthest_impl(ds, metric, c1_vector, skip_row, members, m, &key));
  if (key == 0) { *out_dist = 0.0f; *out_row = ~0ull; return SPF_OK; }
  const uint32_t bits = (uint32_t)(key >> 32);
  memcpy(out_dist, &bits, sizeof(float));
  *out_row = members[0xffffffffu - (uint32_t)(key & 0xffffffffull)];
  return SPF_OK;
}

int spf_distance_pairs(spf_ctx* c, int metric, const float* a, const float* b, uint32_t d, uint64_t pairs,
                       float* out) {
  if (!c || !a || !b || !out) return fail(SPF_E_INVALID, "spf_distance_pairs: NULL argument");
  if (metric < 0 || metric > 2) return fail(SPF_E_INVALID, "unknown metric %d", metric);
  // ndarray-stats returns Err(EmptyInput) for d == 0 and the reference unwraps it (panic)
  if (d == 0) return fail(SPF_E_INVALID, "empty vectors (the reference panics on them)");
  if (pairs == 0) return SPF_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const uint32_t ld = round_up(d, 4);
  DevBuf<float> da, db, dout;
  SPF_TRY(da.alloc(st, (size_t)pairs * ld));
  SPF_TRY(db.alloc(st, (size_t)pairs * ld));
  SPF_TRY(dout.alloc(st, pairs));
  if (ld != d) {
    SPF_CUDA(cudaMemsetAsync(da.p, 0, (size_t)pairs * ld * sizeof(float), st));
    SPF_CUDA(cudaMemsetAsync(db.p, 0, (size_t)pairs * ld * sizeof(float), st));
  }
  SPF_CUDA(cudaMemcpy2DAsync(da.p, (size_t)ld * 4, a, (size_t)d * 4, (size_t)d * 4, pairs, cudaMemcpyHostToDevice, st));
  SPF_CUDA(cudaMemcpy2DAsync(db.p, (size_t)ld * 4, b, (size_t)d * 4, (size_t)d * 4, pairs, cudaMemcpyHostToDevice, st));
  SPF_TRY(launch_pair_dist(c, metric, da.p, ld, nullptr, db.p, ld, nullptr, UINT64_MAX, ld, pairs, dout.p));
  SPF_CUDA(cudaMemcpyAsync(out, dout.p, pairs * sizeof(float), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  return SPF_OK;
}

// The fold of hierarchical.rs:278 on its own (test hook for the scan-based kernel): mode 1 = scan,
// 2 = serial chain.  Both must return the same bits for any input.
int spf_seq_sum_f32(spf_ctx* c, const float* values, uint64_t n, int mode, float* out) {
  if (!c || !out || (n && !values)) return fail(SPF_E_INVALID, "spf_seq_sum_f32: NULL argument");
  if (mode < 1 || mode > 3) return fail(SPF_E_INVALID, "mode must be 1 (scan), 2 (serial) or 3 (cluster scan)");
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  DevBuf<float> d_v, d_out;
  SPF_TRY(d_v.alloc(st, n));
  SPF_TRY(d_out.alloc(st, 1));
  if (n) SPF_CUDA(cudaMemcpyAsync(d_v.p, values, n * sizeof(float), cudaMemcpyHostToDevice, st));
  {
    KernelTimer t(c, "seq_sum");
    if (mode == 1) SPF_TRY(launch_seq_sum_scan(c, d_v.p, n, d_out.p));
    else if (mode == 3) {
      DevBuf<unsigned long long> d_dbg;
      const bool dbg = getenv("SPF_SEQSUM_DEBUG") != nullptr;   // iteration count and clocks of the cluster scan on stderr
      if (dbg) {
        SPF_TRY(d_dbg.alloc(st, 128));
        SPF_CUDA(cudaMemsetAsync(d_dbg.p, 0, 128 * sizeof(unsigned long long), st));
      }
      SPF_TRY(launch_seq_sum_cluster(c, d_v.p, n, d_out.p, nullptr, dbg ? d_dbg.p : nullptr));
      if (dbg) {
        unsigned long long h[128];
        SPF_CUDA(cudaMemcpyAsync(h, d_dbg.p, sizeof(h), cudaMemcpyDeviceToHost, st));
        SPF_CUDA(cudaStreamSynchronize(st));
        fprintf(stderr, "seq_sum_cluster: C=%llu iterations=%llu prefix=%llu cyc;", h[0], h[1], h[3] - h[2]);
        for (unsigned i = 0; i < h[1] && i < 60; ++i) fprintf(stderr, " [%llu]%llu", h[4 + 2 * i], h[5 + 2 * i] - h[3 + 2 * i]);
        fprintf(stderr, "\n");
      }
    }
    else seq_sum_kernel<<<1, SEQ_THREADS, 0, st>>>(d_v.p, n, d_out.p);
    SPF_TRY(check_launch(c, "seq_sum kernel"));
  }
  SPF_CUDA(cudaMemcpyAsync(out, d_out.p, sizeof(float), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  return SPF_OK;
}

// ---- k-means++ session ----------------------------------------------------------------------
int spf_kmpp_begin(spf_dataset* ds, int metric, uint64_t first_row, spf_kmpp** out) {
  return spf::guarded([&]() -> int {
  if (!ds || !out) return fail(SPF_E_INVALID, "spf_kmpp_begin: NULL argument");
  *out = nullptr;
  if (metric < 0 || metric > 2) return fail(SPF_E_INVALID, "unknown metric %d", metric);
  if (first_row >= ds->n) return fail(SPF_E_INVALID, "first_row >= n");
  spf_ctx* c = ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  spf_kmpp* s = new (std::nothrow) spf_kmpp();
  if (!s) return fail(SPF_E_OOM, "out of host memory");
  s->ds = ds;
  s->metric = metric;
  s->newest = first_row;
  s->nblocks = ceil_div(ds->n, WBLOCK);
  cudaError_t e = cudaMalloc((void**)&s->mind, ds->n * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->block_sums, s->nblocks * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->bad, sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->res, 2 * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_sum, sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_total, sizeof(double));
  if (e != cudaSuccess) {
    spf_kmpp_free(s);
    return fail(SPF_E_OOM, "k-means++ session allocation failed: %s", cudaGetErrorString(e));
  }
  *out = s;
  return SPF_OK;
  });
}

// `count` rounds of hierarchical.rs:259-291 back to back on the device: the row a round picks is read
// by the next round's update kernel from device memory, so the host synchronises once per batch.
static int kmpp_rounds_locked(spf_kmpp* s, const double* u01, uint32_t count, uint64_t* chosen, uint32_t* done) {
  spf_dataset* ds = s->ds;
  spf_ctx* c = ds->ctx;
  cudaStream_t st = c->stream;
  const uint64_t n = ds->n;
  if (count > s->batch_cap) {
    if (s->d_u01) cudaFree(s->d_u01);
    if (s->d_chosen) cudaFree(s->d_chosen);
    s->d_u01 = nullptr; s->d_chosen = nullptr; s->batch_cap = 0;
    const uint32_t cap = count < 64u ? 64u : count;
    if (cudaMalloc((void**)&s->d_u01, (size_t)cap * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&s->d_chosen, (size_t)cap * sizeof(uint64_t)) != cudaSuccess)
      return fail(SPF_E_OOM, "k-means++ batch buffers: out of device memory");
    s->batch_cap = cap;
  }
  if (!s->d_ctl && cudaMalloc((void**)&s->d_ctl, 2 * sizeof(int)) != cudaSuccess)
    return fail(SPF_E_OOM, "k-means++ batch buffers: out of device memory");
  SPF_CUDA(cudaMemcpyAsync(s->d_u01, u01, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, st));
  SPF_CUDA(cudaMemsetAsync(s->d_ctl, 0, 2 * sizeof(int), st));
  SPF_CUDA(cudaMemsetAsync(s->bad, 0, sizeof(int), st));
  const uint64_t res0[2] = {0, s->newest};
  SPF_CUDA(cudaMemcpyAsync(s->res, res0, sizeof(res0), cudaMemcpyHostToDevice, st));   // pageable source: staged before the call returns
  const int* stop = s->d_ctl;
  for (uint32_t r = 0; r < count; ++r) {
    {
      KernelTimer t(c, "kmpp_update");
      const int first = (s->rounds == 0 && r == 0) ? 1 : 0;
      SPF_TRY(dispatch_metric(s->metric, [&](auto M) {
        return launch_kmpp_update<decltype(M)::value>(c, ds->x, ds->ld, n, ds->x, first, s->mind, s->res + 1, stop);
      }));
    }
    {
      KernelTimer t(c, "kmpp_sum");
      if (c->params.kmpp_exact_sum == 1) SPF_TRY(launch_kmpp_sum(c, s->mind, n, s->d_sum, stop));
      else if (c->params.kmpp_exact_sum == 2) seq_sum_kernel<<<1, SEQ_THREADS, 0, st>>>(s->mind, n, s->d_sum, stop);
      else tree_sum_kernel<<<1, 1024, 0, st>>>(s->mind, n, s->d_sum, stop);
      SPF_TRY(check_launch(c, "kmpp sum kernel"));
    }
    {
      KernelTimer t(c, "kmpp_pick");
      kmpp_block_sums_kernel<<<(unsigned)s->nblocks, 256, 0, st>>>(s->mind, n, s->d_sum, s->block_sums, s->bad, stop);
      SPF_TRY(check_launch(c, "kmpp_block_sums_kernel"));
      KmppBatch bt;
      bt.u01 = s->d_u01; bt.chosen = s->d_chosen; bt.stop = s->d_ctl; bt.done = reinterpret_cast<uint32_t*>(s->d_ctl + 1); bt.round = r;
      kmpp_pick_kernel<<<1, PICK_THREADS, 0, st>>>(s->mind, n, s->d_sum, s->block_sums, s->nblocks, s->bad, 0.0, -1.0, s->res, s->d_total, bt);
      SPF_TRY(check_launch(c, "kmpp_pick_kernel"));
    }
  }
  int ctl[2] = {0, 0};
  SPF_CUDA(cudaMemcpyAsync(ctl, s->d_ctl, sizeof(ctl), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaMemcpyAsync(chosen, s->d_chosen, (size_t)count * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaMemcpyAsync(&s->last_sum, s->d_sum, sizeof(float), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaMemcpyAsync(&s->last_total, s->d_total, sizeof(double), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  const uint32_t ok = (uint32_t)ctl[1];
  const bool failed = ctl[0] != 0;
  *done = ok;
  s->rounds += ok + (failed ? 1u : 0u);                   // the failing round folded its centroid too
  if (ok) s->newest = chosen[ok - 1];
  s->pending = !failed;
  return failed ? 1 : SPF_OK;
}

int spf_kmpp_round(spf_kmpp* s, double u01, uint64_t* chosen) {
  return spf::guarded([&]() -> int {
  if (!s || !chosen) return fail(SPF_E_INVALID, "spf_kmpp_round: NULL argument");
  if (!(u01 >= 0.0 && u01 < 1.0)) return fail(SPF_E_INVALID, "u01 must be in [0,1)");
  if (!s->pending) return fail(SPF_E_STATE, "previous round needs spf_kmpp_push() before the next one");
  spf_ctx* c = s->ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  uint32_t done = 0;
  uint64_t row = 0;
  const int rc = kmpp_rounds_locked(s, &u01, 1, &row, &done);
  if (rc != SPF_OK) return rc;   // 1: weighted pick impossible → host draws uniformly and pushes
  *chosen = row;
  return SPF_OK;
  });
}

int spf_kmpp_rounds(spf_kmpp* s, const double* u01, uint32_t count, uint64_t* chosen, uint32_t* done) {
  return spf::guarded([&]() -> int {
  if (!s || !u01 || !chosen || !done) return fail(SPF_E_INVALID, "spf_kmpp_rounds: NULL argument");
  *done = 0;
  if (count == 0) return SPF_OK;
  if (count > (1u << 20)) return fail(SPF_E_INVALID, "spf_kmpp_rounds: at most 2^20 rounds per call");
  for (uint32_t i = 0; i < count; ++i)
    if (!(u01[i] >= 0.0 && u01[i] < 1.0)) return fail(SPF_E_INVALID, "u01 must be in [0,1)");
  if (!s->pending) return fail(SPF_E_STATE, "previous round needs spf_kmpp_push() before the next one");
  spf_ctx* c = s->ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  return kmpp_rounds_locked(s, u01, count, chosen, done);
  });
}

// ---- row-sharded k-means++: hierarchical.rs:249-293 split at its reductions ----
int spf_kmpp_begin_sharded(spf_dataset* ds, int metric, spf_kmpp** out) {
  SPF_TRY(spf_kmpp_begin(ds, metric, 0, out));
  spf_kmpp* s = *out;
  s->pending = false;      // no centroid yet: the first one arrives through spf_kmpp_fold_vector
  if (cudaMalloc((void**)&s->d_vec, (size_t)ds->ld * sizeof(float)) != cudaSuccess) {
    spf_kmpp_free(s);
    *out = nullptr;
    return fail(SPF_E_OOM, "k-means++ session allocation failed");
  }
  return SPF_OK;
}

int spf_kmpp_fold_vector(spf_kmpp* s, const float* centroid, float* local_sum) {
  if (!s || !centroid || !local_sum || !s->d_vec) return fail(SPF_E_INVALID, "spf_kmpp_fold_vector: bad argument");
  spf_dataset* ds = s->ds;
  spf_ctx* c = ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const uint64_t n = ds->n;
  SPF_CUDA(cudaMemsetAsync(s->d_vec, 0, (size_t)ds->ld * sizeof(float), st));
  SPF_CUDA(cudaMemcpyAsync(s->d_vec, centroid, (size_t)ds->d * sizeof(float), cudaMemcpyHostToDevice, st));
  const int first = s->rounds == 0 ? 1 : 0;
  SPF_TRY(dispatch_metric(s->metric, [&](auto M) {
    return launch_kmpp_update<decltype(M)::value>(c, ds->x, ds->ld, n, s->d_vec, first, s->mind);
  }));
  s->rounds += 1;
  if (c->params.kmpp_exact_sum == 1) SPF_TRY(launch_kmpp_sum(c, s->mind, n, s->d_sum, nullptr));
    else if (c->params.kmpp_exact_sum == 2) seq_sum_kernel<<<1, SEQ_THREADS, 0, st>>>(s->mind, n, s->d_sum);
  else tree_sum_kernel<<<1, 1024, 0, st>>>(s->mind, n, s->d_sum);
  SPF_TRY(check_launch(c, "kmpp sum kernel"));
  SPF_CUDA(cudaMemcpyAsync(local_sum, s->d_sum, sizeof(float), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  return SPF_OK;
}

int spf_kmpp_weight_total(spf_kmpp* s, float global_sum, double* local_total) {
  if (!s || !local_total) return fail(SPF_E_INVALID, "spf_kmpp_weight_total: NULL argument");
  spf_dataset* ds = s->ds;
  spf_ctx* c = ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  SPF_CUDA(cudaMemcpyAsync(s->d_sum, &global_sum, sizeof(float), cudaMemcpyHostToDevice, st));
  SPF_CUDA(cudaMemsetAsync(s->bad, 0, sizeof(int), st));
  kmpp_block_sums_kernel<<<(unsigned)s->nblocks, 256, 0, st>>>(s->mind, ds->n, s->d_sum, s->block_sums, s->bad);
  SPF_TRY(check_launch(c, "kmpp_block_sums_kernel"));
  // total + validity through the pick kernel (target beyond the total: the pick itself is discarded)
  kmpp_pick_kernel<<<1, PICK_THREADS, 0, st>>>(s->mind, ds->n, s->d_sum, s->block_sums, s->nblocks, s->bad, 0.0, 0.0, s->res, s->d_total);
  SPF_TRY(check_launch(c, "kmpp_pick_kernel"));
  uint64_t res[2] = {0, 0};
  SPF_CUDA(cudaMemcpyAsync(res, s->res, sizeof(res), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaMemcpyAsync(local_total, s->d_total, sizeof(double), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  int bad = 0;
  SPF_CUDA(cudaMemcpy(&bad, s->bad, sizeof(int), cudaMemcpyDeviceToHost));
  return bad ? 1 : SPF_OK;   // 1: an invalid weight on this shard (the Err arm of WeightedIndex::new)
}

int spf_kmpp_pick_local(spf_kmpp* s, double target, uint64_t* row) {
  if (!s || !row) return fail(SPF_E_INVALID, "spf_kmpp_pick_local: NULL argument");
  if (!(target >= 0.0)) return fail(SPF_E_INVALID, "target must be >= 0");
  spf_dataset* ds = s->ds;
  spf_ctx* c = ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  // block sums and the global denominator are those of the preceding spf_kmpp_weight_total
  kmpp_pick_kernel<<<1, PICK_THREADS, 0, st>>>(s->mind, ds->n, s->d_sum, s->block_sums, s->nblocks, s->bad, 0.0, target, s->res, s->d_total);
  SPF_TRY(check_launch(c, "kmpp_pick_kernel"));
  uint64_t res[2] = {0, 0};
  SPF_CUDA(cudaMemcpyAsync(res, s->res, sizeof(res), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  if (res[0] != 0) return 1;
  *row = res[1];
  return SPF_OK;
}

int spf_kmpp_set_vector(spf_kmpp* s, const float* centroid) {
  return spf::guarded([&]() -> int {
  if (!s || !centroid || !s->d_vec) return fail(SPF_E_INVALID, "spf_kmpp_set_vector: needs a session of spf_kmpp_begin_sharded");
  spf_dataset* ds = s->ds;
  spf_ctx* c = ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  SPF_CUDA(cudaMemsetAsync(s->d_vec, 0, (size_t)ds->ld * sizeof(float), st));
  SPF_CUDA(cudaMemcpyAsync(s->d_vec, centroid, (size_t)ds->d * sizeof(float), cudaMemcpyHostToDevice, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  s->vec_pending = true;
  return SPF_OK;
  });
}

int spf_kmpp_rounds_sharded(spf_kmpp* s, spf_comm* comm, uint64_t row_base, const double* u01, uint32_t count,
                            uint64_t* chosen, uint32_t* done) {
  return spf::guarded([&]() -> int {
  if (!s || !u01 || !chosen || !done || !s->d_vec) return fail(SPF_E_INVALID, "spf_kmpp_rounds_sharded: bad argument");
  *done = 0;
  if (count == 0) return SPF_OK;
  if (count > (1u << 20)) return fail(SPF_E_INVALID, "spf_kmpp_rounds_sharded: at most 2^20 rounds per call");
  for (uint32_t i = 0; i < count; ++i)
    if (!(u01[i] >= 0.0 && u01[i] < 1.0)) return fail(SPF_E_INVALID, "u01 must be in [0,1)");
  if (!s->vec_pending) return fail(SPF_E_STATE, "no centroid pending: call spf_kmpp_set_vector() first");
  spf_dataset* ds = s->ds;
  spf_ctx* c = ds->ctx;
  if (comm && comm->ctx != c) return fail(SPF_E_INVALID, "communicator and dataset belong to different contexts");
  const int world = comm ? comm->world : 1, rank = comm ? comm->rank : 0;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const uint64_t n = ds->n;
  const uint32_t ld = ds->ld;
  const size_t cb = 16 + (size_t)ld * sizeof(float);     // candidate slot: row, pad, vector
  if (s->sh_world != world) {
    for (void* p : {(void*)s->sh_sums, (void*)s->sh_tinfo, (void*)s->sh_owner, (void*)s->sh_target, (void*)s->sh_cand})
      if (p) cudaFree(p);
    s->sh_sums = nullptr; s->sh_tinfo = nullptr; s->sh_owner = nullptr; s->sh_target = nullptr; s->sh_cand = nullptr;
    s->sh_world = 0;
    if (cudaMalloc((void**)&s->sh_sums, (size_t)(world + 1) * sizeof(float)) != cudaSuccess ||
        cudaMalloc((void**)&s->sh_tinfo, (size_t)(world + 1) * 2 * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&s->sh_owner, sizeof(int)) != cudaSuccess ||
        cudaMalloc((void**)&s->sh_target, sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&s->sh_cand, (size_t)(world + 1) * cb) != cudaSuccess)
      return fail(SPF_E_OOM, "sharded k-means++ buffers: out of device memory");
    s->sh_world = world;
  }
  if (count > s->batch_cap) {
    if (s->d_u01) cudaFree(s->d_u01);
    if (s->d_chosen) cudaFree(s->d_chosen);
    s->d_u01 = nullptr; s->d_chosen = nullptr; s->batch_cap = 0;
    const uint32_t cap = count < 64u ? 64u : count;
    if (cudaMalloc((void**)&s->d_u01, (size_t)cap * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&s->d_chosen, (size_t)cap * sizeof(uint64_t)) != cudaSuccess)
      return fail(SPF_E_OOM, "k-means++ batch buffers: out of device memory");
    s->batch_cap = cap;
  }
  if (!s->d_ctl && cudaMalloc((void**)&s->d_ctl, 2 * sizeof(int)) != cudaSuccess)
    return fail(SPF_E_OOM, "k-means++ batch buffers: out of device memory");
  SPF_CUDA(cudaMemcpyAsync(s->d_u01, u01, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, st));
  SPF_CUDA(cudaMemsetAsync(s->d_ctl, 0, 2 * sizeof(int), st));
  int* stop = s->d_ctl;
  uint32_t* d_done = reinterpret_cast<uint32_t*>(s->d_ctl + 1);
  float* my_sum = s->sh_sums + world;                      // slot `world`: this rank's contribution
  double* my_tinfo = s->sh_tinfo + 2 * (size_t)world;
  uint8_t* my_cand = s->sh_cand + (size_t)world * cb;
  for (uint32_t r = 0; r < count; ++r) {
    {
      KernelTimer t(c, "kmpp_update");
      const int first = (s->rounds == 0 && r == 0) ? 1 : 0;
      SPF_TRY(dispatch_metric(s->metric, [&](auto M) {
        return launch_kmpp_update<decltype(M)::value>(c, ds->x, ld, n, s->d_vec, first, s->mind, nullptr, stop);
      }));
    }
    {
      KernelTimer t(c, "kmpp_sum");
      if (c->params.kmpp_exact_sum == 1) SPF_TRY(launch_kmpp_sum(c, s->mind, n, my_sum, stop));
      else if (c->params.kmpp_exact_sum == 2) seq_sum_kernel<<<1, SEQ_THREADS, 0, st>>>(s->mind, n, my_sum, stop);
      else tree_sum_kernel<<<1, 1024, 0, st>>>(s->mind, n, my_sum, stop);
      SPF_TRY(check_launch(c, "kmpp sum kernel"));
    }
    {
      KernelTimer t(c, "kmpp_exchange");
      SPF_TRY(comm_allgather(c, comm, my_sum, s->sh_sums, sizeof(float)));
    }
    {
      KernelTimer t(c, "kmpp_pick");
      kmpp_combine_sums_kernel<<<1, 1, 0, st>>>(s->sh_sums, world, s->d_sum, stop);
      SPF_CUDA(cudaMemsetAsync(s->bad, 0, sizeof(int), st));
      kmpp_block_sums_kernel<<<(unsigned)s->nblocks, 256, 0, st>>>(s->mind, n, s->d_sum, s->block_sums, s->bad, stop);
      kmpp_total_kernel<<<1, PICK_THREADS, 0, st>>>(s->block_sums, s->nblocks, s->bad, my_tinfo, stop);
      SPF_TRY(check_launch(c, "kmpp total kernels", 3));
    }
    {
      KernelTimer t(c, "kmpp_exchange");
      SPF_TRY(comm_allgather(c, comm, my_tinfo, s->sh_tinfo, 2 * sizeof(double)));
    }
    {
      KernelTimer t(c, "kmpp_pick");
      kmpp_owner_kernel<<<1, 1, 0, st>>>(s->sh_tinfo, world, s->d_u01, r, s->sh_owner, s->sh_target, stop);
      KmppBatch bt;
      bt.stop = stop; bt.owner = s->sh_owner; bt.d_target = s->sh_target; bt.my_rank = rank; bt.round = r;
      kmpp_pick_kernel<<<1, PICK_THREADS, 0, st>>>(s->mind, n, s->d_sum, s->block_sums, s->nblocks, s->bad, 0.0, 0.0, s->res, s->d_total, bt);
      kmpp_pack_kernel<<<1, 128, 0, st>>>(ds->x, ld, s->res, s->sh_owner, rank, row_base, my_cand, stop);
      SPF_TRY(check_launch(c, "kmpp pick kernels", 3));
    }
    {
      KernelTimer t(c, "kmpp_exchange");
      SPF_TRY(comm_allgather(c, comm, my_cand, s->sh_cand, cb));
    }
    kmpp_take_kernel<<<1, 128, 0, st>>>(s->sh_cand, cb, s->sh_owner, s->d_vec, ld, s->d_chosen, r, d_done, stop);
    SPF_TRY(check_launch(c, "kmpp_take_kernel"));
  }
  int ctl[2] = {0, 0};
  SPF_CUDA(cudaMemcpyAsync(ctl, s->d_ctl, sizeof(ctl), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaMemcpyAsync(chosen, s->d_chosen, (size_t)count * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaMemcpyAsync(&s->last_sum, s->d_sum, sizeof(float), cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  const uint32_t ok = (uint32_t)ctl[1];
  const bool failed = ctl[0] != 0;
  *done = ok;
  s->rounds += ok + (failed ? 1u : 0u);                   // the failing round folded its centroid too
  s->vec_pending = !failed;
  return failed ? 1 : SPF_OK;
  });
}

int spf_kmpp_push(spf_kmpp* s, uint64_t row) {
  if (!s) return fail(SPF_E_INVALID, "session is NULL");
  if (s->pending) return fail(SPF_E_STATE, "a centroid is already pending");
  if (row >= s->ds->n) return fail(SPF_E_INVALID, "row >= n");
  s->newest = row;
  s->pending = true;
  return SPF_OK;
}

int spf_kmpp_last_sums(const spf_kmpp* s, float* sum, double* total) {
  if (!s) return fail(SPF_E_INVALID, "session is NULL");
  if (sum) *sum = s->last_sum;
  if (total) *total = s->last_total;
  return SPF_OK;
}

void spf_kmpp_free(spf_kmpp* s) {
  if (!s) return;
  cudaSetDevice(s->ds->ctx->device);
  cudaStreamSynchronize(s->ds->ctx->stream);
  if (s->mind) cudaFree(s->mind);
  if (s->block_sums) cudaFree(s->block_sums);
  if (s->bad) cudaFree(s->bad);
  if (s->res) cudaFree(s->res);
  if (s->d_sum) cudaFree(s->d_sum);
  if (s->d_total) cudaFree(s->d_total);
  if (s->d_vec) cudaFree(s->d_vec);
  if (s->d_u01) cudaFree(s->d_u01);
  if (s->d_chosen) cudaFree(s->d_chosen);
  if (s->d_ctl) cudaFree(s->d_ctl);
  for (void* p : {(void*)s->sh_sums, (void*)s->sh_tinfo, (void*)s->sh_owner, (void*)s->sh_target, (void*)s->sh_cand})
    if (p) cudaFree(p);
  delete s;
}

}  // extern "C"
