// common.cuh — shared host/device plumbing for libspfresh_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <exception>
#include <map>
#include <new>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/spfresh_b200.h"

namespace spf {

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define SPF_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      return ::spf::fail(_e == cudaErrorMemoryAllocation ? SPF_E_OOM : SPF_E_CUDA,          \
                         "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                         __LINE__);                                                         \
    }                                                                                       \
  } while (0)

// No exception may cross the C ABI (include/spfresh_b200.h): entry points that touch std containers
// run their body through this guard.
template <typename F>
inline int guarded(F&& f) noexcept {
  try {
    return f();
  } catch (const std::bad_alloc&) {
    return fail(SPF_E_OOM, "out of host memory");
  } catch (const std::exception& e) {
    return fail(SPF_E_INVALID, "unexpected exception: %s", e.what());
  } catch (...) {
    return fail(SPF_E_INVALID, "unexpected exception");
  }
}

#define SPF_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc < 0) return _rc;     \
  } while (0)

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct Params {
  int cand_cap = 128;        // candidate group records per point kept by the assign kernels
  int short_cap = 64;        // short-list entries per point kept by resolve (power of two, <= 64)
  int force_exact = 0;       // 1: never use the tcgen05 candidate GEMM
  int tc_min_k = 64;         // use the tensor path only when k >= this
  int tc_min_m = 1024;       // ... and m >= this
  int finalize_lanes = 16;   // finalize_kernel: lanes per point (8, 16 or 32)
  int medoid_direct = 1;     // medoid_kernel: 1 stage only the member rows and read the cluster mean in place, 0 stage both rows of every pair
  int tc_pipe = 0;           // assign_tc_kernel epilogue: 0 load the whole slice first, 1 half-slice pipeline across tiles
  int no_host_staging = 0;   // 1: pageable host buffers go to cudaMemcpyAsync directly (driver-staged), for comparison
  int kmpp_exact_sum = 1;    // 1: sequential f32 sum (bit-parity with the reference)
  int cc_matrix_max_k = 16384;  // precompute the k x k centroid-centroid matrix up to this k
  int scan_threads = 256;
  int scan_list_major = 1;   // 0: query-major scan only, 1: automatic, 2: always list-major (k <= 32)
  int scan_tc = 1;           // tensor-core candidate scan: 0 never, 1 automatic, 2 whenever supported (k <= 16, d <= 128)
  int scan_tc_bucket = 0;    // candidate entries per query kept by the tensor scan (overflow: exact fallback); 0: 256, or 1024 for d > 256
  int scan_tc_cmax_mb = 40960; // keep the bound pass's chunk maxima (one GEMM pass) while they fit in this many MB; 0: always two passes
  int scan_tc_tau_probes = 0;  // probes per query that take part in the bound pass (0: all)
  int search_upload_pieces = 2; // spf_search_batch, >= 32768 queries: upload in this many pieces, probe of a piece under the next upload (<= 8; 1: off)
  int scan_tc_split = 1;       // bound pass as two concurrent launches (multi-unit lists | single-unit lists): 0 off, 1 automatic, > 1 SMs of the first
  int exact_tma = 6;         // CUDA-core direct-form kernel: bit m set = metric m uses the TMA-staged 128 x 128 kernel (default: Manhattan, Chebyshev)
  int sum_hub = 0;           // compute_mean: clusters of at least this many members use the deep, column-sliced launch (0: 8192)
  int sum_slices = 0;        // ... column slices per hub cluster (0: one; measured slower when > 1)
  int sum_fast = 1;          // compute_mean producers: suspended barrier waits + three-instruction copy loop for 512-byte rows (0: polling waits, generic loop)
  int exact_packed = 6;      // ... bit per metric: differences of two dimensions from one packed FADD2
  int exact_three_cta = 0;   // ... bit per metric: the packed kernel in its 64-point shape, three CTAs of 128 threads per SM
  int exact_one_cta = 2;     // ... bit per metric: the packed kernel compiled for one CTA per SM (more registers)
  int exact_seed = 1;        // Manhattan / Chebyshev: seed the running minimum from a tensor-core L2 pre-pass (0 off, 1 long rows, 2 always)
  int scratch_cache = 1;     // keep the large temporaries of assign between calls (see spf_ctx::scratch)
  int exact_tma_min_pairs = 1 << 16;   // ... for problems of at least this many (point, centroid) pairs
  int cc_cache = 1;          // keep the k x k centroid matrix while the centroid vectors do not change
  int chunk_rows = 0;        // points per assign chunk (0: automatic)
  int work_cap = 0;          // exact-evaluation work-list entries per chunk (0: automatic, 8 per point)
};

}  // namespace spf

struct spf_ctx {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // host-streamed assign: uploads overlap the main stream
  cudaStream_t aux_stream = nullptr;    // independent kernels that may overlap the main stream (hub clusters of the mean)
  cudaEvent_t aux_ev[2] = {nullptr, nullptr};
  std::mutex mu;
  bool profiling = false;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  std::map<std::string, float> kernel_ms;
  uint64_t launches = 0;
  uint32_t last_overflow_rows = 0;   // rows the last assign resolved through the dense fallback
  spf::Params params;
  void* tma_encode = nullptr;  // cuTensorMapEncodeTiled, resolved at context creation
  // k x k centroid-centroid matrix of the last assign, kept while the centroid vectors stay the
  // same bit for bit (checked on the device each call: assign_api.cu)
  struct CcCache {
    float* cc = nullptr;   // k x k
    float* C = nullptr;    // k x ld centroid vectors the matrix was computed from
    int* same = nullptr;   // device flag: 1 = this call's centroids equal the cached ones
    uint32_t k = 0, ld = 0;
    int metric = -1;
    bool valid = false;
  } cc_cache;
  // Large per-call temporaries of the assign path (candidate records, short lists, member slots,
  // sort buffers) are kept between calls, one grow-only slot per name: the multi-GB blocks of a
  // 100 M-row k-means iteration otherwise make the stream-ordered pool remap memory every
  // iteration (single iterations of 300 ms were measured at 1.8 s).  Freed by spf_ctx_trim / destroy.
  struct ScratchSlot { void* p = nullptr; size_t bytes = 0; bool busy = false; };
  std::map<std::string, ScratchSlot> scratch;
  // Pinned staging ring for PAGEABLE host buffers (the reference hands ndarray views, i.e. ordinary
  // heap memory): worker threads copy blocks into the ring, the copy engine drains it
  // (assign_api.cu: staged_upload / staged_download).  Allocated on first use, freed with the context.
  struct HostStage {
    static constexpr int SLOTS = 12;
    static constexpr size_t SLOT_BYTES = 4u << 20;
    uint8_t* base = nullptr;
    cudaEvent_t ev[SLOTS] = {};
  } stage;
};

struct spf_dataset {
  spf_ctx* ctx = nullptr;
  float* x = nullptr;     // n x ld, rows zero-padded to ld (multiple of 4 floats)
  // tensor path only, made lazily once per dataset (row_prep_kernel):
  float* xtf = nullptr;   // n x ld rows rounded to TF32 (the GEMM's A operand)
  float* xnorm = nullptr; // squared norms |x|^2
  float* xres = nullptr;  // rounding residual norms |x - xtf|
  bool prepped = false;   // xtf / xnorm / xres hold valid data for all rows
  uint64_t n = 0;
  uint32_t d = 0, ld = 0;
};

struct spf_assign_result {
  spf_ctx* ctx = nullptr;
  uint64_t m = 0;
  uint32_t k = 0;
  uint64_t total = 0;
  bool has_csr = false;
  uint32_t* best = nullptr;      // device, m
  float* dmin = nullptr;         // device, m
  uint64_t* offsets = nullptr;   // device, k+1
  uint32_t* members = nullptr;   // device, total: positions into the point list
  uint64_t* point_idx = nullptr; // device, m dataset rows, or NULL for identity
};

namespace spf {

// RAII device temporaries on the context stream (stream-ordered pool).
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t s = nullptr;
  spf_ctx::ScratchSlot* slot = nullptr;   // set when the memory belongs to the context's scratch cache
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  int alloc(cudaStream_t stream, size_t count) {
    release();
    s = stream;
    n = count;
    size_t bytes = (count ? count : 1) * sizeof(T);
    cudaError_t e = cudaMallocAsync((void**)&p, bytes, stream);
    if (e != cudaSuccess) {
      p = nullptr;
      return fail(SPF_E_OOM, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    return SPF_OK;
  }
  // The same from the context's named scratch slot (all work of a context is ordered on its stream,
  // so reusing the block needs no synchronisation).  Small requests and a slot that is in use go
  // through the pool.  A cached buffer must not be take()n.
  int alloc_cached(spf_ctx* c, const char* tag, size_t count) {
    release();
    const size_t bytes = (count ? count : 1) * sizeof(T);
    if (bytes < (1u << 20) || !c->params.scratch_cache) return alloc(c->stream, count);
    spf_ctx::ScratchSlot& sl = c->scratch[tag];
    if (sl.busy) return alloc(c->stream, count);
    if (sl.bytes < bytes) {
      if (sl.p) cudaFreeAsync(sl.p, c->stream);
      sl.p = nullptr;
      sl.bytes = 0;
      const size_t want = bytes + bytes / 8;                 // headroom: sizes drift between iterations
      cudaError_t e = cudaMallocAsync(&sl.p, want, c->stream);
      if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMallocAsync(&sl.p, bytes, c->stream);
        if (e != cudaSuccess) {
          sl.p = nullptr;
          return fail(SPF_E_OOM, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        }
        sl.bytes = bytes;
      } else {
        sl.bytes = want;
      }
    }
    sl.busy = true;
    slot = &sl;
    s = c->stream;
    n = count;
    p = static_cast<T*>(sl.p);
    return SPF_OK;
  }
  void release() {
    if (slot) slot->busy = false;
    else if (p) cudaFreeAsync(p, s);
    slot = nullptr;
    p = nullptr;
    n = 0;
  }
  T* take() { T* q = p; p = nullptr; return q; }
};

// Brackets a named kernel (or group of kernels) with its own pair of events when profiling is
// on, so timers may nest (adds a sync per timer: only for bench / profiling runs).
struct KernelTimer {
  spf_ctx* c;
  const char* name;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  KernelTimer(spf_ctx* ctx, const char* nm) : c(ctx), name(nm) {
    if (c->profiling) {
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0, c->stream);
    }
  }
  ~KernelTimer() {
    if (c->profiling && e0 && e1) {
      cudaEventRecord(e1, c->stream);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      c->kernel_ms[name] += ms;   // accumulates over the chunks of one call (cleared per call)
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  }
};

inline int check_launch(spf_ctx* c, const char* what, int nlaunches = 1) {
  c->launches += (uint64_t)nlaunches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SPF_E_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  return SPF_OK;
}

inline uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }
inline uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

}  // namespace spf

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace spf {

// Exact, un-fused element update of the three metrics (src/distances/distance.rs:16-43 →
// ndarray-stats sq_l2_dist / l1_dist / linf_dist).  The intrinsics stop nvcc from contracting
// mul+add into FMA, which would change the rounding.
template <int METRIC>
__device__ __forceinline__ float dist_step(float acc, float a, float b) {
  float df = __fsub_rn(a, b);
  if (METRIC == SPF_METRIC_EUCLIDEAN) return __fadd_rn(acc, __fmul_rn(df, df));
  if (METRIC == SPF_METRIC_MANHATTAN) return __fadd_rn(acc, fabsf(df));
  return fmaxf(acc, fabsf(df));   // linf: `if |df| > max {max = |df|}`; NaN leaves max unchanged
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace spf
#endif
