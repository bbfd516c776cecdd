// api.cu — context, dataset and error plumbing of the C ABI (include/spfresh_b200.h).
#include <cuda.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "kernels.cuh"

namespace spf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace spf

using namespace spf;

extern "C" {

int spf_abi_version(void) { return SPF_ABI_VERSION; }

const char* spf_last_error(void) { return spf::g_err; }

int spf_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int spf_ctx_create(int device, spf_ctx** out) {
  if (!out) return fail(SPF_E_INVALID, "spf_ctx_create: out is NULL");
  *out = nullptr;
  int n = spf_device_count();
  if (n <= 0) return fail(SPF_E_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
  if (device < 0 || device >= n) return fail(SPF_E_INVALID, "device %d out of range [0,%d)", device, n);
  cudaDeviceProp prop;
  SPF_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(SPF_E_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                device, prop.major, prop.minor);
  SPF_CUDA(cudaSetDevice(device));
  spf_ctx* c = new (std::nothrow) spf_ctx();
  if (!c) return fail(SPF_E_OOM, "out of host memory");
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev[0]);
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev[1]);
  if (e != cudaSuccess) {
    delete c;
    return fail(SPF_E_CUDA, "context setup failed: %s", cudaGetErrorString(e));
  }
  // keep freed temporaries cached in the stream-ordered pool
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    fn = nullptr;
  cudaGetLastError();
  c->tma_encode = fn;
  const char* env = getenv("SPF_FORCE_EXACT");
  if (env && atoi(env)) c->params.force_exact = 1;
  env = getenv("SPF_TC_PIPE");                     // experiments: epilogue variant of assign_tc_kernel
  if (env) c->params.tc_pipe = atoi(env);
  *out = c;
  return SPF_OK;
}

void spf_ctx_destroy(spf_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->ev[0]) cudaEventDestroy(c->ev[0]);
  if (c->ev[1]) cudaEventDestroy(c->ev[1]);
  for (auto& kv : c->scratch)
    if (kv.second.p) cudaFree(kv.second.p);
  if (c->cc_cache.cc) cudaFree(c->cc_cache.cc);
  if (c->cc_cache.C) cudaFree(c->cc_cache.C);
  if (c->cc_cache.same) cudaFree(c->cc_cache.same);
  if (c->stage.base) {
    cudaFreeHost(c->stage.base);
    for (cudaEvent_t e : c->stage.ev) if (e) cudaEventDestroy(e);
  }
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
  if (c->aux_ev[0]) cudaEventDestroy(c->aux_ev[0]);
  if (c->aux_ev[1]) cudaEventDestroy(c->aux_ev[1]);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int spf_ctx_device(const spf_ctx* c) { return c ? c->device : -1; }
void* spf_ctx_stream(spf_ctx* c) { return c ? (void*)c->stream : nullptr; }

int spf_ctx_synchronize(spf_ctx* c) {
  if (!c) return fail(SPF_E_INVALID, "ctx is NULL");
  SPF_CUDA(cudaStreamSynchronize(c->stream));
  return SPF_OK;
}

int spf_ctx_trim(spf_ctx* c) {
  if (!c) return fail(SPF_E_INVALID, "ctx is NULL");
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  SPF_CUDA(cudaStreamSynchronize(c->stream));
  for (auto& kv : c->scratch) {
    if (kv.second.busy) continue;
    if (kv.second.p) cudaFreeAsync(kv.second.p, c->stream);
    kv.second.p = nullptr;
    kv.second.bytes = 0;
  }
  SPF_CUDA(cudaStreamSynchronize(c->stream));
  cudaMemPool_t pool;
  SPF_CUDA(cudaDeviceGetDefaultMemPool(&pool, c->device));
  SPF_CUDA(cudaMemPoolTrimTo(pool, 0));
  return SPF_OK;
}

int spf_ctx_set_profiling(spf_ctx* c, int enabled) {
  if (!c) return fail(SPF_E_INVALID, "ctx is NULL");
  std::lock_guard<std::mutex> lk(c->mu);
  c->profiling = enabled != 0;
  return SPF_OK;
}

float spf_ctx_kernel_ms(spf_ctx* c, const char* name) {
  if (!c || !name) return -1.0f;
  try {
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->kernel_ms.find(name);
    return it == c->kernel_ms.end() ? -1.0f : it->second;
  } catch (...) {
    return -1.0f;
  }
}

uint64_t spf_ctx_launch_count(const spf_ctx* c) { return c ? c->launches : 0; }
uint32_t spf_ctx_last_overflow_rows(const spf_ctx* c) { return c ? c->last_overflow_rows : 0; }

// Internal tuning knobs (tests use them to force the rare paths; not part of the drop-in ABI).
int spf_ctx_set_param(spf_ctx* c, const char* name, int value) {
  return spf::guarded([&]() -> int {
  if (!c || !name) return fail(SPF_E_INVALID, "ctx/name is NULL");
  std::lock_guard<std::mutex> lk(c->mu);
  std::string s(name);
  if (s == "cand_cap") {
    if (value < 2 || value > 1024 || (value & 1)) return fail(SPF_E_INVALID, "cand_cap must be even and in [2,1024]");
    c->params.cand_cap = value;
  } else if (s == "short_cap") {
    if (value < 2 || value > 64 || (value & (value - 1))) return fail(SPF_E_INVALID, "short_cap must be a power of two in [2,64]");
    c->params.short_cap = value;
  } else if (s == "force_exact") c->params.force_exact = value;
  else if (s == "chunk_rows") c->params.chunk_rows = value;
  else if (s == "work_cap") c->params.work_cap = value;
  else if (s == "scan_list_major") c->params.scan_list_major = value;
  else if (s == "scan_tc") c->params.scan_tc = value;
  else if (s == "scan_tc_bucket") {
    if (value < 0 || value > 65536) return fail(SPF_E_INVALID, "scan_tc_bucket must be in [0,65536]");
    c->params.scan_tc_bucket = value;
  } else if (s == "scan_tc_tau_probes") c->params.scan_tc_tau_probes = value;
  else if (s == "scan_tc_cmax_mb") c->params.scan_tc_cmax_mb = value;
  else if (s == "scan_tc_split") c->params.scan_tc_split = value;
  else if (s == "search_upload_pieces") {
    if (value < 1 || value > 8) return fail(SPF_E_INVALID, "search_upload_pieces must be in [1,8]");
    c->params.search_upload_pieces = value;
  }
  else if (s == "tc_min_k") c->params.tc_min_k = value;
  else if (s == "tc_min_m") c->params.tc_min_m = value;
  else if (s == "tc_pipe") c->params.tc_pipe = value;
  else if (s == "finalize_lanes") c->params.finalize_lanes = value;
  else if (s == "medoid_direct") c->params.medoid_direct = value;
  else if (s == "sum_fast") c->params.sum_fast = value;
  else if (s == "kmpp_exact_sum") c->params.kmpp_exact_sum = value;
  else if (s == "no_host_staging") c->params.no_host_staging = value;
  else if (s == "cc_matrix_max_k") c->params.cc_matrix_max_k = value;
  else if (s == "exact_tma") c->params.exact_tma = value;
  else if (s == "exact_tma_min_pairs") c->params.exact_tma_min_pairs = value;
  else if (s == "exact_packed") c->params.exact_packed = value;
  else if (s == "scratch_cache") c->params.scratch_cache = value;
  else if (s == "exact_seed") c->params.exact_seed = value;
  else if (s == "exact_one_cta") c->params.exact_one_cta = value;
  else if (s == "exact_three_cta") c->params.exact_three_cta = value;
  else if (s == "sum_hub") c->params.sum_hub = value;
  else if (s == "sum_slices") c->params.sum_slices = value;
  else if (s == "cc_cache") c->params.cc_cache = value;
  else return fail(SPF_E_INVALID, "unknown parameter '%s'", name);
  return SPF_OK;
  });
}

// ---------------------------------------------------------------------------------------------
// dataset
// ---------------------------------------------------------------------------------------------
}  // extern "C"

int spf::dataset_alloc(spf_ctx* c, uint64_t n, uint32_t d, spf_dataset** out) {
  if (!c || !out) return fail(SPF_E_INVALID, "ctx/out is NULL");
  if (n == 0 || d == 0) return fail(SPF_E_INVALID, "dataset must have n > 0 and d > 0");
  if (n >= (1ull << 32)) return fail(SPF_E_INVALID, "n must be < 2^32 rows per device shard");
  SPF_CUDA(cudaSetDevice(c->device));
  spf_dataset* ds = new (std::nothrow) spf_dataset();
  if (!ds) return fail(SPF_E_OOM, "out of host memory");
  ds->ctx = c;
  ds->n = n;
  ds->d = d;
  ds->ld = round_up(d, 4);
  size_t bytes = (size_t)n * ds->ld * sizeof(float);
  // stream-ordered pool: re-uploading a dataset of the same size reuses the cached block
  cudaError_t e = cudaMallocAsync((void**)&ds->x, bytes, c->stream);
  if (e != cudaSuccess) {
    delete ds;
    return fail(SPF_E_OOM, "allocation of %zu bytes for the dataset failed: %s", bytes, cudaGetErrorString(e));
  }
  *out = ds;
  return SPF_OK;
}

extern "C" {

int spf_dataset_upload(spf_ctx* c, const float* rows, uint64_t n, uint32_t d, uint64_t row_stride,
                       spf_dataset** out) {
  return spf::guarded([&]() -> int {
  if (!rows) return fail(SPF_E_INVALID, "rows is NULL");
  if (row_stride < d) return fail(SPF_E_INVALID, "row_stride (%llu) < d (%u)", (unsigned long long)row_stride, d);
  spf_dataset* ds = nullptr;
  SPF_TRY(spf::dataset_alloc(c, n, d, &ds));
  std::lock_guard<std::mutex> lk(c->mu);
  cudaError_t e = cudaSuccess;
  if ((size_t)n * d * sizeof(float) >= (8u << 20) && spf::host_pointer_is_pageable(rows) && !c->params.no_host_staging) {
    // ordinary heap memory: threaded staging through the pinned ring instead of the driver's one-thread path
    const int rc = spf::staged_upload(c, rows, n, d, row_stride, ds->x, ds->ld, n, c->stream, [](uint64_t) { return SPF_OK; });
    if (rc == SPF_OK) e = cudaStreamSynchronize(c->stream);
    if (rc != SPF_OK || e != cudaSuccess) {
      spf_dataset_free(ds);
      return rc != SPF_OK ? rc : fail(SPF_E_CUDA, "dataset upload failed: %s", cudaGetErrorString(e));
    }
    *out = ds;
    return SPF_OK;
  }
  if (ds->ld != d) e = cudaMemsetAsync(ds->x, 0, (size_t)n * ds->ld * sizeof(float), c->stream);
  if (e == cudaSuccess)
    e = cudaMemcpy2DAsync(ds->x, (size_t)ds->ld * sizeof(float), rows, (size_t)row_stride * sizeof(float),
                          (size_t)d * sizeof(float), (size_t)n, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    spf_dataset_free(ds);
    return fail(SPF_E_CUDA, "dataset upload failed: %s", cudaGetErrorString(e));
  }
  *out = ds;
  return SPF_OK;
  });
}

int spf_dataset_from_device(spf_ctx* c, const void* dev_rows, uint64_t n, uint32_t d, spf_dataset** out) {
  if (!dev_rows) return fail(SPF_E_INVALID, "dev_rows is NULL");
  spf_dataset* ds = nullptr;
  SPF_TRY(spf::dataset_alloc(c, n, d, &ds));
  std::lock_guard<std::mutex> lk(c->mu);
  cudaError_t e = cudaSuccess;
  if (ds->ld != d) e = cudaMemsetAsync(ds->x, 0, (size_t)n * ds->ld * sizeof(float), c->stream);
  if (e == cudaSuccess)
    e = cudaMemcpy2DAsync(ds->x, (size_t)ds->ld * sizeof(float), dev_rows, (size_t)d * sizeof(float),
                          (size_t)d * sizeof(float), (size_t)n, cudaMemcpyDeviceToDevice, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    spf_dataset_free(ds);
    return fail(SPF_E_CUDA, "dataset device copy failed: %s", cudaGetErrorString(e));
  }
  *out = ds;
  return SPF_OK;
}

void spf_dataset_free(spf_dataset* ds) {
  if (!ds) return;
  cudaSetDevice(ds->ctx->device);
  cudaStreamSynchronize(ds->ctx->stream);
  cudaStream_t st = ds->ctx->stream;
  if (ds->x) cudaFreeAsync(ds->x, st);
  if (ds->xtf) cudaFreeAsync(ds->xtf, st);
  if (ds->xnorm) cudaFreeAsync(ds->xnorm, st);
  if (ds->xres) cudaFreeAsync(ds->xres, st);
  delete ds;
}

int spf_dataset_fetch_rows(spf_dataset* ds, const uint64_t* rows, uint64_t m, float* out) {
  if (!ds || (!rows && m) || (!out && m)) return fail(SPF_E_INVALID, "spf_dataset_fetch_rows: NULL argument");
  if (m == 0) return SPF_OK;
  for (uint64_t i = 0; i < m; ++i)
    if (rows[i] >= ds->n) return fail(SPF_E_INVALID, "row %llu >= n", (unsigned long long)rows[i]);
  spf_ctx* c = ds->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  spf::DevBuf<uint64_t> d_idx;
  spf::DevBuf<float> g;
  SPF_TRY(d_idx.alloc(st, m));
  SPF_TRY(g.alloc(st, (size_t)m * ds->ld));
  SPF_CUDA(cudaMemcpyAsync(d_idx.p, rows, m * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  SPF_TRY(spf::launch_gather_rows(c, ds->x, ds->ld, d_idx.p, m, g.p));
  SPF_CUDA(cudaMemcpy2DAsync(out, (size_t)ds->d * sizeof(float), g.p, (size_t)ds->ld * sizeof(float),
                             (size_t)ds->d * sizeof(float), m, cudaMemcpyDeviceToHost, st));
  SPF_CUDA(cudaStreamSynchronize(st));
  return SPF_OK;
}

uint64_t spf_dataset_rows(const spf_dataset* ds) { return ds ? ds->n : 0; }
uint32_t spf_dataset_dim(const spf_dataset* ds) { return ds ? ds->d : 0; }

}  // extern "C"
