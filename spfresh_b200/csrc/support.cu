// support.cu — small streaming kernels shared by the assign / update / k-means++ / query paths.
#include "kernels.cuh"
#include "pairdist.cuh"

namespace spf {

namespace {

__global__ void gather_rows_kernel(const float* __restrict__ src, uint32_t ld4, const uint64_t* __restrict__ idx,
                                   uint64_t m, float* __restrict__ dst) {
  // one float4 per thread; a row is ld4 float4s
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t total = m * ld4;
  if (t >= total) return;
  const uint64_t r = t / ld4;
  const uint32_t c = (uint32_t)(t - r * ld4);
  const float4* s = reinterpret_cast<const float4*>(src) + (size_t)idx[r] * ld4 + c;
  reinterpret_cast<float4*>(dst)[t] = __ldg(s);
}

__global__ void row_sqnorm_kernel(const float* __restrict__ rows, uint32_t ld4, uint64_t m, float* __restrict__ out) {
  // one warp per row: lanes stride over float4s, shuffle reduce (the norm only feeds the
  // tensor-path error bound and approximate distances; its summation order is free).
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= m) return;
  const float4* r = reinterpret_cast<const float4*>(rows) + (size_t)w * ld4;
  float acc = 0.f;
  for (uint32_t c = lane; c < ld4; c += 32) {
    const float4 v = __ldg(r + c);
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc);
    acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[w] = acc;
}

// Rounded GEMM operands + norms (see launch_row_prep in kernels.cuh).  One warp per row.
__global__ void row_prep_kernel(const float* __restrict__ src, uint32_t ld4, const uint64_t* __restrict__ idx,
                                uint64_t m, float* __restrict__ tf, float* __restrict__ norm,
                                float* __restrict__ res) {
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= m) return;
  const float4* r = reinterpret_cast<const float4*>(src) + (size_t)(idx ? idx[w] : w) * ld4;
  float4* o = reinterpret_cast<float4*>(tf) + (size_t)w * ld4;
  float acc = 0.f, racc = 0.f;
  auto rnd = [](float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
  };
  for (uint32_t c = lane; c < ld4; c += 32) {
    const float4 v = __ldg(r + c);
    const float4 t = make_float4(rnd(v.x), rnd(v.y), rnd(v.z), rnd(v.w));
    o[c] = t;
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc);
    acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    const float ex = v.x - t.x, ey = v.y - t.y, ez = v.z - t.z, ew = v.w - t.w;   // exact (Sterbenz)
    racc = fmaf(ex, ex, racc); racc = fmaf(ey, ey, racc);
    racc = fmaf(ez, ez, racc); racc = fmaf(ew, ew, racc);
  }
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, of);
    racc += __shfl_xor_sync(0xffffffffu, racc, of);
  }
  if (lane == 0) {
    norm[w] = acc;
    res[w] = sqrtf(racc);
  }
}

__global__ void max2_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, uint64_t n,
                                float* out2) {
  // single block; values are >= 0
  __shared__ float sa[32], sb[32];
  float va = 0.f, vb = 0.f;
  for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) {
    va = fmaxf(va, a[i]);
    vb = fmaxf(vb, b[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    va = fmaxf(va, __shfl_xor_sync(0xffffffffu, va, o));
    vb = fmaxf(vb, __shfl_xor_sync(0xffffffffu, vb, o));
  }
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = va; sb[threadIdx.x >> 5] = vb; }
  __syncthreads();
  if (threadIdx.x < 32) {
    va = threadIdx.x < (blockDim.x >> 5) ? sa[threadIdx.x] : 0.f;
    vb = threadIdx.x < (blockDim.x >> 5) ? sb[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      va = fmaxf(va, __shfl_xor_sync(0xffffffffu, va, o));
      vb = fmaxf(vb, __shfl_xor_sync(0xffffffffu, vb, o));
    }
    if (threadIdx.x == 0) { out2[0] = va; out2[1] = vb; }
  }
}

__global__ void centroid_ext_kernel(const float* __restrict__ cnorm, uint32_t k, uint32_t kpad, float* __restrict__ cext) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= kpad) return;
  auto rnd = [](float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
  };
  float h = -__int_as_float(0x7f800000), m = 0.f, l = 0.f;
  if (j < k) {
    const float v = -0.5f * cnorm[j];          // exact (power of two)
    h = rnd(v);
    const float r1 = v - h;                    // exact: |r1| <= 2^-11 |v|
    m = rnd(r1);
    l = rnd(r1 - m);                           // exact difference, then rounded once more
  }
  float4* o = reinterpret_cast<float4*>(cext + (size_t)j * 8);
  o[0] = make_float4(h, m, 0.f, 0.f);
  o[1] = make_float4(l, 0.f, 0.f, 0.f);
}

__global__ void fill_f32_kernel(float* p, uint64_t n, float v) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) p[t] = v;
}
__global__ void fill_u64_kernel(uint64_t* p, uint64_t n, uint64_t v) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) p[t] = v;
}

__global__ void max_f32_kernel(const float* __restrict__ p, uint64_t n, float* out) {
  // single block; values are >= 0 (squared norms)
  __shared__ float sm[32];
  float v = 0.f;
  for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) v = fmaxf(v, p[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) out[0] = v;
  }
}

__global__ void check_rows_kernel(const uint64_t* __restrict__ idx, uint64_t m, uint64_t n, int* flag) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < m && idx[t] >= n) *flag = 1;
}

template <int METRIC>
__global__ void __launch_bounds__(PD_THREADS)
pair_dist_kernel(const float* __restrict__ A, uint32_t ldA, const uint64_t* __restrict__ idxA,
                 const float* __restrict__ B, uint32_t ldB, const uint32_t* __restrict__ idxB32,
                 uint64_t fixedB, uint32_t ld, uint64_t count, float* __restrict__ out) {
  __shared__ PairDistSmem sm[PD_THREADS / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t nwarps = (uint64_t)gridDim.x * (PD_THREADS / 32);
  for (uint64_t base = ((uint64_t)blockIdx.x * (PD_THREADS / 32) + warp) * 32; base < count;
       base += nwarps * 32) {
    const uint64_t i = base + lane;
    const bool valid = i < count;
    const float* pa = nullptr;
    const float* pb = nullptr;
    if (valid) {
      pa = A + (size_t)(idxA ? idxA[i] : i) * ldA;
      pb = B + (size_t)(idxB32 ? (uint64_t)idxB32[i] : (fixedB == UINT64_MAX ? i : fixedB)) * ldB;
    }
    const float dv = warp_pair_dist<METRIC>(pa, pb, ld, sm[warp]);
    if (valid) out[i] = dv;
  }
}

}  // namespace

int launch_gather_rows(spf_ctx* c, const float* src, uint32_t ld, const uint64_t* d_idx, uint64_t m,
                       float* dst) {
  if (m == 0) return SPF_OK;
  const uint32_t ld4 = ld / 4;
  const uint64_t total = m * ld4;
  gather_rows_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, c->stream>>>(src, ld4, d_idx, m, dst);
  return check_launch(c, "gather_rows_kernel");
}

int launch_row_sqnorm(spf_ctx* c, const float* rows, uint32_t ld, uint64_t m, float* out) {
  if (m == 0) return SPF_OK;
  row_sqnorm_kernel<<<(unsigned)ceil_div(m * 32, 256), 256, 0, c->stream>>>(rows, ld / 4, m, out);
  return check_launch(c, "row_sqnorm_kernel");
}

int launch_row_prep(spf_ctx* c, const float* src, uint32_t ld, const uint64_t* d_idx, uint64_t m,
                    float* tf, float* norm, float* res) {
  if (m == 0) return SPF_OK;
  row_prep_kernel<<<(unsigned)ceil_div(m * 32, 256), 256, 0, c->stream>>>(src, ld / 4, d_idx, m, tf, norm, res);
  return check_launch(c, "row_prep_kernel");
}

int launch_centroid_ext(spf_ctx* c, const float* cnorm, uint32_t k, uint32_t kpad, float* cext) {
  centroid_ext_kernel<<<(kpad + 255) / 256, 256, 0, c->stream>>>(cnorm, k, kpad, cext);
  return check_launch(c, "centroid_ext_kernel");
}

int launch_max2_f32(spf_ctx* c, const float* a, const float* b, uint64_t n, float* out2) {
  max2_f32_kernel<<<1, 1024, 0, c->stream>>>(a, b, n, out2);
  return check_launch(c, "max2_f32_kernel");
}

int launch_fill_f32(spf_ctx* c, float* p, uint64_t n, float v) {
  if (n == 0) return SPF_OK;
  fill_f32_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, c->stream>>>(p, n, v);
  return check_launch(c, "fill_f32_kernel");
}

__global__ void add_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __fadd_rn(dst[i], src[i]);
}

__global__ void scale_u64_f32_kernel(const uint64_t* __restrict__ src, float scale, uint64_t n, float* __restrict__ dst) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __fmul_rn(scale, __ull2float_rn(src[i]));
}

int launch_add_f32(spf_ctx* c, float* dst, const float* src, uint64_t n) {
  if (n == 0) return SPF_OK;
  add_f32_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, c->stream>>>(dst, src, n);
  return check_launch(c, "add_f32_kernel");
}

int launch_scale_u64_f32(spf_ctx* c, const uint64_t* src, float scale, uint64_t n, float* dst) {
  if (n == 0) return SPF_OK;
  scale_u64_f32_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, c->stream>>>(src, scale, n, dst);
  return check_launch(c, "scale_u64_f32_kernel");
}

int launch_fill_u64(spf_ctx* c, uint64_t* p, uint64_t n, uint64_t v) {
  if (n == 0) return SPF_OK;
  fill_u64_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, c->stream>>>(p, n, v);
  return check_launch(c, "fill_u64_kernel");
}

int launch_max_f32(spf_ctx* c, const float* p, uint64_t n, float* out1) {
  max_f32_kernel<<<1, 1024, 0, c->stream>>>(p, n, out1);
  return check_launch(c, "max_f32_kernel");
}

__global__ void rows_equal_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint64_t n, int* flag) {
  bool diff = false;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    diff = diff || a[i] != b[i];
  if (__syncthreads_or(diff) && threadIdx.x == 0) *flag = 0;
}

int launch_rows_equal(spf_ctx* c, const float* a, const float* b, uint64_t n, int* d_flag) {
  if (n == 0) return SPF_OK;
  uint64_t blocks = ceil_div(n, 256 * 8);
  if (blocks > (uint64_t)c->sm_count * 8) blocks = (uint64_t)c->sm_count * 8;
  rows_equal_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(reinterpret_cast<const uint32_t*>(a),
                                                             reinterpret_cast<const uint32_t*>(b), n, d_flag);
  return check_launch(c, "rows_equal_kernel");
}

int launch_check_rows(spf_ctx* c, const uint64_t* d_idx, uint64_t m, uint64_t n, int* d_flag) {
  if (m == 0) return SPF_OK;
  check_rows_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, c->stream>>>(d_idx, m, n, d_flag);
  return check_launch(c, "check_rows_kernel");
}

int launch_pair_dist(spf_ctx* c, int metric, const float* A, uint32_t ldA, const uint64_t* idxA,
                     const float* B, uint32_t ldB, const uint32_t* idxB32, uint64_t fixedB,
                     uint32_t ld, uint64_t count, float* out) {
  if (count == 0) return SPF_OK;
  uint64_t blocks = ceil_div(count, PD_THREADS);   // 32 pairs per warp, 2 warps per block
  if (blocks > (uint64_t)c->sm_count * 32) blocks = (uint64_t)c->sm_count * 32;
  dim3 grid((unsigned)blocks), block(PD_THREADS);
  switch (metric) {
    case SPF_METRIC_EUCLIDEAN:
      pair_dist_kernel<SPF_METRIC_EUCLIDEAN><<<grid, block, 0, c->stream>>>(A, ldA, idxA, B, ldB, idxB32, fixedB, ld, count, out);
      break;
    case SPF_METRIC_MANHATTAN:
      pair_dist_kernel<SPF_METRIC_MANHATTAN><<<grid, block, 0, c->stream>>>(A, ldA, idxA, B, ldB, idxB32, fixedB, ld, count, out);
      break;
    case SPF_METRIC_CHEBYSHEV:
      pair_dist_kernel<SPF_METRIC_CHEBYSHEV><<<grid, block, 0, c->stream>>>(A, ldA, idxA, B, ldB, idxB32, fixedB, ld, count, out);
      break;
    default:
      return fail(SPF_E_INVALID, "unknown metric %d", metric);
  }
  return check_launch(c, "pair_dist_kernel");
}

}  // namespace spf
