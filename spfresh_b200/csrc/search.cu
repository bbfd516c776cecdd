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

struct ScanArgs {
  const float* vecs; const uint64_t* slot_ids; const uint64_t* grp_off; const uint32_t* lens;
  uint32_t ld; uint32_t d;
  const float* Q; const uint32_t* probe; const float* thr; const uint32_t* seqbase;
  uint32_t nprobe; uint32_t K;
  uint64_t* out_ids; float* out_dists; uint32_t* out_counts; unsigned long long* out_keys;
  unsigned long long* out_slots;   // nq x K slot index of each result (vector gather)
  unsigned long long* bytes;
};

constexpr int SCAN_WARPS = 8;

// One CTA per query.  Warps take (probed list, group) units round-robin; a lane computes the
// exact squared L2 of its slot's vector (spann_index.rs:172), keeps it if <= thr (:176) and
// the warp maintains the K smallest (distance, encounter index) keys.
template <int R>
__global__ void __launch_bounds__(SCAN_WARPS * 32)
scan_kernel(ScanArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_q = reinterpret_cast<float4*>(smem_raw);                       // ld/4 float4
  uint32_t* s_gpre = reinterpret_cast<uint32_t*>(s_q + a.ld / 4);          // nprobe+1
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(
      (reinterpret_cast<uintptr_t>(s_gpre + a.nprobe + 1) + 7) & ~(uintptr_t)7);   // SCAN_WARPS*K
  unsigned long long* s_pay = s_keys + SCAN_WARPS * a.K;

  const uint64_t q = blockIdx.x;
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
    if (bytes) atomicAdd(a.bytes, bytes);
  }
  __syncthreads();
  const float thr = a.thr[q];
  const uint32_t units = s_gpre[a.nprobe];

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
    unsigned bal = __ballot_sync(0xffffffffu, valid && acc <= thr && ck < kth);
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
      a.out_ids[q * a.K + e] = ok ? a.slot_ids[pay[r]] : ~0ull;
      a.out_dists[q * a.K + e] = ok ? __uint_as_float((uint32_t)(key[r] >> 32)) : __int_as_float(0x7f800000);
      a.out_keys[q * a.K + e] = key[r];
      a.out_slots[q * a.K + e] = ok ? pay[r] : ~0ull;
    }
  }
  if (lane == 0) a.out_counts[q] = count;
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

}  // namespace
}  // namespace spf

extern "C" {

int spf_index_pack(spf_dataset* ds, const uint64_t* offsets, const uint64_t* members,
                   const uint64_t* centroid_rows, uint32_t nlists, uint32_t list_begin, uint32_t list_end,
                   spf_index** out) {
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
}

int spf_index_load_dir(spf_ctx* c, const char* dir, const float* centroids, uint32_t nlists, uint32_t d,
                       spf_index** out) {
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
}

int spf_index_save_dir(const spf_index* idx, const char* dir) {
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
  buf.clear();
  put_u64(buf, ids_written.size());
  for (uint64_t id : ids_written) put_u64(buf, id);
  FILE* f = fopen((base + "/cluster_ids.bin").c_str(), "wb");
  if (!f || fwrite(buf.data(), 1, buf.size(), f) != buf.size()) {
    if (f) fclose(f);
    return fail(SPF_E_IO, "cannot write %s/cluster_ids.bin", dir);
  }
  fclose(f);
  return SPF_OK;
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
  delete idx;
}

uint32_t spf_index_lists(const spf_index* idx) { return idx ? idx->nlists : 0; }
uint64_t spf_index_vectors(const spf_index* idx) { return idx ? idx->total_vectors : 0; }
uint64_t spf_index_last_scan_bytes(const spf_index* idx) { return idx ? idx->last_scan_bytes : 0; }

int spf_search_batch(spf_index* idx, const float* queries, uint64_t nq, uint32_t k, uint32_t nprobe,
                     float prune_factor, uint64_t* ids, float* dists, uint32_t* counts, float* vectors,
                     uint64_t* keys) {
  if (!idx || !queries || !ids || !dists || !counts) return fail(SPF_E_INVALID, "spf_search_batch: NULL argument");
  if (k == 0 || k > 128) return fail(SPF_E_INVALID, "k must be in [1,128]");
  if (nq == 0) return SPF_OK;
  if (nprobe == 0) nprobe = k;                       // spann_index.rs:164 nearest_n(query, k)
  if (nprobe > idx->nlists) nprobe = idx->nlists;
  if (nprobe > 1024) return fail(SPF_E_INVALID, "nprobe must be <= 1024");
  spf_ctx* c = idx->ctx;
  std::lock_guard<std::mutex> lk(c->mu);
  SPF_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
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
  SPF_CUDA(cudaMemcpy2DAsync(Q.p, (size_t)ld * 4, queries, (size_t)d * 4, (size_t)d * 4, nq, cudaMemcpyHostToDevice, st));

  // centroid probe in query chunks (dense nq_chunk x nlists exact distances, then selection)
  uint64_t chunk = (256ull << 20) / nlists;
  if (chunk == 0) chunk = 1;
  if (chunk > nq) chunk = nq;
  DevBuf<float> Dqc;
  SPF_TRY(Dqc.alloc(st, (size_t)chunk * nlists));
  {
    KernelTimer t(c, "probe");
    for (uint64_t q0 = 0; q0 < nq; q0 += chunk) {
      const uint64_t nc = (nq - q0) < chunk ? (nq - q0) : chunk;
      SPF_TRY(launch_assign_exact(c, SPF_METRIC_EUCLIDEAN, Q.p + q0 * ld, nc, idx->centroids, nlists, ld, 1.0f,
                                  nullptr, Dqc.p));
      probe_select_kernel<<<(unsigned)nc, 128, 0, st>>>(Dqc.p, nlists, nprobe, prune_factor, idx->lens,
                                                        probe.p + q0 * nprobe, thr.p + q0, seqbase.p + q0 * nprobe);
      SPF_TRY(check_launch(c, "probe_select_kernel"));
    }
  }
  ScanArgs a;
  a.vecs = idx->vecs; a.slot_ids = idx->slot_ids; a.grp_off = idx->grp_off; a.lens = idx->lens;
  a.ld = ld; a.d = d; a.Q = Q.p; a.probe = probe.p; a.thr = thr.p; a.seqbase = seqbase.p;
  a.nprobe = nprobe; a.K = k;
  a.out_ids = o_ids.p; a.out_dists = o_dists.p; a.out_counts = o_counts.p; a.out_keys = o_keys.p;
  a.out_slots = o_slots.p; a.bytes = d_bytes.p;
  {
    KernelTimer t(c, "scan");
    if (k <= 32) SPF_TRY(launch_scan<1>(c, a, nq));
    else if (k <= 64) SPF_TRY(launch_scan<2>(c, a, nq));
    else SPF_TRY(launch_scan<4>(c, a, nq));
  }
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
}

// Host-side merge of per-rank partial top-k (the reference's final stable sort, spann_index.rs:
// 188-193, applied across list shards).  Keys are globally consistent, so the k smallest win.
int spf_topk_merge(uint32_t parts, uint64_t nq, uint32_t k, const uint64_t* keys, const uint64_t* ids,
                   const float* dists, const uint32_t* counts, uint64_t* out_ids, float* out_dists,
                   uint32_t* out_counts) {
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
}

}  // extern "C"
