// pairdist.cuh — warp-cooperative exact distance of 32 (row, row) pairs, one pair per lane.
//
// The reference evaluates DistanceMetric::compute (src/distances/distance.rs:16-43) as one
// sequential f32 chain per pair.  A lane therefore owns one pair and walks its dimensions in
// order; the warp only cooperates on the memory side: for every pair the 32 lanes copy a
// 64-dimension chunk of both rows into shared memory with cp.async (lanes 0-15 the first row,
// lanes 16-31 the second, 16 bytes each — one coalesced 256-byte request per row chunk, all 64
// requests of a step in flight together, no registers held), then each lane reads its own pair's
// values back conflict-free.
#pragma once
#include "common.cuh"

namespace spf {

constexpr int PD_THREADS = 64;
constexpr int PD_CHUNK = 64;              // dimensions staged per step
constexpr int PD_STRIDE = PD_CHUNK + 4;   // floats per staged row (16-byte aligned, conflict-free)

struct PairDistSmem {
  float ta[32][PD_STRIDE];
  float tb[32][PD_STRIDE];
};

// pa / pb: this lane's two rows (nullptr for an inactive lane), 16-byte aligned, ld a multiple of 4.
// All 32 lanes must call.
template <int METRIC>
__device__ __forceinline__ float warp_pair_dist(const float* pa, const float* pb, uint32_t ld,
                                                PairDistSmem& s) {
  const int lane = threadIdx.x & 31;
  const int second = lane >> 4, l16 = lane & 15;
  const unsigned long long ua = (unsigned long long)(uintptr_t)pa, ub = (unsigned long long)(uintptr_t)pb;
  float acc = 0.0f;
  for (uint32_t c0 = 0; c0 < ld; c0 += PD_CHUNK) {
    const uint32_t col = c0 + l16 * 4;
    const bool col_ok = col < ld;
#pragma unroll 8
    for (int p = 0; p < 32; ++p) {
      const unsigned long long qa = __shfl_sync(0xffffffffu, ua, p);
      const unsigned long long qb = __shfl_sync(0xffffffffu, ub, p);
      if (qa != 0 && col_ok) {
        const float* src = reinterpret_cast<const float*>((uintptr_t)(second ? qb : qa)) + col;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(second ? &s.tb[p][l16 * 4] : &s.ta[p][l16 * 4]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (pa != nullptr) {
      const int n4 = (int)(((ld - c0) < (uint32_t)PD_CHUNK ? (ld - c0) : (uint32_t)PD_CHUNK) >> 2);
      const float4* xp = reinterpret_cast<const float4*>(&s.ta[lane][0]);
      const float4* yp = reinterpret_cast<const float4*>(&s.tb[lane][0]);
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
  return acc;
}

// One row per lane against a vector the lane reads straight from global memory (the medoid pass:
// the second operand is the mean of the lane's cluster — 32 consecutive members nearly always
// share it, so the read is a broadcast served by L1 — and staging it per pair doubled the copies).
// Only the member rows are staged: the whole warp copies 128 dimensions of ONE row per cp.async
// instruction (32 lanes x 16 bytes = one contiguous 512-byte request), 32 rows in flight, then every
// lane walks its own row in dimension order — the same sequential f32 chain as warp_pair_dist.
constexpr int RD_CHUNK = 128;             // dimensions staged per step
constexpr int RD_STRIDE = RD_CHUNK + 4;   // floats per staged row: (33 l + i) mod 8 distinct over 8 lanes -> conflict-free LDS.128

struct RowDistSmem {
  float t[32][RD_STRIDE];
};

// pa: this lane's staged row, pb: the vector it is compared with (both nullptr for an inactive lane),
// 16-byte aligned, ld a multiple of 4; pb must not be written during the kernel.  All 32 lanes must call.
template <int METRIC>
__device__ __forceinline__ float warp_row_dist(const float* pa, const float* pb, uint32_t ld, RowDistSmem& s) {
  const int lane = threadIdx.x & 31;
  const unsigned long long ua = (unsigned long long)(uintptr_t)pa;
  float acc = 0.0f;
  for (uint32_t c0 = 0; c0 < ld; c0 += RD_CHUNK) {
    const uint32_t col = c0 + lane * 4;
    const bool col_ok = col < ld;
#pragma unroll 8
    for (int p = 0; p < 32; ++p) {
      const unsigned long long qa = __shfl_sync(0xffffffffu, ua, p);
      if (qa != 0 && col_ok) {
        const float* src = reinterpret_cast<const float*>((uintptr_t)qa) + col;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s.t[p][lane * 4]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (pa != nullptr) {
      const int n4 = (int)(((ld - c0) < (uint32_t)RD_CHUNK ? (ld - c0) : (uint32_t)RD_CHUNK) >> 2);
      const float4* xp = reinterpret_cast<const float4*>(&s.t[lane][0]);
      const float4* yp = reinterpret_cast<const float4*>(pb + c0);
#pragma unroll 8
      for (int i = 0; i < n4; ++i) {
        const float4 xv = xp[i];
        const float4 yv = __ldg(yp + i);
        acc = dist_step<METRIC>(acc, xv.x, yv.x);
        acc = dist_step<METRIC>(acc, xv.y, yv.y);
        acc = dist_step<METRIC>(acc, xv.z, yv.z);
        acc = dist_step<METRIC>(acc, xv.w, yv.w);
      }
    }
    __syncwarp();
  }
  return acc;
}

// Plain per-thread sequential distance (used where only a handful of pairs are needed).
template <int METRIC>
__device__ __forceinline__ float thread_dist(const float* __restrict__ a, const float* __restrict__ b,
                                             uint32_t ld) {
  float acc = 0.0f;
  for (uint32_t i = 0; i < ld; i += 4) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(a + i));
    const float4 y = __ldg(reinterpret_cast<const float4*>(b + i));
    acc = dist_step<METRIC>(acc, x.x, y.x);
    acc = dist_step<METRIC>(acc, x.y, y.y);
    acc = dist_step<METRIC>(acc, x.z, y.z);
    acc = dist_step<METRIC>(acc, x.w, y.w);
  }
  return acc;
}

}  // namespace spf
