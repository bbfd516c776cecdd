// pairdist.cuh — warp-cooperative exact distance of 32 (row, row) pairs, one pair per lane.
//
// The reference evaluates DistanceMetric::compute (src/distances/distance.rs:16-43) as one
// sequential f32 chain per pair.  A lane therefore owns one pair and walks its dimensions in
// order; the warp only cooperates on the memory side: the 32 lanes load 32 consecutive floats
// of every pair's two rows (one coalesced 128-byte request per row chunk) into a padded shared
// tile, then each lane reads its own pair's values back conflict-free.
#pragma once
#include "common.cuh"

namespace spf {

constexpr int PD_THREADS = 128;

struct PairDistSmem {
  float ta[32][33];
  float tb[32][33];
  const float* pa[32];
  const float* pb[32];
};

// pa / pb: this lane's two rows (nullptr for an inactive lane).  All 32 lanes must call.
template <int METRIC>
__device__ __forceinline__ float warp_pair_dist(const float* pa, const float* pb, uint32_t ld,
                                                PairDistSmem& s) {
  const int lane = threadIdx.x & 31;
  s.pa[lane] = pa;
  s.pb[lane] = pb;
  __syncwarp();
  float acc = 0.0f;
  for (uint32_t chunk = 0; chunk < ld; chunk += 32) {
    const uint32_t col = chunk + lane;
    const bool col_ok = col < ld;
#pragma unroll 8
    for (int p = 0; p < 32; ++p) {
      const float* qa = s.pa[p];
      const float* qb = s.pb[p];
      float va = 0.f, vb = 0.f;
      if (qa != nullptr && col_ok) {
        va = __ldg(qa + col);
        vb = __ldg(qb + col);
      }
      s.ta[p][lane] = va;
      s.tb[p][lane] = vb;
    }
    __syncwarp();
    const int nn = (ld - chunk) < 32u ? (int)(ld - chunk) : 32;
    if (nn == 32) {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc = dist_step<METRIC>(acc, s.ta[lane][i], s.tb[lane][i]);
    } else {
      for (int i = 0; i < nn; ++i) acc = dist_step<METRIC>(acc, s.ta[lane][i], s.tb[lane][i]);
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
