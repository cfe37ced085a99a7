// Sub-warp ray groups for the per-ray scan kernels when N is 32, 64 or 128: 8 lanes own one ray, each lane keeps
// E = N/8 contiguous intervals in registers (4 rays per warp, 16 per 128-thread block).  Rows are read with
// 16-byte vector loads (every lane touches whole 32-byte sectors), prefix sums are lane-local loops plus a
// 3-step shuffle scan inside the group.  Other N fall back to the generic one-warp-per-ray kernels.
#pragma once
#include "common.cuh"

namespace mip360 {

constexpr int RG_LANES = 8;
constexpr int RG_THREADS = 128;
constexpr int RG_RAYS_PER_BLOCK = RG_THREADS / RG_LANES;

__device__ __forceinline__ bool rg_supported(int N) { return N == 32 || N == 64 || N == 128; }
static inline bool rg_supported_host(int N) { return (N == 32 || N == 64 || N == 128) && option(OPT_RAY_GROUP); }

template <typename T>
__device__ __forceinline__ T rg_sum(T v) {
  v += __shfl_xor_sync(FULL_MASK, v, 4);
  v += __shfl_xor_sync(FULL_MASK, v, 2);
  v += __shfl_xor_sync(FULL_MASK, v, 1);
  return v;
}
// exclusive prefix over the 8 lanes of a group (gl = lane & 7)
template <typename T>
__device__ __forceinline__ T rg_scan_excl(T v, int gl) {
  T x = v;
#pragma unroll
  for (int o = 1; o < RG_LANES; o <<= 1) {
    const T n = __shfl_up_sync(FULL_MASK, x, o, RG_LANES);
    if (gl >= o) x += n;
  }
  return x - v;
}
// exclusive suffix (sum over lanes of the group with a larger index)
template <typename T>
__device__ __forceinline__ T rg_scan_excl_rev(T v, int gl) {
  T x = v;
#pragma unroll
  for (int o = 1; o < RG_LANES; o <<= 1) {
    const T n = __shfl_down_sync(FULL_MASK, x, o, RG_LANES);
    if (gl + o < RG_LANES) x += n;
  }
  return x - v;
}

// E contiguous floats (E in {4, 8, 16}; p 16-byte aligned)
template <int E>
__device__ __forceinline__ void rg_load(const float* __restrict__ p, float (&v)[E]) {
#pragma unroll
  for (int i = 0; i < E / 4; ++i) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(p) + i);
    v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
  }
}
template <int E>
__device__ __forceinline__ void rg_store(float* __restrict__ p, const float (&v)[E]) {
#pragma unroll
  for (int i = 0; i < E / 4; ++i)
    reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
// E + 1 knots t[j0 .. j0 + E] of a row with N + 1 entries (rows are not 16-byte aligned: scalar, L1-resident)
template <int E>
__device__ __forceinline__ void rg_load_knots(const float* __restrict__ row, int j0, float (&t)[E + 1]) {
#pragma unroll
  for (int i = 0; i <= E; ++i) t[i] = __ldg(row + j0 + i);
}

// The lane's E + 1 uniforms of a draw (common.cuh): slot c < E is number gl + 8c of the ray, slot E is number N = 8E (used
// by lane 0 only).  E / 4 Philox calls for the first E slots (at least one), one more for slot E.
template <int E>
__device__ __forceinline__ void rg_draw(const RngArgs& rng, uint32_t epoch, uint32_t ray, int gl, float (&u)[E + 1]) {
#pragma unroll
  for (int j = 0; j < (E + 3) / 4; ++j) {
    float q[4];
    rng_uniform4(rng, epoch, ray, (uint32_t)(gl + 8 * j), q);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (4 * j + i < E) u[4 * j + i] = q[i];
  }
  // number N = 8E: lane 0 of block 8 * (E / 4), word E & 3 (the whole warp executes the call, lane 0 uses it)
  float q[4];
  rng_uniform4(rng, epoch, ray, (uint32_t)(8 * (E / 4)), q);
  u[E] = (E & 3) == 0 ? q[0] : q[E & 3];
}

static inline int rg_grid(int B) {
  long long b = ((long long)B + RG_RAYS_PER_BLOCK - 1) / RG_RAYS_PER_BLOCK;
  const long long cap = (long long)sm_count() * 32;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace mip360
