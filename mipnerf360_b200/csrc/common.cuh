// Shared device/host helpers for the per-ray kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mip360_b200.h"

#define FULL_MASK 0xffffffffu

namespace mip360 {

void set_error(const char* fmt, ...);
// runtime switches between kernel variants (mip360_set_option): every variant computes the same function, the
// switches exist so that tests can compare them against each other
enum { OPT_RAY_GROUP = 0, OPT_CTA_PAIR = 1, OPT_SHORT_K = 2, OPT_PACKED_EPILOGUE = 3, OPT_FUSED_NARROW = 4, OPT_COUNT = 5 };
bool option(int key);
void count_launch(int n = 1);
constexpr int MAX_DEVICES = 64;
int current_device();  // cudaGetDevice, clamped to [0, MAX_DEVICES)
int sm_count();        // SMs of the current device

#define MIP_REQUIRE(cond, ...)                  \
  do {                                          \
    if (!(cond)) {                              \
      mip360::set_error(__VA_ARGS__);           \
      return MIP360_ERR_ARG;                    \
    }                                           \
  } while (0)

#define MIP_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      mip360::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
      return MIP360_ERR_CUDA;                                                            \
    }                                                                                    \
  } while (0)

#define MIP_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    mip360::count_launch();                                                              \
    cudaError_t e_ = cudaGetLastError();                                                 \
    if (e_ != cudaSuccess) {                                                             \
      mip360::set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return MIP360_ERR_CUDA;                                                            \
    }                                                                                    \
  } while (0)

// ---- warp primitives -------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
// inclusive scan across the 32 lanes
__device__ __forceinline__ float warp_scan_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(FULL_MASK, v, o);
    if (lane >= o) v += n;
  }
  return v;
}
// inclusive suffix scan (sum over lanes >= lane)
__device__ __forceinline__ float warp_scan_incl_rev(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_down_sync(FULL_MASK, v, o);
    if (lane + o < 32) v += n;
  }
  return v;
}

// Block-level sum of one double per thread into partials[blockIdx.x] (deterministic order).
template <int THREADS>
__device__ __forceinline__ void block_sum_to_partial(double v, double* partials) {
  __shared__ double sm_part[THREADS / 32];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sm_part[w] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) s += sm_part[i];
    partials[blockIdx.x] = s;
  }
}

__device__ __forceinline__ float softplus_f(float x) {
  // torch.nn.Softplus(beta=1, threshold=20): x if x > 20 else log1p(exp(x))
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float(uint32_t(b) << 16); }

// torch.nan_to_num defaults: nan -> 0 (or given), +inf -> FLT_MAX, -inf -> -FLT_MAX
__device__ __forceinline__ float nan_to_num_f(float x, float nan_val = 0.f) {
  if (isnan(x)) return nan_val;
  if (isinf(x)) return x > 0 ? 3.402823466e+38f : -3.402823466e+38f;
  return x;
}

constexpr float G_EPS = 1e-6f;  // intern/parameterization.py:19

// ---- counter-based random numbers (Philox4x32-10, Salmon et al. 2011) ------------------------------------------
// The randomized draws of the reference (torch.rand of ray.py:106, uniform_ of ray.py:33) are generated inside the
// kernels that consume them instead of being written to HBM and read back.  Uniform number m of ray r in a draw:
//   counter = (r, block(m), stream id, epoch), key = 64-bit seed, u = (output word word(m) >> 8) * 2^-24 in [0, 1)
//   block(m) = (m & 7) | ((m >> 5) << 3),  word(m) = (m >> 3) & 3
// i.e. one Philox call yields the numbers m, m + 8, m + 16, m + 24 of a ray: exactly the samples one lane of an 8-lane ray
// group owns (ray_group.cuh), so a lane needs one call per four samples instead of one per sample.
// `stream id` distinguishes the call sites of one iteration (host counter), `epoch` is read from device memory so that
// a captured CUDA graph draws fresh numbers on every replay.  (The test suite carries a NumPy restatement, pinned to the
// published Random123 known-answer vectors, that reproduces these uniforms bit for bit.)
struct RngArgs {
  unsigned long long seed;
  const unsigned long long* epoch;  // device counter, may be null (epoch 0)
  unsigned int stream_id;
  int enabled;
};
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, unsigned long long seed,
                                              uint32_t (&out)[4]) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float rng_word_to_uniform(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-08f; }
__device__ __forceinline__ uint32_t rng_epoch(const RngArgs& r) {
  return r.epoch ? (uint32_t)(*r.epoch) : 0u;
}
// the four uniforms m0, m0 + 8, m0 + 16, m0 + 24 (m0 = (block & 7) + 32 * (block >> 3)) of ray `ray`
__device__ __forceinline__ void rng_uniform4(const RngArgs& r, uint32_t epoch, uint32_t ray, uint32_t block, float (&u)[4]) {
  uint32_t w[4];
  philox4x32_10(ray, block, r.stream_id, epoch, r.seed, w);
#pragma unroll
  for (int i = 0; i < 4; ++i) u[i] = rng_word_to_uniform(w[i]);
}
// uniform number m of ray `ray` (one call per number: the generic kernels)
__device__ __forceinline__ float rng_uniform(const RngArgs& r, uint32_t epoch, uint32_t ray, uint32_t m) {
  uint32_t w[4];
  philox4x32_10(ray, (m & 7u) | ((m >> 5) << 3), r.stream_id, epoch, r.seed, w);
  const uint32_t sel = (m >> 3) & 3u;
  return rng_word_to_uniform(sel == 0 ? w[0] : sel == 1 ? w[1] : sel == 2 ? w[2] : w[3]);
}

}  // namespace mip360
