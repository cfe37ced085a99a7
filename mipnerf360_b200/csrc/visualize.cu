// K7: depth / normal visualisation of a rendered frame (intern/pose.py:112-212), SURVEY §8f rank 4.
// Image-space, HBM-bound: every kernel streams depth (+ acc) once and writes the RGB picture, optionally already
// as uint8 (utils.py:17-21), so a frame leaves the device as 3 B/pixel.
//   normals : masked variances of (x, y, depth) in two deterministic fp64 passes -> isotropic scaling ->
//             3x3 blur/edge stencils (scipy convolve2d 'same', zero fill) in fp64 -> shading + acc blend
//   depth   : optional automatic near/far = acc-weighted quantiles of depth.  The reference argsorts the frame
//             and cumsums acc; here the two quantile keys are found by a 32-step bitwise bisection over the
//             order-preserving integer image of the depths with exact integer weights (acc in 2^-24 units), no
//             sort and no host round trip; then curve, normalise, colour map (LUT or sinebow), acc blend in fp32.
#include <float.h>

#include "common.cuh"

namespace mip360 {

constexpr int VIS_THREADS = 256;
constexpr int VIS_MAX_BLOCKS = 1024;  // partial sums per reduced quantity
enum { CURVE_NEG_LOG = 0, CURVE_IDENTITY = 1, CURVE_INV = 2, CURVE_LOG = 3 };

static inline int vis_grid(long long n) {
  long long b = (n + VIS_THREADS - 1) / VIS_THREADS;
  const long long cap = (long long)sm_count() * 4 < VIS_MAX_BLOCKS ? (long long)sm_count() * 4 : VIS_MAX_BLOCKS;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

__device__ __forceinline__ uint8_t to8b_f(float v) {  // utils.py:17-21
  return (uint8_t)(255.f * fminf(fmaxf(nan_to_num_f(v), 0.f), 1.f));
}

// ---- normals: masked variances ---------------------------------------------------------------
// stats: [0] count  [1..3] mean x, y, z  [4..6] var x, y, z  [7] scaling = sqrt(((var x + var y) / 2) / var z)
template <int PASS>
__global__ void __launch_bounds__(VIS_THREADS) depth_moments_kernel(const float* __restrict__ depth, int H, int W,
                                                                     const double* __restrict__ stats,
                                                                     double* __restrict__ partials) {
  const long long n = (long long)H * W;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  double mx = 0.0, my = 0.0, mz = 0.0;
  if (PASS == 1) { mx = stats[1]; my = stats[2]; mz = stats[3]; }
  for (long long i = (long long)blockIdx.x * VIS_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * VIS_THREADS) {
    const float z = depth[i];
    if (isnan(z)) continue;  // pose.py:131 mask = ~isnan(depth)
    const double x = (double)(i % W), y = (double)(i / W);
    if (PASS == 0) {
      a0 += 1.0; a1 += x; a2 += y; a3 += (double)z;
    } else {
      a1 += (x - mx) * (x - mx); a2 += (y - my) * (y - my); a3 += ((double)z - mz) * ((double)z - mz);
    }
  }
  __shared__ double sm[4][VIS_THREADS / 32];
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sm[0][w] = a0; sm[1][w] = a1; sm[2][w] = a2; sm[3][w] = a3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int i = 0; i < VIS_THREADS / 32; ++i) s += sm[threadIdx.x][i];
    partials[threadIdx.x * VIS_MAX_BLOCKS + blockIdx.x] = s;
  }
}

template <int PASS>
__global__ void __launch_bounds__(32) depth_moments_reduce_kernel(const double* __restrict__ partials, int nblocks,
                                                                   double* __restrict__ stats) {
  // one warp, fixed order: lane-strided partial sums, then the shuffle tree
  double s[4];
  for (int q = 0; q < 4; ++q) {
    double a = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 32) a += partials[q * VIS_MAX_BLOCKS + i];
    s[q] = warp_sum(a);
  }
  if (threadIdx.x == 0) {
    if (PASS == 0) {
      stats[0] = s[0];
      stats[1] = s[1] / s[0]; stats[2] = s[2] / s[0]; stats[3] = s[3] / s[0];
    } else {
      const double cnt = stats[0];
      stats[4] = s[1] / cnt; stats[5] = s[2] / cnt; stats[6] = s[3] / cnt;
      stats[7] = sqrt(((stats[4] + stats[5]) * 0.5) / stats[6]);  // pose.py:134-136
    }
  }
}

// Pictures are written four pixels per thread: 12 bytes (uint8) or 48 bytes (fp32) of RGB per thread go out as
// three 32-bit / 128-bit stores instead of twelve byte / word stores.
constexpr int VIS_PPT = 4;

__device__ __forceinline__ void store_rgb4(float* __restrict__ vis, uint8_t* __restrict__ vis8, long long i0, long long n,
                                           const float (&v)[VIS_PPT][3], const uint8_t (&b)[VIS_PPT][3]) {
  if (i0 + VIS_PPT <= n) {
    if (vis) {
      float4* o = reinterpret_cast<float4*>(vis + 3 * i0);  // 48-byte groups: 16-byte aligned with the tensor
      o[0] = make_float4(v[0][0], v[0][1], v[0][2], v[1][0]);
      o[1] = make_float4(v[1][1], v[1][2], v[2][0], v[2][1]);
      o[2] = make_float4(v[2][2], v[3][0], v[3][1], v[3][2]);
    }
    if (vis8) {
      uint32_t* o = reinterpret_cast<uint32_t*>(vis8 + 3 * i0);
      o[0] = b[0][0] | (b[0][1] << 8) | (b[0][2] << 16) | ((uint32_t)b[1][0] << 24);
      o[1] = b[1][1] | (b[1][2] << 8) | (b[2][0] << 16) | ((uint32_t)b[2][1] << 24);
      o[2] = b[2][2] | (b[3][0] << 8) | (b[3][1] << 16) | ((uint32_t)b[3][2] << 24);
    }
  } else {
    for (int p = 0; p < VIS_PPT && i0 + p < n; ++p)
      for (int c = 0; c < 3; ++c) {
        if (vis) vis[3 * (i0 + p) + c] = v[p][c];
        if (vis8) vis8[3 * (i0 + p) + c] = b[p][c];
      }
  }
}

// pose.py:112-121 (depth_to_normals on scaling * depth) + :138-145 (shading, white where nothing accumulated)
__global__ void __launch_bounds__(VIS_THREADS) normals_kernel(const float* __restrict__ depth, const float* __restrict__ acc,
                                                               const double* __restrict__ stats, int H, int W,
                                                               float* __restrict__ vis, uint8_t* __restrict__ vis8) {
  const long long n = (long long)H * W;
  const long long i0 = ((long long)blockIdx.x * VIS_THREADS + threadIdx.x) * VIS_PPT;
  if (i0 >= n) return;
  const double sc = stats[7];
  float vf[VIS_PPT][3];
  uint8_t vb[VIS_PPT][3];
#pragma unroll
  for (int p = 0; p < VIS_PPT; ++p) {
    const long long i = i0 + p < n ? i0 + p : n - 1;
    const int x = (int)(i % W), y = (int)(i / W);
    double z[3][3];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int yy = y + dy, xx = x + dx;
        z[dy + 1][dx + 1] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? sc * (double)__ldg(depth + (long long)yy * W + xx) : 0.0;
      }
    // true convolution (kernel flipped): out = sum_ij k[i][j] z[2-i][2-j] with k_dy = edge (rows) x blur (cols) and
    // k_dx = blur (rows) x edge (cols).  The zero taps are kept: 0 * NaN = NaN, so a NaN depth poisons its whole 3x3
    // neighbourhood in both derivatives, exactly as scipy's convolve2d does.
    const double fe[3] = {-0.5, 0.0, 0.5}, fb[3] = {0.25, 0.5, 0.25};
    double gy = 0.0, gx = 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        gy += (fb[j] * fe[r]) * z[2 - r][2 - j];
        gx += (fb[r] * fe[j]) * z[2 - r][2 - j];
      }
    const double inv = 1.0 / sqrt(1.0 + gx * gx + gy * gy);
    const double nrm[3] = {gx * inv, gy * inv, inv};
    // pose.py:143: vis * acc + (1 - acc) — (1 - acc) is formed in float32 (acc's dtype), then promoted
    const double a = acc ? (double)acc[i] : 1.0, one_minus_a = acc ? (double)(1.f - acc[i]) : 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double v = (nrm[c] + 1.0) * 0.5;
      if (isnan(nrm[c])) v = 1.0;  // isnan(normals) + nan_to_num(...): NaN -> 1 + 0
      else if (isinf(v)) v = v > 0 ? DBL_MAX : -DBL_MAX;
      if (acc) v = v * a + one_minus_a;
      vf[p][c] = (float)v;
      double u = isnan(v) ? 0.0 : v;
      u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
      vb[p][c] = (uint8_t)(255.0 * u);
    }
  }
  store_rgb4(vis, vis8, i0, n, vf, vb);
}

// ---- depth: automatic near / far ----------------------------------------------------------------
// order-preserving integer image of a float; every NaN maps to the largest key (np.argsort puts NaN last)
__device__ __forceinline__ uint32_t depth_key(float z) {
  if (isnan(z)) return 0xFFFFFFFFu;
  const uint32_t u = __float_as_uint(z);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_depth(uint32_t k) {
  if (k == 0xFFFFFFFFu) return __uint_as_float(0x7FC00000u);
  return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
// acc in exact integer units of 2^-24; pixels with NaN depth carry no weight (pose.py:178)
__device__ __forceinline__ unsigned long long depth_weight(float z, float a) {
  if (isnan(z) || !(a > 0.f)) return 0ull;
  return __float2ull_rn(fminf(a, 1024.f) * 16777216.f);
}

// work layout (unsigned long long): [0] total weight, [1 + 2p], [2 + 2p] = S(candidate) of bisection step p for the
// low / high quantile, [65] min key >= k_lo, [66] max key <= k_hi
constexpr int QW_TOTAL = 0, QW_STEP = 1, QW_KMIN = 65, QW_KMAX = 66, QW_LEN = 67;

struct QuantileKeys { uint32_t k_lo, k_hi, c_lo, c_hi; };
// Replays the first `steps` bisection decisions from the recorded sums (identical in every block).
//   k_lo = min { k : S(k) >= total * frac }          built from the top bit down, candidate = k | (2^bit - 1)
//   k_hi = max { k : S(k) <= total * (1 - frac) }    candidate = k | 2^bit
__device__ __forceinline__ QuantileKeys replay_bisection(const unsigned long long* work, int steps, double frac) {
  const double total = (double)work[QW_TOTAL];
  const double thr_lo = total * frac, thr_hi = total * (1.0 - frac);
  QuantileKeys q{0u, 0u, 0u, 0u};
  for (int p = 0; p <= steps && p <= 32; ++p) {
    const int bit = 31 - p;
    if (p > 0) {  // decision of step p-1 from its recorded sums
      if (!((double)work[QW_STEP + 2 * (p - 1)] >= thr_lo)) q.k_lo |= 1u << (bit + 1);
      if ((double)work[QW_STEP + 2 * (p - 1) + 1] <= thr_hi) q.k_hi |= 1u << (bit + 1);
    }
    if (p < 32) {
      q.c_lo = q.k_lo | ((1u << bit) - 1u);
      q.c_hi = q.k_hi | (1u << bit);
    }
  }
  return q;
}

// step < 0: total weight; 0..31: one bisection step; 32: nearest existing keys
__global__ void __launch_bounds__(VIS_THREADS) depth_quantile_kernel(const float* __restrict__ depth, const float* __restrict__ acc,
                                                                      long long n, double frac, int step,
                                                                      unsigned long long* __restrict__ work) {
  __shared__ QuantileKeys qs;
  __shared__ unsigned long long sm[2][VIS_THREADS / 32];
  if (step >= 0) {
    if (threadIdx.x == 0) qs = replay_bisection(work, step, frac);
    __syncthreads();
  }
  const QuantileKeys q = step >= 0 ? qs : QuantileKeys{0xFFFFFFFFu, 0u, 0xFFFFFFFFu, 0u};
  unsigned long long s_lo = 0ull, s_hi = 0ull;
  uint32_t kmin = 0xFFFFFFFFu, kmax = 0u;
  bool any_max = false;
  auto visit = [&](float z, float a) {
    const uint32_t k = depth_key(z);
    if (step < 32) {
      const unsigned long long w = depth_weight(z, a);
      if (k <= q.c_lo) s_lo += w;
      if (k <= q.c_hi) s_hi += w;
    } else {
      if (k >= q.k_lo && k < kmin) kmin = k;
      if (k <= q.k_hi && k >= kmax) { kmax = k; any_max = true; }
    }
  };
  const long long tid = (long long)blockIdx.x * VIS_THREADS + threadIdx.x, nthr = (long long)gridDim.x * VIS_THREADS;
  long long done = 0;
  if (((reinterpret_cast<uintptr_t>(depth) | reinterpret_cast<uintptr_t>(acc)) & 15) == 0) {
    const long long n4 = n / 4;  // 16-byte loads, two groups in flight per thread
    const float4* d4 = reinterpret_cast<const float4*>(depth);
    const float4* a4 = reinterpret_cast<const float4*>(acc);
    const float4 ones = make_float4(1.f, 1.f, 1.f, 1.f);
    long long g = tid;
    for (; g + nthr < n4; g += 2 * nthr) {
      const float4 z0 = d4[g], z1 = d4[g + nthr];
      const float4 w0 = acc ? a4[g] : ones, w1 = acc ? a4[g + nthr] : ones;
      visit(z0.x, w0.x); visit(z0.y, w0.y); visit(z0.z, w0.z); visit(z0.w, w0.w);
      visit(z1.x, w1.x); visit(z1.y, w1.y); visit(z1.z, w1.z); visit(z1.w, w1.w);
    }
    for (; g < n4; g += nthr) {
      const float4 z0 = d4[g];
      const float4 w0 = acc ? a4[g] : ones;
      visit(z0.x, w0.x); visit(z0.y, w0.y); visit(z0.z, w0.z); visit(z0.w, w0.w);
    }
    done = n4 * 4;
  }
  for (long long i = done + tid; i < n; i += nthr) visit(depth[i], acc ? acc[i] : 1.f);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (step < 32) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s_lo += __shfl_xor_sync(FULL_MASK, s_lo, o);
      s_hi += __shfl_xor_sync(FULL_MASK, s_hi, o);
    }
    if (lane == 0) { sm[0][w] = s_lo; sm[1][w] = s_hi; }
    __syncthreads();
    if (threadIdx.x < 2) {
      unsigned long long s = 0ull;
      for (int i = 0; i < VIS_THREADS / 32; ++i) s += sm[threadIdx.x][i];
      // integer atomics: exact and order independent
      if (step < 0) { if (threadIdx.x == 0) atomicAdd(work + QW_TOTAL, s); }
      else atomicAdd(work + QW_STEP + 2 * step + threadIdx.x, s);
    }
  } else {
    // keys are shifted by one so that "no key <= k_hi" (0) differs from key 0
    unsigned long long vmax = any_max ? (unsigned long long)kmax + 1ull : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      kmin = min(kmin, __shfl_xor_sync(FULL_MASK, kmin, o));
      vmax = max(vmax, __shfl_xor_sync(FULL_MASK, vmax, o));
    }
    if (lane == 0) {
      atomicMin(work + QW_KMIN, (unsigned long long)kmin);
      atomicMax(work + QW_KMAX, vmax);
    }
  }
}

// pose.py:191-194: near = near or depth_keep[0] - eps; far = far or depth_keep[-1] + eps (float32 arithmetic)
__global__ void depth_range_finalize_kernel(const unsigned long long* __restrict__ work, float near_in, float far_in,
                                            int auto_near, int auto_far, float* __restrict__ range) {
  const float nan = __uint_as_float(0x7FC00000u);
  float near = near_in, far = far_in;
  if (auto_near) {
    const unsigned long long k = work[QW_KMIN];
    near = (k >= 0xFFFFFFFFull ? nan : key_depth((uint32_t)k)) - FLT_EPSILON;
  }
  if (auto_far) {
    const unsigned long long k = work[QW_KMAX];
    far = (k == 0ull ? nan : key_depth((uint32_t)(k - 1ull))) + FLT_EPSILON;
  }
  range[0] = near;
  range[1] = far;
}

__device__ __forceinline__ float curve_f(float x, int curve) {
  switch (curve) {
    case CURVE_NEG_LOG: return -logf(x + FLT_EPSILON);  // pose.py:153 (the default)
    case CURVE_INV: return 1.f / (x + FLT_EPSILON);
    case CURVE_LOG: return logf(x + FLT_EPSILON);
    default: return x;
  }
}
__device__ __forceinline__ float sinebow_f(float x) {  // pose.py:123-126: sin(pi x)^2
  const float s = sinf(3.14159274101257324f * x);
  return s * s;
}

// pose.py:196-212
__global__ void __launch_bounds__(VIS_THREADS) depth_vis_kernel(const float* __restrict__ depth, const float* __restrict__ acc,
                                                                 const float* __restrict__ range, int curve, float modulus,
                                                                 const float* __restrict__ lut, int n_lut, long long n,
                                                                 float* __restrict__ vis, uint8_t* __restrict__ vis8) {
  const long long i0 = ((long long)blockIdx.x * VIS_THREADS + threadIdx.x) * VIS_PPT;
  if (i0 >= n) return;
  const float cn = curve_f(range[0], curve), cf = curve_f(range[1], curve);
  const float lo = (isnan(cn) || isnan(cf)) ? __uint_as_float(0x7FC00000u) : fminf(cn, cf);  // np.minimum propagates NaN
  const float span = fabsf(cf - cn);
  float zs[VIS_PPT], as[VIS_PPT];
  if (i0 + VIS_PPT <= n && ((reinterpret_cast<uintptr_t>(depth) | reinterpret_cast<uintptr_t>(acc)) & 15) == 0) {
    const float4 z4 = *reinterpret_cast<const float4*>(depth + i0);
    zs[0] = z4.x; zs[1] = z4.y; zs[2] = z4.z; zs[3] = z4.w;
    if (acc) {
      const float4 a4 = *reinterpret_cast<const float4*>(acc + i0);
      as[0] = a4.x; as[1] = a4.y; as[2] = a4.z; as[3] = a4.w;
    }
  } else {
#pragma unroll
    for (int p = 0; p < VIS_PPT; ++p) {
      const long long i = i0 + p < n ? i0 + p : n - 1;
      zs[p] = depth[i];
      if (acc) as[p] = acc[i];
    }
  }
  float vf[VIS_PPT][3];
  uint8_t vb[VIS_PPT][3];
#pragma unroll
  for (int p = 0; p < VIS_PPT; ++p) {
    const float z = zs[p];
    float a = acc ? as[p] : 1.f;
    if (isnan(z)) a = 0.f;  // pose.py:178
    const float d = curve_f(z, curve);
    float value;
    if (modulus > 0.f) {
      float r = fmodf(d, modulus);  // np.mod: result takes the sign of the divisor
      if (r != 0.f && r < 0.f) r += modulus;
      value = r / modulus;
    } else {
      value = (d - lo) / span;
      value = isnan(value) ? 0.f : fminf(fmaxf(value, 0.f), 1.f);  // nan_to_num(clip(., 0, 1))
    }
    float rgb[3];
    if (lut) {
      // a listed colour map called with floats: index = trunc(value * N), value == 1 -> N - 1
      int idx = isnan(value) ? 0 : (int)(value * (float)n_lut);
      idx = idx < 0 ? 0 : (idx >= n_lut ? n_lut - 1 : idx);
      rgb[0] = __ldg(lut + 3 * idx); rgb[1] = __ldg(lut + 3 * idx + 1); rgb[2] = __ldg(lut + 3 * idx + 2);
    } else {
      rgb[0] = sinebow_f(0.5f - value);
      rgb[1] = sinebow_f(0.833333313465118408f - value);
      rgb[2] = sinebow_f(1.16666662693023682f - value);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      vf[p][c] = rgb[c] * a + (1.f - a);
      vb[p][c] = to8b_f(vf[p][c]);
    }
  }
  store_rgb4(vis, vis8, i0, n, vf, vb);
}

}  // namespace mip360

using namespace mip360;

extern "C" {

int mip360_vis_partials_len(void) { return 4 * VIS_MAX_BLOCKS; }
int mip360_vis_work_len(void) { return QW_LEN; }

int mip360_normals_scaling(const float* depth, int H, int W, double* partials, double* stats, mip360_stream_t stream) {
  MIP_REQUIRE(depth && partials && stats && H > 0 && W > 0, "normals_scaling: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = vis_grid((long long)H * W);
  depth_moments_kernel<0><<<grid, VIS_THREADS, 0, s>>>(depth, H, W, stats, partials);
  MIP_LAUNCH_CHECK();
  depth_moments_reduce_kernel<0><<<1, 32, 0, s>>>(partials, grid, stats);
  MIP_LAUNCH_CHECK();
  depth_moments_kernel<1><<<grid, VIS_THREADS, 0, s>>>(depth, H, W, stats, partials);
  MIP_LAUNCH_CHECK();
  depth_moments_reduce_kernel<1><<<1, 32, 0, s>>>(partials, grid, stats);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_visualize_normals(const float* depth, const float* acc, const double* stats, int H, int W, float* vis,
                             uint8_t* vis8, mip360_stream_t stream) {
  MIP_REQUIRE(H >= 0 && W >= 0, "visualize_normals: bad shape");
  const long long n = (long long)H * W;
  if (n == 0) return MIP360_OK;
  MIP_REQUIRE(depth && stats && (vis || vis8), "visualize_normals: null pointer");
  MIP_REQUIRE((reinterpret_cast<uintptr_t>(vis) & 15) == 0 && (reinterpret_cast<uintptr_t>(vis8) & 3) == 0, "visualize_normals: vis must be 16-byte and vis8 4-byte aligned");
  normals_kernel<<<(int)((n + VIS_THREADS * VIS_PPT - 1) / (VIS_THREADS * VIS_PPT)), VIS_THREADS, 0, (cudaStream_t)stream>>>(depth, acc, stats, H, W, vis, vis8);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_depth_range(const float* depth, const float* acc, long long n, double ignore_frac, float near, float far,
                       int auto_near, int auto_far, unsigned long long* work, float* range, mip360_stream_t stream) {
  MIP_REQUIRE(range && n >= 0, "depth_range: bad arguments");
  MIP_REQUIRE(ignore_frac >= 0.0 && ignore_frac <= 0.5, "depth_range: ignore_frac=%g outside [0, 0.5]", ignore_frac);
  cudaStream_t s = (cudaStream_t)stream;
  if (auto_near || auto_far) {
    MIP_REQUIRE(depth && work && n > 0, "depth_range: automatic near/far needs the frame and a work buffer");
    MIP_CUDA(cudaMemsetAsync(work, 0, sizeof(unsigned long long) * QW_LEN, s));
    const unsigned long long init_min = 0xFFFFFFFFull;
    MIP_CUDA(cudaMemcpyAsync(work + QW_KMIN, &init_min, sizeof(init_min), cudaMemcpyHostToDevice, s));
    const int grid = vis_grid(n);
    // ignore_frac == 0 keeps every pixel: the bisection passes are skipped, and replaying the (all-zero) record
    // yields k_lo = 0 and k_hi = 0xFFFFFFFF, so the last pass returns the smallest / largest key of the frame
    for (int step = ignore_frac > 0.0 ? -1 : 32; step <= 32; ++step) {
      depth_quantile_kernel<<<grid, VIS_THREADS, 0, s>>>(depth, acc, n, ignore_frac, step, work);
      MIP_LAUNCH_CHECK();
    }
  }
  depth_range_finalize_kernel<<<1, 1, 0, s>>>(work, near, far, auto_near, auto_far, range);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_visualize_depth(const float* depth, const float* acc, const float* range, int curve, float modulus,
                           const float* lut, int n_lut, long long n, float* vis, uint8_t* vis8, mip360_stream_t stream) {
  MIP_REQUIRE(n >= 0, "visualize_depth: bad size");
  if (n == 0) return MIP360_OK;
  MIP_REQUIRE(depth && range && (vis || vis8), "visualize_depth: null pointer");
  MIP_REQUIRE((reinterpret_cast<uintptr_t>(vis) & 15) == 0 && (reinterpret_cast<uintptr_t>(vis8) & 3) == 0, "visualize_depth: vis must be 16-byte and vis8 4-byte aligned");
  MIP_REQUIRE(curve >= 0 && curve <= 3, "visualize_depth: curve=%d", curve);
  MIP_REQUIRE(!lut || n_lut >= 1, "visualize_depth: empty colour table");
  depth_vis_kernel<<<(int)((n + VIS_THREADS * VIS_PPT - 1) / (VIS_THREADS * VIS_PPT)), VIS_THREADS, 0, (cudaStream_t)stream>>>(
      depth, acc, range, curve, modulus, lut, n_lut, n, vis, vis8);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

}  // extern "C"
