// K4: hierarchical resampling (intern/ray.py:12-57, :118-153).  One warp per ray; the ray's weights,
// CDF and bin edges are staged in shared memory, the CDF is a lane-chunked warp scan and the inverse
// CDF a binary search per output sample.  Compiled with --fmad=false so that the lerp and the CDF
// arithmetic round like the reference's unfused ops: given the same fp32 CDF and uniforms, the bin
// indices and samples are bit-identical to the reference's mask/max/min formulation (App. B2).
#include "common.cuh"
#include "ray_group.cuh"

namespace mip360 {

constexpr int RS_WARPS = 4;
constexpr int RS_MAXK = MIP360_MAX_SAMPLES + 1;  // knots

struct __align__(16) ResampleSmem {
  float wpad[MIP360_MAX_SAMPLES + 2];
  float w[MIP360_MAX_SAMPLES];
  float cdf[RS_MAXK];
  float bins[RS_MAXK];
};

// ray.py:137-142: w_pad = [w0, w, w_{N-1}]; w_max = max(adjacent) [N+1]; blur = .5*(adjacent) [N]; + padding
__device__ __forceinline__ void warp_blur(const float* wpad, float* wout, int N, float padding, int lane) {
  for (int j = lane; j < N; j += 32) {
    const float m0 = fmaxf(wpad[j], wpad[j + 1]);
    const float m1 = fmaxf(wpad[j + 1], wpad[j + 2]);
    wout[j] = 0.5f * (m0 + m1) + padding;
  }
}

// ray.py:15-27 on w[N] (shared memory) -> cdf[N+1].  MAXC = intervals per lane (4: N <= 128, 16: N <= 512)
constexpr int RS_MAXC_SMALL = 4, RS_MAXC_LARGE = (MIP360_MAX_SAMPLES + 31) / 32;
#define RS_GENERIC(kernel, N) ((N) <= 32 * RS_MAXC_SMALL ? kernel<RS_MAXC_SMALL> : kernel<RS_MAXC_LARGE>)
template <int MAXC>
__device__ __forceinline__ void warp_cdf(const float* w, float* cdf, int N, int lane) {
  const int C = (N + 31) >> 5;  // contiguous chunk per lane
  const int j0 = lane * C;
  float loc = 0.f;
  for (int c = 0; c < C; ++c) {
    const int j = j0 + c;
    if (j < N) loc += w[j];
  }
  float wsum = warp_sum(loc);
  const float eps = 1e-5f;
  const float padding = fmaxf(0.f, eps - wsum);
  const float add = padding / (float)N;
  wsum = wsum + padding;
  // pdf and its inclusive scan
  float run = 0.f;
  float pdf_loc[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int j = j0 + c;
    float p = 0.f;
    if (c < C && j < N) p = (w[j] + add) / wsum;
    run += p;
    pdf_loc[c] = run;  // inclusive within the chunk
  }
  const float incl = warp_scan_incl(run, lane);
  const float off = incl - run;
  // The tree scan can leave a lane's offset one ulp below the last value of the lane before it.  A CDF must be
  // non-decreasing (the search relies on it), so every value is raised to the running maximum of the preceding
  // lanes' last values; max is exact and associative, so this is a plain shuffle scan.
  float pm = off + run;  // this lane's last value
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(FULL_MASK, pm, o);
    if (lane >= o) pm = fmaxf(pm, n);
  }
  pm = __shfl_up_sync(FULL_MASK, pm, 1);
  if (lane == 0) pm = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const int j = j0 + c;
    if (c < C && j < N - 1) cdf[j + 1] = fminf(1.f, fmaxf(off + pdf_loc[c], pm));
  }
  if (lane == 0) {
    cdf[0] = 0.f;
    cdf[N] = 1.f;
  }
}

// |d t_mean|^2 of one interval, formed like frustum_norm_sq_rg_kernel (parameterization.py:103, App. A1)
__device__ __forceinline__ double interval_norm_sq(float t0, float t1, float d0, float d1, float d2) {
  const float mu = (t0 + t1) / 2.f, hw = (t1 - t0) / 2.f;
  const float hw2 = hw * hw;
  const float t_mean = mu + (2.f * mu * hw2) / (3.f * (mu * mu) + hw2);
  const float m0 = d0 * t_mean, m1 = d1 * t_mean, m2 = d2 * t_mean;
  return (double)(m0 * m0 + m1 * m1 + m2 * m2);
}
// block sum of one double per thread, one atomicAdd per block
template <int THREADS>
__device__ __forceinline__ void block_atomic_add(double v, double* out) {
  __shared__ double sm_part[THREADS / 32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm_part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) s += sm_part[i];
    atomicAdd(out, s);
  }
}

// ray.py:41-56: last knot with cdf <= u (upper_bound - 1), lerp inside the interval
__device__ __forceinline__ float invert_one(const float* cdf, const float* bins, int N, float u, int* idx_out) {
  int lo = 0, hi = N + 1;  // first k in [0, N+1) with cdf[k] > u
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  int i0 = lo - 1;
  i0 = i0 < 0 ? 0 : i0;  // u < cdf[0] cannot happen for u >= 0; the reference then falls back to knot 0
  const int i1 = (lo > N) ? N : lo;
  const float c0 = cdf[i0], c1 = cdf[i1];
  const float b0 = bins[i0], b1 = bins[i1];
  float t = nan_to_num_f((u - c0) / (c1 - c0), 0.f);
  t = fminf(fmaxf(t, 0.f), 1.f);
  if (idx_out) *idx_out = i0;
  return b0 + t * (b1 - b0);
}

template <int MAXC>
__global__ void __launch_bounds__(RS_WARPS * 32)
resample_kernel(const float* __restrict__ t_vals, const float* __restrict__ weights, const float* __restrict__ u_base,
                const float* __restrict__ jitter, RngArgs rng, float jitter_scale, const float* __restrict__ directions,
                double* __restrict__ norm_sq, int B, int N, float padding, int blur, float* __restrict__ new_t) {
  __shared__ ResampleSmem sm[RS_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  ResampleSmem& s = sm[warp];
  const int K = N + 1;
  const float one_m_eps = 1.f - 1.1920928955078125e-07f;
  const uint32_t epoch = rng.enabled ? rng_epoch(rng) : 0u;
  double norm_acc = 0.0;
  for (int b = blockIdx.x * RS_WARPS + warp; b < B; b += gridDim.x * RS_WARPS) {
    const float* wrow = weights + (long long)b * N;
    const float* trow = t_vals + (long long)b * K;
    for (int j = lane; j < N; j += 32) s.wpad[j + 1] = wrow[j];
    for (int k = lane; k < K; k += 32) s.bins[k] = trow[k];
    if (lane == 0) {
      s.wpad[0] = wrow[0];
      s.wpad[N + 1] = wrow[N - 1];
    }
    __syncwarp();
    if (blur) {
      warp_blur(s.wpad, s.w, N, padding, lane);
    } else {
      for (int j = lane; j < N; j += 32) s.w[j] = s.wpad[j + 1];
    }
    __syncwarp();
    warp_cdf<MAXC>(s.w, s.cdf, N, lane);
    __syncwarp();
    for (int m = lane; m < K; m += 32) {
      float u = u_base[m];
      if (jitter || rng.enabled) {
        // ray.py:33: uniform_(0, 1/M - eps) — a given draw, or the in-kernel generator scaled the way torch scales it
        const float jit = jitter ? jitter[(long long)b * K + m]
                                 : rng_uniform(rng, epoch, (uint32_t)b, (uint32_t)m) * jitter_scale;
        u = (u + u) + jit;  // the doubled stratum offset is the reference's (App. A5)
        u = fminf(u, one_m_eps);
      }
      const float x = invert_one(s.cdf, s.bins, N, u, nullptr);
      new_t[(long long)b * K + m] = x;
      if (norm_sq) s.wpad[m] = x;  // the padded weights are dead by now (N + 2 >= K slots)
    }
    __syncwarp();
    if (norm_sq) {
      const float d0 = directions[b * 3], d1 = directions[b * 3 + 1], d2 = directions[b * 3 + 2];
      for (int j = lane; j < N; j += 32) norm_acc += interval_norm_sq(s.wpad[j], s.wpad[j + 1], d0, d1, d2);
      __syncwarp();
    }
  }
  if (norm_sq) block_atomic_add<RS_WARPS * 32>(norm_acc, norm_sq);
}

__global__ void __launch_bounds__(RS_WARPS * 32)
blur_kernel(const float* __restrict__ weights, int B, int N, float padding, float* __restrict__ out) {
  __shared__ ResampleSmem sm[RS_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  ResampleSmem& s = sm[warp];
  for (int b = blockIdx.x * RS_WARPS + warp; b < B; b += gridDim.x * RS_WARPS) {
    const float* wrow = weights + (long long)b * N;
    for (int j = lane; j < N; j += 32) s.wpad[j + 1] = wrow[j];
    if (lane == 0) {
      s.wpad[0] = wrow[0];
      s.wpad[N + 1] = wrow[N - 1];
    }
    __syncwarp();
    warp_blur(s.wpad, s.w, N, padding, lane);
    __syncwarp();
    for (int j = lane; j < N; j += 32) out[(long long)b * N + j] = s.w[j];
    __syncwarp();
  }
}

template <int MAXC>
__global__ void __launch_bounds__(RS_WARPS * 32)
cdf_kernel(const float* __restrict__ weights, int B, int N, float* __restrict__ cdf) {
  __shared__ ResampleSmem sm[RS_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  ResampleSmem& s = sm[warp];
  for (int b = blockIdx.x * RS_WARPS + warp; b < B; b += gridDim.x * RS_WARPS) {
    for (int j = lane; j < N; j += 32) s.w[j] = weights[(long long)b * N + j];
    __syncwarp();
    warp_cdf<MAXC>(s.w, s.cdf, N, lane);
    __syncwarp();
    for (int k = lane; k <= N; k += 32) cdf[(long long)b * (N + 1) + k] = s.cdf[k];
    __syncwarp();
  }
}

__global__ void __launch_bounds__(RS_WARPS * 32)
invert_kernel(const float* __restrict__ bins, const float* __restrict__ cdf, const float* __restrict__ u,
              int u_row_stride, int B, int N, int M, float* __restrict__ samples, int32_t* __restrict__ idx) {
  __shared__ ResampleSmem sm[RS_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  ResampleSmem& s = sm[warp];
  const int K = N + 1;
  for (int b = blockIdx.x * RS_WARPS + warp; b < B; b += gridDim.x * RS_WARPS) {
    for (int k = lane; k < K; k += 32) {
      s.cdf[k] = cdf[(long long)b * K + k];
      s.bins[k] = bins[(long long)b * K + k];
    }
    __syncwarp();
    for (int m = lane; m < M; m += 32) {
      int i0;
      const float x = invert_one(s.cdf, s.bins, N, u[(long long)b * u_row_stride + m], &i0);
      samples[(long long)b * M + m] = x;
      if (idx) idx[(long long)b * M + m] = i0;
    }
    __syncwarp();
  }
}

// N in {32, 64, 128}: 8 lanes per ray (ray_group.cuh).  Weights live in registers (blocked: E contiguous per
// lane), the blur needs one neighbour value from each adjacent lane, the CDF is a lane-local loop + 3-step group
// scan.  Knots are read with coalesced cyclic loads straight into shared memory; CDF and bins sit there in a
// skewed layout (index + index/8: blocked stores and cyclic stores are both bank-conflict free) for the searches,
// and the 8 lanes write 8 consecutive samples per step.
// skewed shared-memory index: one pad word per lane chunk of E knots (E = 4, 8, 16)
template <int E>
__device__ __forceinline__ int rg_skew(int i) {
  return i + (i >> (E == 4 ? 2 : E == 8 ? 3 : 4));
}

template <int E>
__global__ void __launch_bounds__(RG_THREADS)
resample_rg_kernel(const float* __restrict__ t_vals, const float* __restrict__ weights, const float* __restrict__ u_base,
                   const float* __restrict__ jitter, RngArgs rng, float jitter_scale,
                   const float* __restrict__ directions, double* __restrict__ norm_sq, int B, float padding, int blur,
                   float* __restrict__ new_t) {
  constexpr int N = E * RG_LANES, K = N + 1, ROW = K + K / E + 2;
  __shared__ float s_cdf[RG_RAYS_PER_BLOCK][ROW];
  __shared__ float s_bins[RG_RAYS_PER_BLOCK][ROW];
  const int gl = threadIdx.x & 7, g = threadIdx.x >> 3, j0 = gl * E;
  const float one_m_eps = 1.f - 1.1920928955078125e-07f;
  float* cdf = s_cdf[g];
  float* bins = s_bins[g];
  const bool randomized = jitter != nullptr || rng.enabled;
  const uint32_t epoch = rng.enabled ? rng_epoch(rng) : 0u;
  double norm_acc = 0.0;
  for (long long base = (long long)blockIdx.x * RG_RAYS_PER_BLOCK; base < B; base += (long long)gridDim.x * RG_RAYS_PER_BLOCK) {
    const long long ray_raw = base + g;
    const bool active = ray_raw < B;
    const long long ray = active ? ray_raw : B - 1;
    float w[E];
    rg_load<E>(weights + ray * N + j0, w);
    const float* trow = t_vals + ray * K;
#pragma unroll
    for (int c = 0; c < E; ++c) bins[rg_skew<E>(gl + RG_LANES * c)] = __ldg(trow + gl + RG_LANES * c);
    if (gl == 0) bins[rg_skew<E>(N)] = __ldg(trow + N);
    // the jitter row is fetched now, with the weights and knots, not after the CDF has been built (a load issued
    // behind the scan would cost a second round trip to memory)
    float jit[E + 1];
    if (jitter) {
#pragma unroll
      for (int c = 0; c <= E; ++c) jit[c] = __ldg(jitter + ray * K + ((c < E) ? gl + RG_LANES * c : N));
    } else if (rng.enabled) {
      // ray.py:33 drawn here instead of being read: uniform_(0, 1/M - eps) = u01 * (1/M - eps)
      rg_draw<E>(rng, epoch, (uint32_t)ray, gl, jit);
#pragma unroll
      for (int c = 0; c <= E; ++c) jit[c] = jit[c] * jitter_scale;
    }
    float d0 = 0.f, d1 = 0.f, d2 = 0.f;
    if (norm_sq) { d0 = __ldg(directions + ray * 3); d1 = __ldg(directions + ray * 3 + 1); d2 = __ldg(directions + ray * 3 + 2); }
    if (blur) {
      float wl = __shfl_up_sync(FULL_MASK, w[E - 1], 1, RG_LANES);
      float wr = __shfl_down_sync(FULL_MASK, w[0], 1, RG_LANES);
      if (gl == 0) wl = w[0];
      if (gl == RG_LANES - 1) wr = w[E - 1];
      float mx[E + 1];  // mx[i] = max(w[j0+i-1], w[j0+i]) with replicated ends
      mx[0] = fmaxf(wl, w[0]);
#pragma unroll
      for (int i = 1; i < E; ++i) mx[i] = fmaxf(w[i - 1], w[i]);
      mx[E] = fmaxf(w[E - 1], wr);
#pragma unroll
      for (int i = 0; i < E; ++i) w[i] = 0.5f * (mx[i] + mx[i + 1]) + padding;
    }
    float loc = 0.f;
#pragma unroll
    for (int i = 0; i < E; ++i) loc += w[i];
    float wsum = rg_sum(loc);
    const float pad = fmaxf(0.f, 1e-5f - wsum);
    const float add = pad / (float)N;
    wsum = wsum + pad;
    // pdf = (w + add) / wsum as a multiplication by the (correctly rounded) reciprocal: the CDF only has to agree
    // with the reference to rounding (its scan order differs anyway); indices are exact GIVEN a CDF
    const float inv_wsum = 1.f / wsum;
    float run = 0.f, incl[E];
#pragma unroll
    for (int i = 0; i < E; ++i) {
      run += (w[i] + add) * inv_wsum;
      incl[i] = run;
    }
    const float off = rg_scan_excl(run, gl);
    // keep the CDF non-decreasing across lane boundaries (see warp_cdf): exact running max of the lanes' last values
    float pm = off + run;
#pragma unroll
    for (int o = 1; o < RG_LANES; o <<= 1) {
      const float n = __shfl_up_sync(FULL_MASK, pm, o, RG_LANES);
      if (gl >= o) pm = fmaxf(pm, n);
    }
    pm = __shfl_up_sync(FULL_MASK, pm, 1, RG_LANES);
    if (gl == 0) pm = 0.f;
    float my_last = 0.f;  // cdf[(gl+1)*E]: the last knot of this lane's chunk
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const float cv = (j0 + i == N - 1) ? 1.f : fminf(1.f, fmaxf(off + incl[i], pm));
      cdf[rg_skew<E>(j0 + i + 1)] = cv;
      my_last = cv;
    }
    if (gl == 0) cdf[0] = 0.f;
    // two-level search, level 1 in registers: every lane of the group gets the 8 chunk-end knots
    float chunk_end[RG_LANES];
#pragma unroll
    for (int l = 0; l < RG_LANES; ++l) chunk_end[l] = __shfl_sync(FULL_MASK, my_last, l, RG_LANES);
    __syncwarp();
    // cnt = number of knots <= u (upper bound).  Level 1: whole chunks below u (sorted, so a sum of predicates);
    // level 2: branch-free counting search inside the one partial chunk (at most E-1 of its knots are <= u), the
    // lane's E (+1 for lane 0: sample N) searches advancing in lock step.
    constexpr int S = E + 1;
    float u[S];
    int cbase[S], cnt[S];
#pragma unroll
    for (int c = 0; c < S; ++c) {
      const int m = (c < E) ? gl + RG_LANES * c : N;
      u[c] = __ldg(u_base + m);
      if (randomized) {
        u[c] = (u[c] + u[c]) + jit[c];  // the doubled stratum offset is the reference's (App. A5)
        u[c] = fminf(u[c], one_m_eps);
      }
      // number of chunk ends <= u (sorted; the last one, cdf[N] = 1 > u, never counts): a 3-step bisection over the
      // 8 registers with selects instead of 8 compare-and-add pairs
      const bool p1 = chunk_end[3] <= u[c];
      const float e2 = p1 ? chunk_end[5] : chunk_end[1];
      const bool p2 = e2 <= u[c];
      const float e3 = p1 ? (p2 ? chunk_end[6] : chunk_end[4]) : (p2 ? chunk_end[2] : chunk_end[0]);
      const int nfull = (p1 ? 4 : 0) + (p2 ? 2 : 0) + (e3 <= u[c] ? 1 : 0);
      // the partial chunk holds knots E*n+1 .. E*n+E, skewed position (E+1)*n + 1 + offset (contiguous up to E-1)
      cbase[c] = min(nfull, RG_LANES - 1) * (E + 1);
      cnt[c] = 0;  // nfull == 8 (u >= cdf[N] = 1) cannot happen: u <= 1 - eps
    }
#pragma unroll
    for (int step = E / 2; step > 0; step >>= 1) {
#pragma unroll
      for (int c = 0; c < S; ++c) {
        const float v = cdf[cbase[c] + cnt[c] + step];
        if (v <= u[c]) cnt[c] += step;
      }
    }
#pragma unroll
    for (int c = 0; c < S; ++c) {
      // knots <= u: knot 0, the full chunks and cnt of the partial chunk  =>  i0 = E*n + cnt, i1 = i0 + 1
      const int p0 = cbase[c] + cnt[c], p1 = p0 + 1 + (cnt[c] == E - 1 ? 1 : 0);
      const float c0 = cdf[p0], c1 = cdf[p1];
      const float b0 = bins[p0], b1 = bins[p1];
      // clip(nan_to_num((u-c0)/(c1-c0), 0), 0, 1) without the special-value tests: a zero-width CDF step gives
      // +inf -> 1 when u > c0 and nan -> 0 when u == c0 (c0 <= u always holds for the selected knot)
      const float den = c1 - c0, num = u[c] - c0;
      const float tt = den > 0.f ? fminf(fmaxf(__fdividef(num, den), 0.f), 1.f) : (num > 0.f ? 1.f : 0.f);
      const int m = (c < E) ? gl + RG_LANES * c : N;
      const float x = b0 + tt * (b1 - b0);
      if (active && (c < E || gl == 0)) new_t[ray * K + m] = x;
      u[c] = x;  // kept for the norm below
    }
    __syncwarp();
    if (norm_sq) {
      // park the new knots in the (now dead) CDF row so that every lane can read interval ends t[m], t[m+1]
#pragma unroll
      for (int c = 0; c < S; ++c)
        if (c < E || gl == 0) cdf[rg_skew<E>((c < E) ? gl + RG_LANES * c : N)] = u[c];
      __syncwarp();
      if (active) {
#pragma unroll
        for (int c = 0; c < E; ++c) {
          const int m = gl + RG_LANES * c;
          norm_acc += interval_norm_sq(cdf[rg_skew<E>(m)], cdf[rg_skew<E>(m + 1)], d0, d1, d2);
        }
      }
      __syncwarp();
    }
  }
  if (norm_sq) block_atomic_add<RG_THREADS>(norm_acc, norm_sq);
}

static inline int ray_grid(int B, int warps) {
  long long b = ((long long)B + warps - 1) / warps;
  const long long cap = (long long)sm_count() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace mip360

using namespace mip360;

extern "C" {

int mip360_blur_weights(const float* weights, int B, int N, float resample_padding, float* out,
                        mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (weights && out), "blur_weights: null pointer");
  MIP_REQUIRE(N >= 1 && N <= MIP360_MAX_SAMPLES, "blur_weights: N=%d outside [1,%d]", N, MIP360_MAX_SAMPLES);
  if (B <= 0) return MIP360_OK;
  blur_kernel<<<ray_grid(B, RS_WARPS), RS_WARPS * 32, 0, (cudaStream_t)stream>>>(weights, B, N, resample_padding, out);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_resample_cdf(const float* weights, int B, int N, float* cdf, mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (weights && cdf), "resample_cdf: null pointer");
  MIP_REQUIRE(N >= 1 && N <= MIP360_MAX_SAMPLES, "resample_cdf: N=%d outside [1,%d]", N, MIP360_MAX_SAMPLES);
  if (B <= 0) return MIP360_OK;
  RS_GENERIC(cdf_kernel, N)<<<ray_grid(B, RS_WARPS), RS_WARPS * 32, 0, (cudaStream_t)stream>>>(weights, B, N, cdf);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_resample_invert(const float* bins, const float* cdf, const float* u, int u_row_stride, int B, int N, int M,
                           float* samples, int32_t* idx, mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (bins && cdf && u && samples), "resample_invert: null pointer");
  MIP_REQUIRE(N >= 1 && N <= MIP360_MAX_SAMPLES, "resample_invert: N=%d outside [1,%d]", N, MIP360_MAX_SAMPLES);
  MIP_REQUIRE(M >= 1, "resample_invert: M=%d", M);
  if (B <= 0) return MIP360_OK;
  invert_kernel<<<ray_grid(B, RS_WARPS), RS_WARPS * 32, 0, (cudaStream_t)stream>>>(bins, cdf, u, u_row_stride, B, N, M,
                                                                                   samples, idx);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_resample(const float* t_vals, const float* weights, const float* u_base, const float* jitter, int B, int N,
                    float resample_padding, int blur, float* new_t, mip360_stream_t stream) {
  return mip360_resample_sample(t_vals, weights, u_base, jitter, 0, 0ull, 0u, nullptr, 0.f, nullptr, nullptr, B, N,
                                resample_padding, blur, new_t, stream);
}

int mip360_resample_sample(const float* t_vals, const float* weights, const float* u_base, const float* jitter,
                           int use_rng, unsigned long long rng_seed, unsigned int rng_stream,
                           const unsigned long long* rng_epoch, float jitter_scale, const float* directions,
                           double* norm_sq, int B, int N, float resample_padding, int blur, float* new_t,
                           mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (t_vals && weights && u_base && new_t), "resample: null pointer");
  MIP_REQUIRE(N >= 1 && N <= MIP360_MAX_SAMPLES, "resample: N=%d outside [1,%d]", N, MIP360_MAX_SAMPLES);
  MIP_REQUIRE(!norm_sq || directions, "resample: the norm needs the ray directions");
  MIP_REQUIRE(!(jitter && use_rng), "resample: either a given draw or the in-kernel generator");
  if (B <= 0) return MIP360_OK;
  const bool rg = rg_supported_host(N);
  const RngArgs rng{rng_seed, rng_epoch, rng_stream, use_rng ? 1 : 0};
  cudaStream_t st = (cudaStream_t)stream;
  if (rg && N == 32)
    resample_rg_kernel<4><<<rg_grid(B), RG_THREADS, 0, st>>>(t_vals, weights, u_base, jitter, rng, jitter_scale, directions,
                                                            norm_sq, B, resample_padding, blur, new_t);
  else if (rg && N == 64)
    resample_rg_kernel<8><<<rg_grid(B), RG_THREADS, 0, st>>>(t_vals, weights, u_base, jitter, rng, jitter_scale, directions,
                                                            norm_sq, B, resample_padding, blur, new_t);
  else if (rg && N == 128)
    resample_rg_kernel<16><<<rg_grid(B), RG_THREADS, 0, st>>>(t_vals, weights, u_base, jitter, rng, jitter_scale,
                                                             directions, norm_sq, B, resample_padding, blur, new_t);
  else
    RS_GENERIC(resample_kernel, N)<<<ray_grid(B, RS_WARPS), RS_WARPS * 32, 0, st>>>(t_vals, weights, u_base, jitter, rng, jitter_scale,
                                                                     directions, norm_sq, B, N, resample_padding, blur,
                                                                     new_t);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

}  // extern "C"
