// K5 distortion regulariser (intern/regularization.py:3-19) and K6 interlevel loss
// (intern/distillation.py:4-51), forward and backward, one warp per ray, O(N) per ray instead of the
// reference's O(N^2) / O(N) Python loops.  Prefix sums that are later differenced are carried in fp64
// so that the cancellation does not cost fp32 accuracy.
#include "common.cuh"
#include "ray_group.cuh"

namespace mip360 {

constexpr int LS_WARPS = 4;
// intervals per lane of the generic kernels: 4 covers N <= 128, 16 the N <= 512 limit (template parameter LS_MAXC)
constexpr int LS_MAXC_SMALL = 4, LS_MAXC_LARGE = (MIP360_MAX_SAMPLES + 31) / 32;
#define LS_GENERIC(N, ...) ((N) <= 32 * LS_MAXC_SMALL ? __VA_ARGS__ LS_MAXC_SMALL> : __VA_ARGS__ LS_MAXC_LARGE>)
constexpr int LS_MAX_PARTIALS = 4096;

__device__ __forceinline__ double warp_scan_incl_d(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double n = __shfl_up_sync(FULL_MASK, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

struct __align__(16) DistSmem {
  float s[MIP360_MAX_SAMPLES + 1];
  float w[MIP360_MAX_SAMPLES];
};

// per-ray: 2 sum_i w_i (m_i W_<i - (wm)_<i) + 1/3 sum_i w_i^2 ds_i   (m sorted, App. A9)
template <bool BWD, int LS_MAXC>
__global__ void __launch_bounds__(LS_WARPS * 32)
distortion_kernel(const float* __restrict__ s_vals, const float* __restrict__ weights, int B, int N,
                  float* __restrict__ per_ray, double* __restrict__ partials, const float* __restrict__ g_loss_ptr,
                  float* __restrict__ g_w) {
  __shared__ DistSmem sm[LS_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  DistSmem& s = sm[warp];
  const int C = (N + 31) >> 5, j0 = lane * C;
  double block_acc = 0.0;
  float g_loss = 1.f;
  if (BWD) g_loss = *g_loss_ptr;
  for (int b = blockIdx.x * LS_WARPS + warp; b < B; b += gridDim.x * LS_WARPS) {
    for (int k = lane; k <= N; k += 32) s.s[k] = s_vals[(long long)b * (N + 1) + k];
    for (int j = lane; j < N; j += 32) s.w[j] = weights[(long long)b * N + j];
    __syncwarp();
    double eW[LS_MAXC], eWM[LS_MAXC];
    float m[LS_MAXC], w[LS_MAXC], ds[LS_MAXC];
    double runW = 0.0, runWM = 0.0;
#pragma unroll
    for (int c = 0; c < LS_MAXC; ++c) {
      const int j = j0 + c;
      m[c] = w[c] = ds[c] = 0.f;
      if (c < C && j < N) {
        m[c] = 0.5f * (s.s[j] + s.s[j + 1]);
        ds[c] = s.s[j + 1] - s.s[j];
        w[c] = s.w[j];
      }
      eW[c] = runW;
      eWM[c] = runWM;
      runW += (double)w[c];
      runWM += (double)w[c] * (double)m[c];
    }
    const double inclW = warp_scan_incl_d(runW, lane), inclWM = warp_scan_incl_d(runWM, lane);
    const double offW = inclW - runW, offWM = inclWM - runWM;
    if (!BWD) {
      double loss = 0.0;
#pragma unroll
      for (int c = 0; c < LS_MAXC; ++c) {
        const double Wl = offW + eW[c], WMl = offWM + eWM[c];
        loss += 2.0 * (double)w[c] * ((double)m[c] * Wl - WMl) + (double)w[c] * (double)w[c] * (double)ds[c] / 3.0;
      }
      loss = warp_sum(loss);
      if (lane == 0) {
        if (per_ray) per_ray[b] = (float)loss;
        block_acc += loss;
      }
    } else {
      const double Wtot = __shfl_sync(FULL_MASK, inclW, 31), WMtot = __shfl_sync(FULL_MASK, inclWM, 31);
#pragma unroll
      for (int c = 0; c < LS_MAXC; ++c) {
        const int j = j0 + c;
        if (c < C && j < N) {
          const double Wl = offW + eW[c], WMl = offWM + eWM[c];
          const double Wr = Wtot - Wl - (double)w[c], WMr = WMtot - WMl - (double)w[c] * (double)m[c];
          // d/dw_i = 2 sum_j w_j |m_i - m_j| + 2/3 w_i ds_i
          const double g = 2.0 * ((double)m[c] * (Wl - Wr) - (WMl - WMr)) + (2.0 / 3.0) * (double)w[c] * (double)ds[c];
          g_w[(long long)b * N + j] = (float)(g * (double)g_loss);
        }
      }
    }
    __syncwarp();
  }
  if (!BWD) {
    __shared__ double sm_part[LS_WARPS];
    if (lane == 0) sm_part[warp] = block_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < LS_WARPS; ++i) t += sm_part[i];
      partials[blockIdx.x] = t;
    }
  }
}

// final deterministic reduction of per-block partial sums
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ partials, int n, double scale,
                                                              float* __restrict__ out) {
  __shared__ double sm[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) a += partials[i];
  sm[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)(sm[0] * scale);
}

struct __align__(16) BoundsSmem {
  float tf[MIP360_MAX_SAMPLES + 1];
  float tc[MIP360_MAX_SAMPLES + 1];
  double cumw[MIP360_MAX_SAMPLES + 1];
};

// b[r,i] = sum_j w_j [t0_j <= R_i and t1_j >= L_i]  (closed intervals; distillation.py:25-29, App. B5)
template <int LS_MAXC>
__global__ void __launch_bounds__(LS_WARPS * 32)
bounds_kernel(const float* __restrict__ t_fine, const float* __restrict__ w_fine, const float* __restrict__ t_coarse,
              int B, int N, float* __restrict__ b_out, double* __restrict__ total) {
  __shared__ BoundsSmem sm[LS_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  BoundsSmem& s = sm[warp];
  const int C = (N + 31) >> 5, j0 = lane * C;
  double col[LS_MAXC];  // column sums over this warp's rays (columns lane, lane + 32, ...)
#pragma unroll
  for (int c = 0; c < LS_MAXC; ++c) col[c] = 0.0;
  for (int b = blockIdx.x * LS_WARPS + warp; b < B; b += gridDim.x * LS_WARPS) {
    for (int k = lane; k <= N; k += 32) {
      s.tf[k] = t_fine[(long long)b * (N + 1) + k];
      s.tc[k] = t_coarse[(long long)b * (N + 1) + k];
    }
    // exclusive prefix sums of the fine weights in fp64
    double run = 0.0, ex[LS_MAXC];
#pragma unroll
    for (int c = 0; c < LS_MAXC; ++c) {
      const int j = j0 + c;
      ex[c] = run;
      if (c < C && j < N) run += (double)w_fine[(long long)b * N + j];
    }
    const double incl = warp_scan_incl_d(run, lane);
    const double off = incl - run;
#pragma unroll
    for (int c = 0; c < LS_MAXC; ++c) {
      const int j = j0 + c;
      if (c < C && j < N) s.cumw[j] = off + ex[c];
    }
    if (lane == 31) s.cumw[N] = incl;
    __syncwarp();
#pragma unroll
    for (int ci = 0; ci < LS_MAXC; ++ci) {
      const int i = lane + 32 * ci;
      if (i >= N) break;
      const float L = s.tc[i], R = s.tc[i + 1];
      // lo = first j in [0,N) with t1_j = tf[j+1] >= L
      int lo = 0, hi = N;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s.tf[mid + 1] >= L) hi = mid; else lo = mid + 1;
      }
      const int first = lo;
      // last j with t0_j = tf[j] <= R  -> count of j in [0,N) with tf[j] <= R, minus 1
      lo = 0; hi = N;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s.tf[mid] <= R) lo = mid + 1; else hi = mid;
      }
      const int last = lo - 1;
      float v = 0.f;
      if (last >= first) v = fmaxf((float)(s.cumw[last + 1] - s.cumw[first]), 0.f);
      if (b_out) b_out[(long long)b * N + i] = v;
      col[ci] += (double)v;
    }
    __syncwarp();
  }
  if (total) {
    // distillation.py:25-29 sums the bound of interval i over ALL rays (App. A6): block-level column sums, one atomic
    // per column and block
    __syncthreads();
    double* red = reinterpret_cast<double*>(sm);  // LS_WARPS x N doubles fit in the staging area
#pragma unroll
    for (int ci = 0; ci < LS_MAXC; ++ci) {
      const int i = lane + 32 * ci;
      if (i < N) red[warp * N + i] = col[ci];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += LS_WARPS * 32) {
      double a = 0.0;
      for (int w = 0; w < LS_WARPS; ++w) a += red[w * N + i];
      atomicAdd(&total[i], a);
    }
  }
}

// column sums over rays, accumulated in fp64 into bound_total[N]
__global__ void __launch_bounds__(256)
bounds_reduce_kernel(const float* __restrict__ b, int B, int N, int rows_per_block, double* __restrict__ total) {
  // blockDim = (128, 2): x over columns, y over row parity
  __shared__ double sm[2][128];
  const int y = threadIdx.y;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(B, r0 + rows_per_block);
  for (int n0 = 0; n0 < N; n0 += 128) {  // 128 columns at a time
    const int n = n0 + threadIdx.x;
    double a = 0.0;
    if (n < N)
      for (int r = r0 + y; r < r1; r += 2) a += (double)b[(long long)r * N + n];
    sm[y][threadIdx.x] = a;
    __syncthreads();
    if (y == 0 && n < N) atomicAdd(&total[n], sm[0][threadIdx.x] + sm[1][threadIdx.x]);
    __syncthreads();
  }
}

// loss = sum relu(bnd - w)^2 / (w + 1e-6) / batch_div   (distillation.py:48-49)
template <bool BWD>
__global__ void __launch_bounds__(256)
interlevel_kernel(const float* __restrict__ w_hat, const float* __restrict__ b_per_ray,
                  const double* __restrict__ bound_total, long long total, int N, int bound_mode, float batch_div,
                  double* __restrict__ partials, const float* __restrict__ g_loss_ptr, float* __restrict__ g_w_hat) {
  double acc = 0.0;
  float g_scale = 0.f;
  if (BWD) g_scale = *g_loss_ptr / batch_div;
  if ((N & 3) == 0 && bound_mode == 0) {
    // four consecutive columns per thread (16-byte loads / stores), U quads in flight per thread, the batch bounds as
    // floats in shared memory, the quad's column carried along instead of a 64-bit modulo per quad.
    __shared__ float bnd_s[MIP360_MAX_SAMPLES];
    for (int i = threadIdx.x; i < N; i += blockDim.x) bnd_s[i] = (float)bound_total[i];
    __syncthreads();
    constexpr int U = 4;
    const long long quads = total >> 2, qstride = (long long)gridDim.x * blockDim.x;
    const int nq = N >> 2, qstep = (int)(qstride % nq);
    long long q0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int c0 = (int)(q0 % nq);
    for (; q0 < quads; q0 += U * qstride) {
      float4 w4[U];
      int col[U];
      int c = c0;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long q = q0 + u * qstride;
        col[u] = c;
        c = c + qstep >= nq ? c + qstep - nq : c + qstep;
        if (q < quads) w4[u] = __ldg(reinterpret_cast<const float4*>(w_hat) + q);
      }
      c0 = c;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long q = q0 + u * qstride;
        if (q >= quads) break;
        const float w[4] = {w4[u].x, w4[u].y, w4[u].z, w4[u].w};
        const float4 b4 = *reinterpret_cast<const float4*>(&bnd_s[col[u] << 2]);
        const float bnd[4] = {b4.x, b4.y, b4.z, b4.w};
        float go[4], term[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float r = fmaxf(bnd[k] - w[k], 0.f);
          const float den = w[k] + 1e-6f;
          if (!BWD) term[k] = (r * r) / den;
          else go[k] = g_scale * (-2.f * r / den - (r * r) / (den * den));
        }
        // the four non-negative terms of a quad are summed in fp32 (pairwise), quads in fp64
        if (!BWD) acc += (double)((term[0] + term[1]) + (term[2] + term[3]));
        if (BWD) reinterpret_cast<float4*>(g_w_hat)[q] = make_float4(go[0], go[1], go[2], go[3]);
      }
    }
    if (!BWD) block_sum_to_partial<256>(acc, partials);
    return;
  }
  // column index carried along instead of a 64-bit modulo per element: i = e mod N advances by (stride mod N) per step
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int step = (int)(stride % N);
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int i = (int)(e % N);
  for (; e < total; e += stride, i = (i + step >= N ? i + step - N : i + step)) {
    const float bnd = bound_mode == 0 ? (float)bound_total[i] : b_per_ray[e];
    const float w = w_hat[e];
    const float r = fmaxf(bnd - w, 0.f);
    const float den = w + 1e-6f;
    if (!BWD) {
      acc += (double)((r * r) / den);
    } else {
      // d/dw [ r^2/(w+eps) ] = -2 r/(w+eps) - r^2/(w+eps)^2
      g_w_hat[e] = g_scale * (-2.f * r / den - (r * r) / (den * den));
    }
  }
  if (!BWD) block_sum_to_partial<256>(acc, partials);
}

// N in {32, 64, 128}: 8 lanes per ray (ray_group.cuh); the prefix sums that are differenced stay in fp64
template <int E, bool BWD>
__global__ void __launch_bounds__(RG_THREADS)
distortion_rg_kernel(const float* __restrict__ s_vals, const float* __restrict__ weights, int B,
                     float* __restrict__ per_ray, double* __restrict__ partials, const float* __restrict__ g_loss_ptr,
                     float* __restrict__ g_w) {
  constexpr int N = E * RG_LANES;
  const int gl = threadIdx.x & 7, j0 = gl * E;
  double block_acc = 0.0;
  float g_loss = 1.f;
  if (BWD) g_loss = *g_loss_ptr;
  for (long long base = (long long)blockIdx.x * RG_RAYS_PER_BLOCK; base < B; base += (long long)gridDim.x * RG_RAYS_PER_BLOCK) {
    const long long ray_raw = base + (threadIdx.x >> 3);
    const bool active = ray_raw < B;
    const long long ray = active ? ray_raw : B - 1;
    float s[E + 1], w[E];
    rg_load_knots<E>(s_vals + ray * (N + 1), j0, s);
    rg_load<E>(weights + ray * N + j0, w);
    double eW[E], eWM[E], runW = 0.0, runWM = 0.0;
    float m[E];
#pragma unroll
    for (int i = 0; i < E; ++i) {
      m[i] = 0.5f * (s[i] + s[i + 1]);
      eW[i] = runW;
      eWM[i] = runWM;
      runW += (double)w[i];
      runWM += (double)w[i] * (double)m[i];
    }
    const double offW = rg_scan_excl(runW, gl), offWM = rg_scan_excl(runWM, gl);
    if (!BWD) {
      // the pair term is a difference of prefix sums (fp64); the self term is a plain positive sum (fp32)
      double loss = 0.0;
      float self = 0.f;
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const double Wl = offW + eW[i], WMl = offWM + eWM[i];
        loss += (double)w[i] * ((double)m[i] * Wl - WMl);
        self += (w[i] * w[i]) * (s[i + 1] - s[i]);
      }
      loss = rg_sum(2.0 * loss + (double)self * (1.0 / 3.0));
      if (gl == 0 && active) {
        if (per_ray) per_ray[ray] = (float)loss;
        block_acc += loss;
      }
    } else {
      const double Wtot = __shfl_sync(FULL_MASK, offW + runW, RG_LANES - 1, RG_LANES);
      const double WMtot = __shfl_sync(FULL_MASK, offWM + runWM, RG_LANES - 1, RG_LANES);
      float g[E];
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const double Wl = offW + eW[i], WMl = offWM + eWM[i];
        const double Wr = Wtot - Wl - (double)w[i], WMr = WMtot - WMl - (double)w[i] * (double)m[i];
        const double gi = 2.0 * ((double)m[i] * (Wl - Wr) - (WMl - WMr)) +
                          (2.0 / 3.0) * (double)w[i] * (double)(s[i + 1] - s[i]);
        g[i] = (float)(gi * (double)g_loss);
      }
      if (active) rg_store<E>(g_w + ray * N + j0, g);
    }
  }
  if (!BWD) block_sum_to_partial<RG_THREADS>(block_acc, partials);
}

// per-ray proposal bounds, 8 lanes per ray: fine knots are read with coalesced cyclic loads straight into shared
// memory (skewed layout: index + index/8, conflict-free for both cyclic and blocked access), fp64 prefix weights
// beside them; each lane resolves E coarse intervals (strided, so the group writes 8 consecutive outputs).
template <int E>
__device__ __forceinline__ int rg_skew_l(int i) {
  return i + (i >> (E == 4 ? 2 : E == 8 ? 3 : 4));
}

template <int E>
__global__ void __launch_bounds__(RG_THREADS)
bounds_rg_kernel(const float* __restrict__ t_fine, const float* __restrict__ w_fine, const float* __restrict__ t_coarse,
                 int B, float* __restrict__ b_out, double* __restrict__ total) {
  constexpr int N = E * RG_LANES, K = N + 1, ROW = K + K / E + 2;
  __shared__ float s_tf[RG_RAYS_PER_BLOCK][ROW];
  __shared__ double s_cw[RG_RAYS_PER_BLOCK][ROW];
  const int gl = threadIdx.x & 7, g = threadIdx.x >> 3, j0 = gl * E;
  float* tfs = s_tf[g];
  double* cws = s_cw[g];
  double col[E];  // column sums (columns gl + 8c) over the rays this lane group has seen
#pragma unroll
  for (int c = 0; c < E; ++c) col[c] = 0.0;
  for (long long base = (long long)blockIdx.x * RG_RAYS_PER_BLOCK; base < B; base += (long long)gridDim.x * RG_RAYS_PER_BLOCK) {
    const long long ray_raw = base + g;
    const bool active = ray_raw < B;
    const long long ray = active ? ray_raw : B - 1;
    const float* tfrow = t_fine + ray * K;
#pragma unroll
    for (int c = 0; c < E; ++c) tfs[rg_skew_l<E>(gl + RG_LANES * c)] = __ldg(tfrow + gl + RG_LANES * c);
    if (gl == 0) tfs[rg_skew_l<E>(N)] = __ldg(tfrow + N);
    float w[E];
    rg_load<E>(w_fine + ray * N + j0, w);
    // the coarse knots are fetched now, together with the fine row, so that the searches below do not start a
    // second round trip to memory.  A lane owns the knots gl + 8c (c < E); slot E is knot N (used by lane 0).
    constexpr int S = E + 1;
    const float* tc = t_coarse + ray * K;
    float X[S];
#pragma unroll
    for (int c = 0; c < S; ++c) X[c] = __ldg(tc + ((c < E) ? gl + RG_LANES * c : N));
    double run = 0.0, ex[E];
#pragma unroll
    for (int i = 0; i < E; ++i) {
      ex[i] = run;
      run += (double)w[i];
    }
    const double off = rg_scan_excl(run, gl);
#pragma unroll
    for (int i = 0; i < E; ++i) cws[rg_skew_l<E>(j0 + i)] = off + ex[i];
    if (gl == RG_LANES - 1) cws[rg_skew_l<E>(N)] = off + run;
    __syncwarp();
    // ONE search per coarse knot x (not two per interval): with
    //   A(x) = #{j < N : t1_j = tf[j+1] < x}   and   B(x) = #{j < N : t0_j = tf[j] <= x}
    // interval i = [x_i, x_{i+1}] needs first = A(x_i) and nR = B(x_{i+1}), and the two counts of one knot differ
    // only by its ties:  B(x) = A(x) + [tf[0] < x] + #{m : tf[m] == x} - [tf[N] <= x].
    // A(x) by the two-level counting search.  Level 1 (registers): t1 of each lane chunk's last interval
    // (tf[(l+1)E]), bisection with selects, the last end separately since all 8 chunks can be full.  Level 2:
    // log2(E) shared-memory probes inside the one partial chunk, the lane's knots in lock step.
    float endF[RG_LANES];
#pragma unroll
    for (int l = 0; l < RG_LANES; ++l) endF[l] = tfs[rg_skew_l<E>((l + 1) * E)];
    const float tf0 = tfs[0], tfN = endF[RG_LANES - 1];
    int first[S], baseF[S];
#pragma unroll
    for (int c = 0; c < S; ++c) {
      const bool f1 = endF[3] < X[c];
      const float fe2 = f1 ? endF[5] : endF[1];
      const bool f2 = fe2 < X[c];
      const float fe3 = f1 ? (f2 ? endF[6] : endF[4]) : (f2 ? endF[2] : endF[0]);
      int nf = (f1 ? 4 : 0) + (f2 ? 2 : 0) + (fe3 < X[c] ? 1 : 0);
      if (nf == 7 && endF[7] < X[c]) nf = 8;
      baseF[c] = nf * (E + 1);  // skewed position of the partial chunk: knot E*n + o sits at (E+1)*n + o for o < E
      first[c] = 0;
    }
#pragma unroll
    for (int step = E / 2; step > 0; step >>= 1) {
#pragma unroll
      for (int c = 0; c < S; ++c) {
        // (the all-chunks-full case probes past the row's knots into its padding: harmless, the count is N then)
        const float vf = tfs[min(baseF[c] + first[c] + step, ROW - 1)];  // t1 of interval E*nf + (first+step) - 1
        if (vf < X[c]) first[c] += step;
      }
    }
    int cntA[S], cntB[S];
#pragma unroll
    for (int c = 0; c < S; ++c) {
      const int nf = baseF[c] / (E + 1);
      cntA[c] = nf == RG_LANES ? N : nf * E + first[c];
      int p = cntA[c] + (tf0 < X[c] ? 1 : 0);                        // knots (0..N) strictly below x
      while (p <= N && tfs[rg_skew_l<E>(p)] == X[c]) ++p;             // + its ties (0 or 1 except on collapsed rows)
      cntB[c] = p - (tfN <= X[c] ? 1 : 0);
    }
#pragma unroll
    for (int c = 0; c < E; ++c) {
      // B of knot i + 1: the next lane's knot of the same slot, or lane 0's next slot for the last lane
      const int b_dn = __shfl_down_sync(FULL_MASK, cntB[c], 1, RG_LANES);
      const int b_l0 = __shfl_sync(FULL_MASK, cntB[c + 1], 0, RG_LANES);
      const int f = cntA[c], r = gl == RG_LANES - 1 ? b_l0 : b_dn;
      float v = 0.f;
      if (r - 1 >= f) v = fmaxf((float)(cws[rg_skew_l<E>(r)] - cws[rg_skew_l<E>(f)]), 0.f);
      if (active) {
        if (b_out) b_out[ray * N + gl + RG_LANES * c] = v;
        col[c] += (double)v;
      }
    }
    __syncwarp();
  }
  if (total) {
    // batch totals per coarse interval (App. A6): the 16 lane groups of the block fold their column sums through
    // the (now idle) prefix-weight rows, one atomic per column and block
    __syncthreads();
#pragma unroll
    for (int c = 0; c < E; ++c) cws[gl + RG_LANES * c] = col[c];
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += RG_THREADS) {
      double a = 0.0;
#pragma unroll
      for (int q = 0; q < RG_RAYS_PER_BLOCK; ++q) a += s_cw[q][i];
      atomicAdd(&total[i], a);
    }
  }
}

static inline int rg_loss_grid(int B) {
  const int g = rg_grid(B);
  return g < LS_MAX_PARTIALS ? g : LS_MAX_PARTIALS;
}

static inline int ray_grid(int B, int warps) {
  long long b = ((long long)B + warps - 1) / warps;
  const long long cap = min((long long)sm_count() * 16, (long long)LS_MAX_PARTIALS);
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace mip360

using namespace mip360;

extern "C" {

int mip360_partials_len(int B) {
  (void)B;
  return LS_MAX_PARTIALS;
}

int mip360_distortion_fwd(const float* s_vals, const float* weights, int B, int N, float* per_ray, double* partials,
                          float* loss, mip360_stream_t stream) {
  MIP_REQUIRE(partials && loss && (B <= 0 || (s_vals && weights)), "distortion_fwd: null pointer");
  MIP_REQUIRE(N >= 1 && N <= MIP360_MAX_SAMPLES, "distortion_fwd: N=%d outside [1,%d]", N, MIP360_MAX_SAMPLES);
  MIP_REQUIRE(B >= 0, "distortion_fwd: B=%d", B);
  const bool rg = rg_supported_host(N);
  const int grid = B > 0 ? (rg ? rg_loss_grid(B) : ray_grid(B, LS_WARPS)) : 0;
  if (grid > 0) {
    cudaStream_t st = (cudaStream_t)stream;
    if (rg && N == 32) distortion_rg_kernel<4, false><<<grid, RG_THREADS, 0, st>>>(s_vals, weights, B, per_ray, partials, nullptr, nullptr);
    else if (rg && N == 64) distortion_rg_kernel<8, false><<<grid, RG_THREADS, 0, st>>>(s_vals, weights, B, per_ray, partials, nullptr, nullptr);
    else if (rg && N == 128) distortion_rg_kernel<16, false><<<grid, RG_THREADS, 0, st>>>(s_vals, weights, B, per_ray, partials, nullptr, nullptr);
    else LS_GENERIC(N, distortion_kernel<false,)<<<grid, LS_WARPS * 32, 0, st>>>(s_vals, weights, B, N, per_ray, partials, nullptr, nullptr);
    MIP_LAUNCH_CHECK();
  }
  reduce_partials_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, grid, 1.0, loss);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_distortion_bwd(const float* s_vals, const float* weights, int B, int N, const float* g_loss_ptr, float* g_w,
                          mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (s_vals && weights && g_loss_ptr && g_w), "distortion_bwd: null pointer");
  MIP_REQUIRE(N >= 1 && N <= MIP360_MAX_SAMPLES, "distortion_bwd: N=%d outside [1,%d]", N, MIP360_MAX_SAMPLES);
  if (B <= 0) return MIP360_OK;
  {
    cudaStream_t st = (cudaStream_t)stream;
    const bool rg = rg_supported_host(N);
    if (rg && N == 32) distortion_rg_kernel<4, true><<<rg_grid(B), RG_THREADS, 0, st>>>(s_vals, weights, B, nullptr, nullptr, g_loss_ptr, g_w);
    else if (rg && N == 64) distortion_rg_kernel<8, true><<<rg_grid(B), RG_THREADS, 0, st>>>(s_vals, weights, B, nullptr, nullptr, g_loss_ptr, g_w);
    else if (rg && N == 128) distortion_rg_kernel<16, true><<<rg_grid(B), RG_THREADS, 0, st>>>(s_vals, weights, B, nullptr, nullptr, g_loss_ptr, g_w);
    else LS_GENERIC(N, distortion_kernel<true,)<<<ray_grid(B, LS_WARPS), LS_WARPS * 32, 0, st>>>(s_vals, weights, B, N, nullptr, nullptr, g_loss_ptr, g_w);
  }
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_bounds_per_ray(const float* t_fine, const float* w_fine, const float* t_coarse, int B, int N, float* b_out,
                          mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || b_out, "bounds_per_ray: null pointer");
  return mip360_bounds(t_fine, w_fine, t_coarse, B, N, b_out, nullptr, stream);
}

int mip360_bounds(const float* t_fine, const float* w_fine, const float* t_coarse, int B, int N, float* b_out,
                  double* bound_total, mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (t_fine && w_fine && t_coarse && (b_out || bound_total)), "bounds: null pointer");
  MIP_REQUIRE(N >= 1 && N <= MIP360_MAX_SAMPLES, "bounds: N=%d outside [1,%d]", N, MIP360_MAX_SAMPLES);
  if (B <= 0) return MIP360_OK;
  {
    cudaStream_t st = (cudaStream_t)stream;
    const bool rg = rg_supported_host(N);
    // with totals every block ends in N atomics: cap the grid at a few blocks per SM
    const int grid_rg = bound_total ? min(rg_grid(B), sm_count() * 8) : rg_grid(B);
    if (rg && N == 32) bounds_rg_kernel<4><<<grid_rg, RG_THREADS, 0, st>>>(t_fine, w_fine, t_coarse, B, b_out, bound_total);
    else if (rg && N == 64) bounds_rg_kernel<8><<<grid_rg, RG_THREADS, 0, st>>>(t_fine, w_fine, t_coarse, B, b_out, bound_total);
    else if (rg && N == 128) bounds_rg_kernel<16><<<grid_rg, RG_THREADS, 0, st>>>(t_fine, w_fine, t_coarse, B, b_out, bound_total);
    else LS_GENERIC(N, bounds_kernel<)<<<ray_grid(B, LS_WARPS), LS_WARPS * 32, 0, st>>>(t_fine, w_fine, t_coarse, B, N, b_out, bound_total);
  }
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_bounds_reduce(const float* b, int B, int N, double* bound_total, mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (b && bound_total), "bounds_reduce: null pointer");
  MIP_REQUIRE(N >= 1 && N <= MIP360_MAX_SAMPLES, "bounds_reduce: N=%d outside [1,%d]", N, MIP360_MAX_SAMPLES);
  if (B <= 0) return MIP360_OK;
  const int target_blocks = sm_count() * 4;
  int rows = (B + target_blocks - 1) / target_blocks;
  if (rows < 16) rows = 16;
  const int grid = (B + rows - 1) / rows;
  bounds_reduce_kernel<<<grid, dim3(128, 2), 0, (cudaStream_t)stream>>>(b, B, N, rows, bound_total);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_interlevel_fwd(const float* w_hat, const float* b_per_ray, const double* bound_total, int B, int N,
                          int bound_mode, float batch_div, double* partials, float* loss, mip360_stream_t stream) {
  MIP_REQUIRE(partials && loss && (B <= 0 || w_hat), "interlevel_fwd: null pointer");
  MIP_REQUIRE(B <= 0 || (bound_mode == 0 ? bound_total != nullptr : b_per_ray != nullptr), "interlevel_fwd: bound missing");
  MIP_REQUIRE(N >= 1 && N <= MIP360_MAX_SAMPLES, "interlevel_fwd: N=%d outside [1,%d]", N, MIP360_MAX_SAMPLES);
  const long long total = (long long)B * N;
  int grid = 0;
  if (total > 0) {
    grid = (int)min((total + 255) / 256, (long long)min(sm_count() * 8, LS_MAX_PARTIALS));
    interlevel_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(w_hat, b_per_ray, bound_total, total, N,
                                                                      bound_mode, batch_div, partials, nullptr, nullptr);
    MIP_LAUNCH_CHECK();
  }
  reduce_partials_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, grid, 1.0 / (double)batch_div, loss);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_interlevel_bwd(const float* w_hat, const float* b_per_ray, const double* bound_total, int B, int N,
                          int bound_mode, float batch_div, const float* g_loss_ptr, float* g_w_hat,
                          mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (w_hat && g_loss_ptr && g_w_hat), "interlevel_bwd: null pointer");
  MIP_REQUIRE(B <= 0 || (bound_mode == 0 ? bound_total != nullptr : b_per_ray != nullptr), "interlevel_bwd: bound missing");
  MIP_REQUIRE(N >= 1 && N <= MIP360_MAX_SAMPLES, "interlevel_bwd: N=%d outside [1,%d]", N, MIP360_MAX_SAMPLES);
  const long long total = (long long)B * N;
  if (total <= 0) return MIP360_OK;
  const int grid = (int)min((total + 255) / 256, (long long)sm_count() * 8);
  interlevel_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(w_hat, b_per_ray, bound_total, total, N, bound_mode,
                                                                   batch_div, nullptr, g_loss_ptr, g_w_hat);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

}  // extern "C"
