// K3: volume compositing (intern/ray.py:155-191, model.py:59-78, model.py:184-185) forward and
// backward, plus the t<->s maps of intern/parameterization.py:5-13.  One warp per ray: the ray's knots
// and per-sample values are staged in shared memory with coalesced loads, each lane owns a contiguous
// chunk of ceil(N/32) intervals, the transmittance is a warp exclusive scan.
#include "common.cuh"
#include "ray_group.cuh"

namespace mip360 {

constexpr int CP_WARPS = 2;  // generic kernels: 14 KB of staging per ray at the N = 512 limit
// intervals per lane: 4 covers N <= 128, 16 the N <= 512 limit (template parameter MAXC of the generic kernels)
constexpr int CP_MAXC_SMALL = 4, CP_MAXC_LARGE = (MIP360_MAX_SAMPLES + 31) / 32;

struct __align__(16) CompositeSmem {
  float t[MIP360_MAX_SAMPLES + 1];
  float sigma[MIP360_MAX_SAMPLES];       // density after activation
  float aux[MIP360_MAX_SAMPLES];         // head_mode 1: post-sigmoid density head value y0
  float rgb[MIP360_MAX_SAMPLES * 3];     // colour after padding
  float gw[MIP360_MAX_SAMPLES];          // backward: total dL/dw
};

template <int CP_MAXC>
struct RayScan {
  float dd[CP_MAXC], T[CP_MAXC], w[CP_MAXC], delta[CP_MAXC];
};

// weights of one ray: lane-chunked exclusive scan of sigma*delta
template <int CP_MAXC>
__device__ __forceinline__ void ray_weights(const CompositeSmem& s, int N, int C, int j0, float dnorm, int lane,
                                            RayScan<CP_MAXC>& r) {
  float run = 0.f;
  float excl[CP_MAXC];
#pragma unroll
  for (int c = 0; c < CP_MAXC; ++c) {
    const int j = j0 + c;
    r.dd[c] = 0.f;
    r.delta[c] = 0.f;
    if (c < C && j < N) {
      r.delta[c] = (s.t[j + 1] - s.t[j]) * dnorm;
      r.dd[c] = s.sigma[j] * r.delta[c];
    }
    excl[c] = run;
    run += r.dd[c];
  }
  const float off = warp_scan_incl(run, lane) - run;
#pragma unroll
  for (int c = 0; c < CP_MAXC; ++c) {
    r.T[c] = expf(-(off + excl[c]));
    const float alpha = 1.f - expf(-r.dd[c]);
    r.w[c] = alpha * r.T[c];
  }
}

// head_mode 2: raw holds the heads' pre-activation sums without bias -> Sigmoid(raw + bias) (model.py:150-158)
// The head activation as the GEMM epilogue evaluates it for an unfused head (1 / (1 + e^-z) with the fast exp and
// reciprocal: ~2 ulp, and the same values whichever side applies the Sigmoid)
__device__ __forceinline__ float head_sigmoid_one(float z) { return __fdividef(1.f, 1.f + __expf(-z)); }
__device__ __forceinline__ float4 head_sigmoid(float4 v, const float* __restrict__ head_bias) {
  return make_float4(head_sigmoid_one(v.x + __ldg(head_bias)), head_sigmoid_one(v.y + __ldg(head_bias + 1)),
                     head_sigmoid_one(v.z + __ldg(head_bias + 2)), head_sigmoid_one(v.w + __ldg(head_bias + 3)));
}
// ... and its derivative y (1 - y) folded into the gradient of the four head outputs; y0 is kept, the colours are
// recovered from the padded values c = y * scale - pad
__device__ __forceinline__ float4 head_sigmoid_bwd(float4 g, float y0, float c1, float c2, float c3, float scale, float pad) {
  const float y1 = (c1 + pad) / scale, y2 = (c2 + pad) / scale, y3 = (c3 + pad) / scale;
  return make_float4(g.x * (y0 * (1.f - y0)), g.y * (y1 * (1.f - y1)), g.z * (y2 * (1.f - y2)), g.w * (y3 * (1.f - y3)));
}

// stage one ray.  mode_full: rgb too.  head_mode 1: raw [N,4] post-sigmoid head outputs.
__device__ __forceinline__ void stage_ray(CompositeSmem& s, const float* rgb_or_raw, const float* density,
                                          const float* t_vals, long long b, int N, int head_mode, bool with_rgb,
                                          int density_mode, float density_bias, float rgb_padding, int lane,
                                          const float* head_bias) {
  const float* trow = t_vals + b * (N + 1);
  for (int k = lane; k <= N; k += 32) s.t[k] = trow[k];
  if (head_mode >= 1) {
    const float4* raw4 = reinterpret_cast<const float4*>(rgb_or_raw) + b * N;
    const float scale = 1.f + 2.f * rgb_padding;
    for (int j = lane; j < N; j += 32) {
      float4 v = raw4[j];
      if (head_mode == 2) v = head_sigmoid(v, head_bias);
      s.aux[j] = v.x;
      s.sigma[j] = softplus_f(v.x + density_bias);
      s.rgb[j * 3 + 0] = v.y * scale - rgb_padding;
      s.rgb[j * 3 + 1] = v.z * scale - rgb_padding;
      s.rgb[j * 3 + 2] = v.w * scale - rgb_padding;
    }
  } else {
    const float* drow = density + b * N;
    for (int j = lane; j < N; j += 32) {
      const float v = drow[j];
      s.aux[j] = v;
      s.sigma[j] = density_mode == 1 ? softplus_f(v + density_bias) : v;
    }
    if (with_rgb) {
      const float* crow = rgb_or_raw + b * N * 3;
      for (int e = lane; e < N * 3; e += 32) s.rgb[e] = crow[e];
    }
  }
}

// intern/parameterization.py:5-8 with the eps shifts a single reference call observes (App. A4); see t_to_s_kernel
// per ray: 1/(near+eps) and the denominator 1/(far+eps) - 1/(near+2eps); per knot: two divisions instead of four
struct TToS {
  float inv_n1, den;
};
__device__ __forceinline__ TToS t_to_s_ray(float near, float far) {
  const float n1 = near + G_EPS;
  const float f1 = far + G_EPS;
  const float n2 = n1 + G_EPS;
  return TToS{1.f / n1, 1.f / f1 - 1.f / n2};
}
__device__ __forceinline__ void t_to_s_one(float t, const TToS& r, float& s, float& t_shift) {
  const float t1 = t + G_EPS;
  s = (1.f / t1 - r.inv_n1) / r.den;
  t_shift = t1;
}

template <int CP_MAXC>
__global__ void __launch_bounds__(CP_WARPS * 32)
composite_fwd_kernel(const float* __restrict__ rgb_or_raw, const float* __restrict__ density,
                     const float* __restrict__ t_vals, const float* __restrict__ dirs, int B, int N, int head_mode,
                     int weights_only, int density_mode, float density_bias, float rgb_padding, int white_bkgd,
                     float* __restrict__ comp_rgb, float* __restrict__ distance, float* __restrict__ acc_out,
                     float* __restrict__ weights, const float* __restrict__ near, const float* __restrict__ far,
                     float* __restrict__ s_vals, float* __restrict__ t_shift, const float* __restrict__ head_bias) {
  __shared__ CompositeSmem sm[CP_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  CompositeSmem& s = sm[warp];
  const int C = (N + 31) >> 5, j0 = lane * C;
  for (int b = blockIdx.x * CP_WARPS + warp; b < B; b += gridDim.x * CP_WARPS) {
    stage_ray(s, rgb_or_raw, density, t_vals, b, N, head_mode, !weights_only, density_mode, density_bias, rgb_padding,
              lane, head_bias);
    const float dx = dirs[b * 3], dy = dirs[b * 3 + 1], dz = dirs[b * 3 + 2];
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    __syncwarp();
    if (s_vals) {  // model.py:196 fused: s_vals = t_to_s(t_vals, near, far) of the same knots
      const TToS tr = t_to_s_ray(near[b], far[b]);
      for (int k = lane; k <= N; k += 32) {
        float sv, ts;
        t_to_s_one(s.t[k], tr, sv, ts);
        s_vals[(long long)b * (N + 1) + k] = sv;
        if (t_shift) t_shift[(long long)b * (N + 1) + k] = ts;
      }
    }
    RayScan<CP_MAXC> r;
    ray_weights(s, N, C, j0, dnorm, lane, r);
    float a = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, wt = 0.f;
#pragma unroll
    for (int c = 0; c < CP_MAXC; ++c) {
      const int j = j0 + c;
      if (c < C && j < N) {
        if (weights) weights[(long long)b * N + j] = r.w[c];
        a += r.w[c];
        if (!weights_only) {
          cr += r.w[c] * s.rgb[j * 3 + 0];
          cg += r.w[c] * s.rgb[j * 3 + 1];
          cb += r.w[c] * s.rgb[j * 3 + 2];
          wt += r.w[c] * (0.5f * (s.t[j] + s.t[j + 1]));
        }
      }
    }
    if (!weights_only) {
      a = warp_sum(a); cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); wt = warp_sum(wt);
      if (lane == 0) {
        float dist = nan_to_num_f(wt / a);
        dist = fminf(fmaxf(dist, s.t[0]), s.t[N]);
        if (white_bkgd) {
          const float bg = 1.f - a;
          cr += bg; cg += bg; cb += bg;
        }
        comp_rgb[b * 3 + 0] = cr;
        comp_rgb[b * 3 + 1] = cg;
        comp_rgb[b * 3 + 2] = cb;
        distance[b] = dist;
        acc_out[b] = a;
      }
    }
    __syncwarp();
  }
}

// Backward.  G_j = g_w_j + g_rgb . c_j + g_acc' is the total gradient reaching w_j;
//   dL/d(dd_j) = G_j T_j e^{-dd_j} - sum_{k>j} G_k w_k ,  dL/dsigma_j = dL/d(dd_j) delta_j.
template <int CP_MAXC>
__global__ void __launch_bounds__(CP_WARPS * 32)
composite_bwd_kernel(const float* __restrict__ rgb_or_raw, const float* __restrict__ density,
                     const float* __restrict__ t_vals, const float* __restrict__ dirs, int B, int N, int head_mode,
                     int weights_only, int density_mode, float density_bias, float rgb_padding, int white_bkgd,
                     const float* __restrict__ g_rgb, const float* __restrict__ g_acc, const float* __restrict__ g_dist,
                     const float* __restrict__ g_w, float* __restrict__ g_rgb_in, float* __restrict__ g_density,
                     float* __restrict__ g_raw, const float* __restrict__ head_bias) {
  __shared__ CompositeSmem sm[CP_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  CompositeSmem& s = sm[warp];
  const int C = (N + 31) >> 5, j0 = lane * C;
  const float cscale = 1.f + 2.f * rgb_padding;
  for (int b = blockIdx.x * CP_WARPS + warp; b < B; b += gridDim.x * CP_WARPS) {
    stage_ray(s, rgb_or_raw, density, t_vals, b, N, head_mode, !weights_only, density_mode, density_bias, rgb_padding,
              lane, head_bias);
    if (g_w) {
      for (int j = lane; j < N; j += 32) s.gw[j] = g_w[(long long)b * N + j];
    }
    const float dx = dirs[b * 3], dy = dirs[b * 3 + 1], dz = dirs[b * 3 + 2];
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    float gr = 0.f, gg = 0.f, gb = 0.f, ga = 0.f, gd = 0.f;
    if (!weights_only) {
      if (g_rgb) { gr = g_rgb[b * 3]; gg = g_rgb[b * 3 + 1]; gb = g_rgb[b * 3 + 2]; }
      if (g_acc) ga = g_acc[b];
      if (g_dist) gd = g_dist[b];
      if (white_bkgd) ga -= (gr + gg + gb);
    }
    __syncwarp();
    RayScan<CP_MAXC> r;
    ray_weights(s, N, C, j0, dnorm, lane, r);
    // distance = clamp(nan_to_num(sum w t_mid / acc), t_0, t_N): d/dw_j = (t_mid_j - distance)/acc where the
    // quotient is finite and inside the clamp range (torch.clamp passes the gradient on the closed range)
    float gd_scale = 0.f, dist_raw = 0.f;
    if (g_dist) {
      float a = 0.f, wt = 0.f;
#pragma unroll
      for (int c = 0; c < CP_MAXC; ++c) {
        const int j = j0 + c;
        if (c < C && j < N) { a += r.w[c]; wt += r.w[c] * (0.5f * (s.t[j] + s.t[j + 1])); }
      }
      a = warp_sum(a); wt = warp_sum(wt);
      dist_raw = wt / a;
      const bool pass = !isnan(dist_raw) && !isinf(dist_raw) && dist_raw >= s.t[0] && dist_raw <= s.t[N];
      gd_scale = pass ? gd / a : 0.f;
      if (!pass) dist_raw = 0.f;  // keeps 0 * (t_mid - NaN) out of G
    }
    float G[CP_MAXC], run = 0.f, excl_rev[CP_MAXC];
    // suffix sums of G_k w_k: walk the lane's chunk backwards
#pragma unroll
    for (int c = CP_MAXC - 1; c >= 0; --c) {
      const int j = j0 + c;
      G[c] = 0.f;
      if (c < C && j < N) {
        G[c] = (g_w ? s.gw[j] : 0.f) + ga;
        if (!weights_only) G[c] += gr * s.rgb[j * 3] + gg * s.rgb[j * 3 + 1] + gb * s.rgb[j * 3 + 2];
        if (g_dist) G[c] += gd_scale * (0.5f * (s.t[j] + s.t[j + 1]) - dist_raw);
      }
      excl_rev[c] = run;
      run += G[c] * r.w[c];
    }
    const float off = warp_scan_incl_rev(run, lane) - run;
#pragma unroll
    for (int c = 0; c < CP_MAXC; ++c) {
      const int j = j0 + c;
      if (c < C && j < N) {
        const float g_dd = G[c] * r.T[c] * expf(-r.dd[c]) - (off + excl_rev[c]);
        const float g_sigma = g_dd * r.delta[c];
        const long long e = (long long)b * N + j;
        if (head_mode >= 1) {
          const float y0 = s.aux[j];
          // sigma = softplus(y0 + bias): d/dy0 = sigmoid(y0 + bias); colour = y*(1+2p) - p
          const float g_y0 = g_sigma * sigmoid_f(y0 + density_bias);
          const float g_y1 = r.w[c] * gr * cscale, g_y2 = r.w[c] * gg * cscale, g_y3 = r.w[c] * gb * cscale;
          float4 gy = make_float4(g_y0, g_y1, g_y2, g_y3);
          if (head_mode == 2)
            gy = head_sigmoid_bwd(gy, y0, s.rgb[j * 3], s.rgb[j * 3 + 1], s.rgb[j * 3 + 2], cscale, rgb_padding);
          if (g_raw) reinterpret_cast<float4*>(g_raw)[e] = gy;
        } else {
          if (density_mode == 1) {
            const float g_z = g_sigma * sigmoid_f(s.aux[j] + density_bias);
            if (g_density) g_density[e] = g_z;
          } else if (g_density) {
            g_density[e] = g_sigma;
          }
          if (!weights_only && g_rgb_in) {
            g_rgb_in[e * 3 + 0] = r.w[c] * gr;
            g_rgb_in[e * 3 + 1] = r.w[c] * gg;
            g_rgb_in[e * 3 + 2] = r.w[c] * gb;
          }
        }
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------
// N in {32, 64, 128}: 8 lanes per ray, E = N/8 intervals per lane in registers (ray_group.cuh)
// ---------------------------------------------------------------------------------------------------
template <int E>
struct RgRay {
  float t[E + 1], sigma[E], aux[E], delta[E], dd[E], T[E], w[E];
  float c[E][3];  // colours (only filled when wanted)
  float d[3];     // ray direction
};

// Every global load of the lane is issued here, before any arithmetic: the activation code below branches
// (softplus threshold, log1pf), and loads placed after a branch would start a second round trip to memory.
template <int E>
__device__ __forceinline__ void rg_fetch(RgRay<E>& r, const float* __restrict__ rgb_or_raw,
                                         const float* __restrict__ density, const float* __restrict__ t_vals,
                                         const float* __restrict__ dirs, long long ray, int N, int gl, int head_mode,
                                         bool want_rgb, float rgb_padding, const float* __restrict__ head_bias) {
  const int j0 = gl * E;
  rg_load_knots<E>(t_vals + ray * (N + 1), j0, r.t);
#pragma unroll
  for (int i = 0; i < 3; ++i) r.d[i] = __ldg(dirs + ray * 3 + i);
  if (head_mode >= 1) {
    // one 16-byte load per sample: density logit and colour together (whole sectors, a single pass over the rows)
    const float4* raw4 = reinterpret_cast<const float4*>(rgb_or_raw) + ray * N + j0;
    const float scale = 1.f + 2.f * rgb_padding;
    float4 v[E];
#pragma unroll
    for (int i = 0; i < E; ++i) v[i] = __ldg(raw4 + i);
    if (head_mode == 2) {
#pragma unroll
      for (int i = 0; i < E; ++i) v[i] = head_sigmoid(v[i], head_bias);
    }
#pragma unroll
    for (int i = 0; i < E; ++i) {
      r.aux[i] = v[i].x;
      if (want_rgb) {
        r.c[i][0] = v[i].y * scale - rgb_padding;
        r.c[i][1] = v[i].z * scale - rgb_padding;
        r.c[i][2] = v[i].w * scale - rgb_padding;
      }
    }
  } else {
    rg_load<E>(density + ray * N + j0, r.aux);
    if (want_rgb) {
      float flat[3 * E];
      rg_load<3 * E>(rgb_or_raw + (ray * N + j0) * 3, flat);
#pragma unroll
      for (int i = 0; i < E; ++i) {
        r.c[i][0] = flat[3 * i]; r.c[i][1] = flat[3 * i + 1]; r.c[i][2] = flat[3 * i + 2];
      }
    }
  }
}

// density activation, then the weights of the lane's E intervals
template <int E>
__device__ __forceinline__ void rg_weights(RgRay<E>& r, int gl, int head_mode, int density_mode, float density_bias) {
#pragma unroll
  for (int i = 0; i < E; ++i)
    r.sigma[i] = (head_mode >= 1 || density_mode == 1) ? softplus_f(r.aux[i] + density_bias) : r.aux[i];
  const float dnorm = sqrtf(r.d[0] * r.d[0] + r.d[1] * r.d[1] + r.d[2] * r.d[2]);
  float run = 0.f, excl[E];
#pragma unroll
  for (int i = 0; i < E; ++i) {
    r.delta[i] = (r.t[i + 1] - r.t[i]) * dnorm;
    r.dd[i] = r.sigma[i] * r.delta[i];
    excl[i] = run;
    run += r.dd[i];
  }
  const float off = rg_scan_excl(run, gl);
#pragma unroll
  for (int i = 0; i < E; ++i) {
    r.T[i] = expf(-(off + excl[i]));
    r.w[i] = (1.f - expf(-r.dd[i])) * r.T[i];
  }
}

// WO = weights-only (density_to_weight): a separate instantiation, so the colour registers do not cost it occupancy
template <int E, bool WO>
__global__ void __launch_bounds__(RG_THREADS)
composite_fwd_rg_kernel(const float* __restrict__ rgb_or_raw, const float* __restrict__ density,
                        const float* __restrict__ t_vals, const float* __restrict__ dirs, int B, int head_mode,
                        int /*weights_only*/, int density_mode, float density_bias, float rgb_padding, int white_bkgd,
                        float* __restrict__ comp_rgb, float* __restrict__ distance, float* __restrict__ acc_out,
                        float* __restrict__ weights, const float* __restrict__ near, const float* __restrict__ far,
                        float* __restrict__ s_vals, float* __restrict__ t_shift, const float* __restrict__ head_bias) {
  constexpr int N = E * RG_LANES;
  constexpr bool weights_only = WO;
  const int gl = threadIdx.x & 7, j0 = gl * E;
  for (long long base = (long long)blockIdx.x * RG_RAYS_PER_BLOCK; base < B; base += (long long)gridDim.x * RG_RAYS_PER_BLOCK) {
    const long long ray_raw = base + (threadIdx.x >> 3);
    const bool active = ray_raw < B;
    const long long ray = active ? ray_raw : B - 1;
    RgRay<E> r;
    rg_fetch<E>(r, rgb_or_raw, density, t_vals, dirs, ray, N, gl, head_mode, !weights_only, rgb_padding, head_bias);
    float nr = 0.f, fr = 0.f;
    if (!WO && s_vals) { nr = __ldg(near + ray); fr = __ldg(far + ray); }
    rg_weights<E>(r, gl, head_mode, density_mode, density_bias);
    if (weights && active) rg_store<E>(weights + ray * N + j0, r.w);
    if (weights_only) continue;
    if (!WO && s_vals) {
      // model.py:196 fused: the lane's knots j0 .. j0+E-1 (the last lane also knot N).  The 16 rays of a block are one
      // contiguous run of 16 (N+1) floats in s_vals / t_shift: the lanes park their blocked values in shared memory and
      // the block stores the run with fully coalesced 4-byte stores (a lane's own blocked stores would touch 8 sectors
      // per ray and instruction)
      __shared__ float s_tile[2][RG_RAYS_PER_BLOCK * (N + 1)];
      const int g = threadIdx.x >> 3;
      const int nk = E + (gl == RG_LANES - 1 ? 1 : 0);
      const TToS tr = t_to_s_ray(nr, fr);
#pragma unroll
      for (int i = 0; i <= E; ++i) {
        if (i < nk) {
          float sv, ts;
          t_to_s_one(r.t[i], tr, sv, ts);
          s_tile[0][g * (N + 1) + j0 + i] = sv;
          s_tile[1][g * (N + 1) + j0 + i] = ts;
        }
      }
      __syncthreads();
      const long long rays_here = (B - base) < RG_RAYS_PER_BLOCK ? (B - base) : RG_RAYS_PER_BLOCK;
      const int run = (int)rays_here * (N + 1);
      float* gs = s_vals + base * (N + 1);
      for (int idx = threadIdx.x; idx < run; idx += RG_THREADS) gs[idx] = s_tile[0][idx];
      if (t_shift) {
        float* gt = t_shift + base * (N + 1);
        for (int idx = threadIdx.x; idx < run; idx += RG_THREADS) gt[idx] = s_tile[1][idx];
      }
      __syncthreads();
    }
    float a = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, wt = 0.f;
#pragma unroll
    for (int i = 0; i < E; ++i) {
      a += r.w[i];
      cr += r.w[i] * r.c[i][0];
      cg += r.w[i] * r.c[i][1];
      cb += r.w[i] * r.c[i][2];
      wt += r.w[i] * (0.5f * (r.t[i] + r.t[i + 1]));
    }
    a = rg_sum(a); cr = rg_sum(cr); cg = rg_sum(cg); cb = rg_sum(cb); wt = rg_sum(wt);
    const float t_first = __shfl_sync(FULL_MASK, r.t[0], 0, RG_LANES);
    const float t_last = __shfl_sync(FULL_MASK, r.t[E], RG_LANES - 1, RG_LANES);
    if (gl == 0 && active) {
      float dist = nan_to_num_f(wt / a);
      dist = fminf(fmaxf(dist, t_first), t_last);
      if (white_bkgd) {
        const float bg = 1.f - a;
        cr += bg; cg += bg; cb += bg;
      }
      comp_rgb[ray * 3 + 0] = cr;
      comp_rgb[ray * 3 + 1] = cg;
      comp_rgb[ray * 3 + 2] = cb;
      distance[ray] = dist;
      acc_out[ray] = a;
    }
  }
}

template <int E, bool WO>
__global__ void __launch_bounds__(RG_THREADS)
composite_bwd_rg_kernel(const float* __restrict__ rgb_or_raw, const float* __restrict__ density,
                        const float* __restrict__ t_vals, const float* __restrict__ dirs, int B, int head_mode,
                        int /*weights_only*/, int density_mode, float density_bias, float rgb_padding, int white_bkgd,
                        const float* __restrict__ g_rgb, const float* __restrict__ g_acc, const float* __restrict__ g_dist,
                        const float* __restrict__ g_w, float* __restrict__ g_rgb_in, float* __restrict__ g_density,
                        float* __restrict__ g_raw, const float* __restrict__ head_bias) {
  constexpr int N = E * RG_LANES;
  constexpr bool weights_only = WO;
  const int gl = threadIdx.x & 7, j0 = gl * E;
  const float cscale = 1.f + 2.f * rgb_padding;
  for (long long base = (long long)blockIdx.x * RG_RAYS_PER_BLOCK; base < B; base += (long long)gridDim.x * RG_RAYS_PER_BLOCK) {
    const long long ray_raw = base + (threadIdx.x >> 3);
    const bool active = ray_raw < B;
    const long long ray = active ? ray_raw : B - 1;
    RgRay<E> r;
    rg_fetch<E>(r, rgb_or_raw, density, t_vals, dirs, ray, N, gl, head_mode, !weights_only, rgb_padding, head_bias);
    float gr = 0.f, gg = 0.f, gb = 0.f, ga = 0.f, gd = 0.f;
    if (!weights_only) {
      if (g_rgb) { gr = __ldg(g_rgb + ray * 3); gg = __ldg(g_rgb + ray * 3 + 1); gb = __ldg(g_rgb + ray * 3 + 2); }
      if (g_acc) ga = __ldg(g_acc + ray);
      if (g_dist) gd = __ldg(g_dist + ray);
    }
    float G[E];
    if (g_w) {
      rg_load<E>(g_w + ray * N + j0, G);
    } else {
#pragma unroll
      for (int i = 0; i < E; ++i) G[i] = 0.f;
    }
    // all loads are in flight; arithmetic starts here
    rg_weights<E>(r, gl, head_mode, density_mode, density_bias);
    if (!weights_only && white_bkgd) ga -= (gr + gg + gb);
    // gradient through distance (see composite_bwd_kernel)
    float gd_scale = 0.f, dist_raw = 0.f;
    if (!weights_only && g_dist) {
      float a = 0.f, wt = 0.f;
#pragma unroll
      for (int i = 0; i < E; ++i) { a += r.w[i]; wt += r.w[i] * (0.5f * (r.t[i] + r.t[i + 1])); }
      a = rg_sum(a); wt = rg_sum(wt);
      const float t_first = __shfl_sync(FULL_MASK, r.t[0], 0, RG_LANES);
      const float t_last = __shfl_sync(FULL_MASK, r.t[E], RG_LANES - 1, RG_LANES);
      dist_raw = wt / a;
      const bool pass = !isnan(dist_raw) && !isinf(dist_raw) && dist_raw >= t_first && dist_raw <= t_last;
      gd_scale = pass ? gd / a : 0.f;
      if (!pass) dist_raw = 0.f;  // keeps 0 * (t_mid - NaN) out of G
    }
    float run = 0.f, excl_rev[E];
#pragma unroll
    for (int i = E - 1; i >= 0; --i) {
      G[i] += ga;
      if (!weights_only) G[i] += gr * r.c[i][0] + gg * r.c[i][1] + gb * r.c[i][2];
      if (!weights_only && g_dist) G[i] += gd_scale * (0.5f * (r.t[i] + r.t[i + 1]) - dist_raw);
      excl_rev[i] = run;
      run += G[i] * r.w[i];
    }
    const float off = rg_scan_excl_rev(run, gl);
    float gs[E];
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const float g_dd = G[i] * r.T[i] * expf(-r.dd[i]) - (off + excl_rev[i]);
      gs[i] = g_dd * r.delta[i];
    }
    if (!active) continue;
    if (head_mode >= 1) {
      float4* out = reinterpret_cast<float4*>(g_raw) + ray * N + j0;
#pragma unroll
      for (int i = 0; i < E; ++i) {
        float4 gy = make_float4(gs[i] * sigmoid_f(r.aux[i] + density_bias), r.w[i] * gr * cscale, r.w[i] * gg * cscale,
                                r.w[i] * gb * cscale);
        if (head_mode == 2) gy = head_sigmoid_bwd(gy, r.aux[i], r.c[i][0], r.c[i][1], r.c[i][2], cscale, rgb_padding);
        out[i] = gy;
      }
    } else {
      if (g_density) {
        if (density_mode == 1) {
#pragma unroll
          for (int i = 0; i < E; ++i) gs[i] *= sigmoid_f(r.aux[i] + density_bias);
        }
        rg_store<E>(g_density + ray * N + j0, gs);
      }
      if (!weights_only && g_rgb_in) {
        float flat[3 * E];
#pragma unroll
        for (int i = 0; i < E; ++i) {
          flat[3 * i] = r.w[i] * gr; flat[3 * i + 1] = r.w[i] * gg; flat[3 * i + 2] = r.w[i] * gb;
        }
        rg_store<3 * E>(g_rgb_in + (ray * N + j0) * 3, flat);
      }
    }
  }
}

template <bool WO, typename... Args>
static void launch_composite_fwd(int N, int B, cudaStream_t st, Args... a) {
  if (N == 32) composite_fwd_rg_kernel<4, WO><<<rg_grid(B), RG_THREADS, 0, st>>>(a...);
  else if (N == 64) composite_fwd_rg_kernel<8, WO><<<rg_grid(B), RG_THREADS, 0, st>>>(a...);
  else composite_fwd_rg_kernel<16, WO><<<rg_grid(B), RG_THREADS, 0, st>>>(a...);
}
template <bool WO, typename... Args>
static void launch_composite_bwd(int N, int B, cudaStream_t st, Args... a) {
  if (N == 32) composite_bwd_rg_kernel<4, WO><<<rg_grid(B), RG_THREADS, 0, st>>>(a...);
  else if (N == 64) composite_bwd_rg_kernel<8, WO><<<rg_grid(B), RG_THREADS, 0, st>>>(a...);
  else composite_bwd_rg_kernel<16, WO><<<rg_grid(B), RG_THREADS, 0, st>>>(a...);
}

// intern/parameterization.py:5-8 with the eps shifts a single reference call observes (App. A4)
__global__ void __launch_bounds__(256)
t_to_s_kernel(const float* __restrict__ t_vals, const float* __restrict__ near, const float* __restrict__ far, int B,
              int K, float* __restrict__ s_vals, float* __restrict__ t_shift) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)B * K) return;
  const int b = (int)(e / K);
  const float t1 = t_vals[e] + G_EPS;
  const float n1 = near[b] + G_EPS;
  const float f1 = far[b] + G_EPS;
  const float n2 = n1 + G_EPS;
  s_vals[e] = (1.f / t1 - 1.f / n1) / (1.f / f1 - 1.f / n2);
  if (t_shift) t_shift[e] = t1;
}

// intern/parameterization.py:10-13
__global__ void __launch_bounds__(256)
s_to_t_kernel(const float* __restrict__ s_vals, const float* __restrict__ near, const float* __restrict__ far, int B,
              int K, float* __restrict__ t_vals) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)B * K) return;
  const int b = (int)(e / K);
  const float s = s_vals[e];
  const float gf = 1.f / (far[b] + G_EPS), gn = 1.f / (near[b] + G_EPS);
  t_vals[e] = 1.f / ((s * gf + (1.f - s) * gn) + G_EPS);
}

// Gradient w.r.t. the MLP head outputs -> bf16 rows of 64 for the head dgrad / wgrad GEMMs.
// g [M, nv] fp32 is dL/dy; act 2 (Sigmoid head, model.py:150-158): dL/dz = g * y (1 - y) with the saved
// head output y [M, nv]; act 0: dL/dz = g.  Columns nv..63 are zero.
__global__ void __launch_bounds__(256)
head_grad_pack_kernel(const float* __restrict__ g, const float* __restrict__ y, long long M, int nv, int act,
                      uint16_t* __restrict__ out) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  float z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    z[i] = 0.f;
    if (i < nv) {
      const float gi = g[r * nv + i];
      if (act == 2) {
        const float yi = y[r * nv + i];
        z[i] = gi * yi * (1.f - yi);
      } else {
        z[i] = gi;
      }
    }
  }
  uint4* row = reinterpret_cast<uint4*>(out + r * 64);
  row[0] = make_uint4(pack_bf16x2(z[0], z[1]), pack_bf16x2(z[2], z[3]), pack_bf16x2(z[4], z[5]),
                      pack_bf16x2(z[6], z[7]));
#pragma unroll
  for (int q = 1; q < 8; ++q) row[q] = make_uint4(0u, 0u, 0u, 0u);
}

static inline int ray_grid(int B, int warps) {
  long long b = ((long long)B + warps - 1) / warps;
  const long long cap = (long long)sm_count() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// the generic kernel instantiation for N intervals per ray
#define CP_GENERIC(kernel, N) ((N) <= 32 * CP_MAXC_SMALL ? kernel<CP_MAXC_SMALL> : kernel<CP_MAXC_LARGE>)

}  // namespace mip360

using namespace mip360;

extern "C" {

int mip360_composite_fwd(const float* rgb_or_raw, const float* density, const float* t_vals, const float* dirs, int B,
                         int N, int head_mode, float density_bias, float rgb_padding, int white_bkgd, float* comp_rgb,
                         float* distance, float* acc, float* weights, mip360_stream_t stream) {
  return mip360_composite_fwd_s(rgb_or_raw, density, t_vals, dirs, B, N, head_mode, density_bias, rgb_padding, white_bkgd,
                                comp_rgb, distance, acc, weights, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
}

int mip360_composite_fwd_s(const float* rgb_or_raw, const float* density, const float* t_vals, const float* dirs, int B,
                           int N, int head_mode, float density_bias, float rgb_padding, int white_bkgd, float* comp_rgb,
                           float* distance, float* acc, float* weights, const float* near, const float* far,
                           float* s_vals, float* t_shift, const float* head_bias, mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (rgb_or_raw && t_vals && dirs && comp_rgb && distance && acc), "composite_fwd: null pointer");
  MIP_REQUIRE(head_mode >= 0 && head_mode <= 2 && (head_mode != 2 || head_bias), "composite_fwd: head_mode %d (2 needs head_bias)", head_mode);
  MIP_REQUIRE(B <= 0 || !s_vals || (near && far), "composite_fwd: s_vals needs near and far");
  MIP_REQUIRE(B <= 0 || (head_mode >= 1 || density), "composite_fwd: density missing");
  MIP_REQUIRE(N >= 1 && N <= MIP360_MAX_SAMPLES, "composite_fwd: N=%d outside [1,%d]", N, MIP360_MAX_SAMPLES);
  if (B <= 0) return MIP360_OK;
  if (rg_supported_host(N))
    launch_composite_fwd<false>(N, B, (cudaStream_t)stream, rgb_or_raw, density, t_vals, dirs, B, head_mode, 0, 0, density_bias,
                         rgb_padding, white_bkgd, comp_rgb, distance, acc, weights, near, far, s_vals, t_shift, head_bias);
  else
    CP_GENERIC(composite_fwd_kernel, N)<<<ray_grid(B, CP_WARPS), CP_WARPS * 32, 0, (cudaStream_t)stream>>>(
        rgb_or_raw, density, t_vals, dirs, B, N, head_mode, 0, 0, density_bias, rgb_padding, white_bkgd, comp_rgb,
        distance, acc, weights, near, far, s_vals, t_shift, head_bias);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_composite_bwd(const float* rgb_or_raw, const float* density, const float* t_vals, const float* dirs, int B,
                         int N, int head_mode, float density_bias, float rgb_padding, int white_bkgd,
                         const float* g_rgb, const float* g_acc, const float* g_dist, const float* g_w, float* g_rgb_in,
                         float* g_density, float* g_raw, const float* head_bias, mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (rgb_or_raw && t_vals && dirs), "composite_bwd: null pointer");
  MIP_REQUIRE(head_mode >= 0 && head_mode <= 2 && (head_mode != 2 || head_bias), "composite_bwd: head_mode %d (2 needs head_bias)", head_mode);
  MIP_REQUIRE(B <= 0 || (head_mode >= 1 || density), "composite_bwd: density missing");
  MIP_REQUIRE(N >= 1 && N <= MIP360_MAX_SAMPLES, "composite_bwd: N=%d outside [1,%d]", N, MIP360_MAX_SAMPLES);
  if (B <= 0) return MIP360_OK;
  if (rg_supported_host(N))
    launch_composite_bwd<false>(N, B, (cudaStream_t)stream, rgb_or_raw, density, t_vals, dirs, B, head_mode, 0, 0, density_bias,
                         rgb_padding, white_bkgd, g_rgb, g_acc, g_dist, g_w, g_rgb_in, g_density, g_raw, head_bias);
  else
    CP_GENERIC(composite_bwd_kernel, N)<<<ray_grid(B, CP_WARPS), CP_WARPS * 32, 0, (cudaStream_t)stream>>>(
        rgb_or_raw, density, t_vals, dirs, B, N, head_mode, 0, 0, density_bias, rgb_padding, white_bkgd, g_rgb, g_acc,
        g_dist, g_w, g_rgb_in, g_density, g_raw, head_bias);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_density_to_weight_fwd(const float* density, const float* t_vals, const float* dirs, int B, int N,
                                 int density_mode, float density_bias, float* weights, mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (density && t_vals && dirs && weights), "density_to_weight_fwd: null pointer");
  MIP_REQUIRE(N >= 1 && N <= MIP360_MAX_SAMPLES, "density_to_weight_fwd: N=%d outside [1,%d]", N, MIP360_MAX_SAMPLES);
  if (B <= 0) return MIP360_OK;
  if (rg_supported_host(N))
    launch_composite_fwd<true>(N, B, (cudaStream_t)stream, (const float*)nullptr, density, t_vals, dirs, B, 0, 1, density_mode,
                         density_bias, 0.f, 0, (float*)nullptr, (float*)nullptr, (float*)nullptr, weights,
                         (const float*)nullptr, (const float*)nullptr, (float*)nullptr, (float*)nullptr, (const float*)nullptr);
  else
    CP_GENERIC(composite_fwd_kernel, N)<<<ray_grid(B, CP_WARPS), CP_WARPS * 32, 0, (cudaStream_t)stream>>>(
        nullptr, density, t_vals, dirs, B, N, 0, 1, density_mode, density_bias, 0.f, 0, nullptr, nullptr, nullptr,
        weights, nullptr, nullptr, nullptr, nullptr, nullptr);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_density_to_weight_bwd(const float* density, const float* t_vals, const float* dirs, int B, int N,
                                 int density_mode, float density_bias, const float* g_w, float* g_density,
                                 mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (density && t_vals && dirs && g_w && g_density), "density_to_weight_bwd: null pointer");
  MIP_REQUIRE(N >= 1 && N <= MIP360_MAX_SAMPLES, "density_to_weight_bwd: N=%d outside [1,%d]", N, MIP360_MAX_SAMPLES);
  if (B <= 0) return MIP360_OK;
  if (rg_supported_host(N))
    launch_composite_bwd<true>(N, B, (cudaStream_t)stream, (const float*)nullptr, density, t_vals, dirs, B, 0, 1, density_mode,
                         density_bias, 0.f, 0, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, g_w,
                         (float*)nullptr, g_density, (float*)nullptr, (const float*)nullptr);
  else
    CP_GENERIC(composite_bwd_kernel, N)<<<ray_grid(B, CP_WARPS), CP_WARPS * 32, 0, (cudaStream_t)stream>>>(
        nullptr, density, t_vals, dirs, B, N, 0, 1, density_mode, density_bias, 0.f, 0, nullptr, nullptr, nullptr, g_w,
        nullptr, g_density, nullptr, nullptr);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_t_to_s(const float* t_vals, const float* near, const float* far, int B, int K, float* s_vals,
                  float* t_shift, mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (t_vals && near && far && s_vals), "t_to_s: null pointer");
  if (B <= 0 || K <= 0) return MIP360_OK;
  const long long n = (long long)B * K;
  t_to_s_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(t_vals, near, far, B, K, s_vals, t_shift);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_s_to_t(const float* s_vals, const float* near, const float* far, int B, int K, float* t_vals,
                  mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (s_vals && near && far && t_vals), "s_to_t: null pointer");
  if (B <= 0 || K <= 0) return MIP360_OK;
  const long long n = (long long)B * K;
  s_to_t_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(s_vals, near, far, B, K, t_vals);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_head_grad_pack(const float* g, const float* y, long long M, int n_valid, int act, uint16_t* out_bf16,
                          mip360_stream_t stream) {
  MIP_REQUIRE(g && out_bf16, "head_grad_pack: null pointer");
  MIP_REQUIRE(n_valid >= 1 && n_valid <= 8, "head_grad_pack: n_valid=%d outside [1,8]", n_valid);
  MIP_REQUIRE(act == 0 || (act == 2 && y), "head_grad_pack: act=%d (0 or 2 with y)", act);
  if (M <= 0) return MIP360_OK;
  head_grad_pack_kernel<<<(int)((M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, y, M, n_valid, act, out_bf16);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

}  // extern "C"
