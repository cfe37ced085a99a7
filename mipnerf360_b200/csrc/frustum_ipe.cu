// K0/K1: level-0 sampling, conical frustum -> Gaussian -> contraction -> integrated positional
// encoding, fused.  One thread per sample; per-warp shared-memory staging turns the per-sample rows
// (3, 9, 42 floats or 64 bf16) into fully coalesced 128-byte stores.
//
// Compiled with --fmad=false: every fp32 operation is rounded the way the reference's unfused
// PyTorch elementwise ops round it (reference: intern/parameterization.py, intern/encoding.py).
#include "common.cuh"
#include "ray_group.cuh"

namespace mip360 {

// intern/encoding.py:9-29 — row order defines encoding column order
__constant__ float c_P[21][3] = {
    {0.8506508f, 0.f, 0.5257311f},   {0.809017f, 0.5f, 0.309017f},   {0.5257311f, 0.8506508f, 0.f},
    {1.f, 0.f, 0.f},                 {0.809017f, 0.5f, -0.309017f},  {0.8506508f, 0.f, -0.5257311f},
    {0.309017f, 0.809017f, -0.5f},   {0.f, 0.5257311f, -0.8506508f}, {0.5f, 0.309017f, -0.809017f},
    {0.f, 1.f, 0.f},                 {-0.5257311f, 0.8506508f, 0.f}, {-0.309017f, 0.809017f, -0.5f},
    {0.f, 0.5257311f, 0.8506508f},   {-0.309017f, 0.809017f, 0.5f},  {0.309017f, 0.809017f, 0.5f},
    {0.5f, 0.309017f, 0.809017f},    {0.5f, -0.309017f, 0.809017f},  {0.f, 0.f, 1.f},
    {-0.5f, 0.309017f, 0.809017f},   {-0.809017f, 0.5f, 0.309017f},  {-0.809017f, 0.5f, -0.309017f}};

__device__ __forceinline__ float g_disp(float x) { return 1.0f / (x + G_EPS); }  // parameterization.py:15-21

// sample index -> (ray, interval) without a 64-bit division: magic = floor(2^64 / N) + 1 (host), exact for s*N < 2^64
__device__ __forceinline__ void split_sample(long long s, int N, unsigned long long magic, int& b, int& j) {
  b = (int)__umul64hi((unsigned long long)s, magic);
  j = (int)(s - (long long)b * N);
}
static inline unsigned long long div_magic(int N) { return N > 1 ? ~0ull / (unsigned long long)N + 1ull : 0ull; }

// intern/parameterization.py:108-113: the original (numerically unstable) formula, kept for the callers that ask for it
__device__ __forceinline__ void frustum_moments_unstable(float t0, float t1, float radius, float& t_mean, float& t_var,
                                                         float& r_var) {
  const float a2 = t0 * t0, b2 = t1 * t1;
  const float a3 = a2 * t0, b3 = b2 * t1, a4 = a2 * a2, b4 = b2 * b2, a5 = a4 * t0, b5 = b4 * t1;  // torch: x**k
  t_mean = (3.f * (b4 - a4)) / (4.f * (b3 - a3));
  r_var = (radius * radius) * ((3.f / 20.f) * (b5 - a5) / (b3 - a3));
  const float t_mosq = (3.f / 5.f) * (b5 - a5) / (b3 - a3);
  t_var = t_mosq - t_mean * t_mean;
}

// intern/parameterization.py:101-107 (stable branch)
__device__ __forceinline__ void frustum_moments(float t0, float t1, float radius, float& t_mean, float& t_var,
                                                float& r_var) {
  const float mu = (t0 + t1) / 2.f;
  const float hw = (t1 - t0) / 2.f;
  const float mu2 = mu * mu, hw2 = hw * hw;
  const float hw4 = hw2 * hw2;
  const float den = 3.f * mu2 + hw2;
  t_mean = mu + (2.f * mu * hw2) / den;
  t_var = hw2 / 3.f - (4.f / 15.f) * ((hw4 * (12.f * mu2 - hw2)) / (den * den));
  r_var = (radius * radius) * (mu2 / 4.f + (5.f / 12.f) * hw2 - (4.f / 15.f) * hw4 / den);
}

// intern/parameterization.py:45-61 (diag=False)
__device__ __forceinline__ void lift_to_xyz(const float d[3], float t_mean, float t_var, float r_var, float mean[3],
                                            float cov[9]) {
  const float dmag = fmaxf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2], 1e-10f);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    mean[i] = d[i] * t_mean;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float d_outer = d[i] * d[j];
      const float null_outer = (i == j ? 1.f : 0.f) - d[i] * (d[j] / dmag);
      cov[i * 3 + j] = t_var * d_outer + r_var * null_outer;
    }
  }
}

// cov <- J cov J^T with J = a I + b x x^T (closed form of jacobian(contract, x))
__device__ __forceinline__ void apply_contract_jacobian(const float x[3], float m, float cov[9]) {
  const float m2 = m * m;
  const float a = 2.f / m - 1.f / m2;
  const float b = -2.f / (m2 * m) + 2.f / (m2 * m2);
  float J[9], T[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) J[i * 3 + j] = (i == j ? a : 0.f) + b * (x[i] * x[j]);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      T[i * 3 + j] = J[i * 3 + 0] * cov[0 * 3 + j] + J[i * 3 + 1] * cov[1 * 3 + j] + J[i * 3 + 2] * cov[2 * 3 + j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      cov[i * 3 + j] = T[i * 3 + 0] * J[j * 3 + 0] + T[i * 3 + 1] * J[j * 3 + 1] + T[i * 3 + 2] * J[j * 3 + 2];
}

// intern/parameterization.py:64-83.  mode 0: reference (global norm n, Jacobian at the scaled mean with the
// per-point norm), 1: per-point contraction, 2: none.
__device__ __forceinline__ void contract_gaussian(float mean[3], float cov[9], int mode, float n_global) {
  if (mode == 0) {
    if (!(n_global <= 1.f)) {
      const float c = 2.f - 1.f / n_global;
#pragma unroll
      for (int i = 0; i < 3; ++i) mean[i] = c * (mean[i] / n_global);
    }
    const float m = sqrtf(mean[0] * mean[0] + mean[1] * mean[1] + mean[2] * mean[2]);
    if (m > 1.f) apply_contract_jacobian(mean, m, cov);
  } else if (mode == 1) {
    const float m = sqrtf(mean[0] * mean[0] + mean[1] * mean[1] + mean[2] * mean[2]);
    if (m > 1.f) {
      apply_contract_jacobian(mean, m, cov);
      const float c = 2.f - 1.f / m;
#pragma unroll
      for (int i = 0; i < 3; ++i) mean[i] = c * (mean[i] / m);
    }
  }
}

// intern/encoding.py:43-55: gamma = P mean, sigma_k = P_k cov P_k^T; has_cov = false gives plain PE
__device__ __forceinline__ void ipe_features(const float mean[3], const float cov[9], bool has_cov, float* out_sin,
                                             float* out_cos, int stride) {
#pragma unroll
  for (int k = 0; k < 21; ++k) {
    const float p0 = c_P[k][0], p1 = c_P[k][1], p2 = c_P[k][2];
    const float gamma = p0 * mean[0] + p1 * mean[1] + p2 * mean[2];
    float damp = 1.f;
    if (has_cov) {
      // A = cov P^T (column k), sigma = sum_i P_ki A_ik
      const float a0 = cov[0] * p0 + cov[1] * p1 + cov[2] * p2;
      const float a1 = cov[3] * p0 + cov[4] * p1 + cov[5] * p2;
      const float a2 = cov[6] * p0 + cov[7] * p1 + cov[8] * p2;
      const float sigma = p0 * a0 + p1 * a1 + p2 * a2;
      damp = expf(-0.5f * sigma);
    }
    float s, c;
    sincosf(gamma, &s, &c);
    out_sin[k * stride] = damp * s;
    out_cos[k * stride] = damp * c;
  }
}

// Fast variant for the bf16 MLP rows (model path): the 21 directions are compile-time immediates (zero terms
// vanish), sigma uses the symmetric 6-term form with FMAs, exp/sin/cos are the MUFU approximations after an
// (abs. error ~1e-6, far below the bf16 rounding of the result: 2^-9 relative).
// Output: 21 packed bf16 pairs = columns 0..41 of the MLP row.
__device__ __forceinline__ void ipe_features_bf16(const float mean[3], const float cov[9], uint32_t packed[21]) {
  constexpr float A = 0.8506508f, B = 0.5257311f, C = 0.809017f, D = 0.5f, E = 0.309017f;
  constexpr float P[21][3] = {{A, 0, B},  {C, D, E},  {B, A, 0},  {1, 0, 0},   {C, D, -E},  {A, 0, -B}, {E, C, -D},
                              {0, B, -A}, {D, E, -C}, {0, 1, 0},  {-B, A, 0},  {-E, C, -D}, {0, B, A},  {-E, C, D},
                              {E, C, D},  {D, E, C},  {D, -E, C}, {0, 0, 1},   {-D, E, C},  {-C, D, E}, {-C, D, -E}};
  float sn[21], cs[21];
  const float c01 = cov[1] + cov[3], c02 = cov[2] + cov[6], c12 = cov[5] + cov[7];  // symmetric part x 2
#pragma unroll
  for (int k = 0; k < 21; ++k) {
    const float p0 = P[k][0], p1 = P[k][1], p2 = P[k][2];
    float gamma = 0.f, sigma = 0.f;
    if (p0 != 0.f) gamma = fmaf(p0, mean[0], gamma);
    if (p1 != 0.f) gamma = fmaf(p1, mean[1], gamma);
    if (p2 != 0.f) gamma = fmaf(p2, mean[2], gamma);
    if (p0 != 0.f) sigma = fmaf(p0 * p0, cov[0], sigma);
    if (p1 != 0.f) sigma = fmaf(p1 * p1, cov[4], sigma);
    if (p2 != 0.f) sigma = fmaf(p2 * p2, cov[8], sigma);
    if (p0 * p1 != 0.f) sigma = fmaf(p0 * p1, c01, sigma);
    if (p0 * p2 != 0.f) sigma = fmaf(p0 * p2, c02, sigma);
    if (p1 * p2 != 0.f) sigma = fmaf(p1 * p2, c12, sigma);
    const float damp = exp2f(-0.72134752f * sigma);  // exp(-sigma/2) = 2^(-sigma/(2 ln 2)): one MUFU.EX2
    // MUFU.SIN/COS reduce the argument themselves; |gamma| stays O(scene radius), where their absolute error
    // (~|gamma| * 2^-22) is far below the bf16 rounding of the product
    sn[k] = damp * __sinf(gamma);
    cs[k] = damp * __cosf(gamma);
  }
#pragma unroll
  for (int c = 0; c < 10; ++c) packed[c] = pack_bf16x2(sn[2 * c], sn[2 * c + 1]);
  packed[10] = pack_bf16x2(sn[20], cs[0]);
#pragma unroll
  for (int c = 0; c < 10; ++c) packed[11 + c] = pack_bf16x2(cs[2 * c + 1], cs[2 * c + 2]);
}

constexpr int K1_THREADS = 256;
constexpr int K1_STAGE_FLOATS = 32 * 43;  // per-warp staging: 32 rows x (42 + 1 pad) floats

// Write V floats per lane (row-major rows of length V, rows of a warp contiguous in global memory)
template <int V>
__device__ __forceinline__ void warp_store_rows(float* stage, const float* vals, float* gbase, int lane,
                                                int rows_valid) {
  constexpr int LD = (V % 2 == 0) ? V + 1 : V;  // odd leading dimension -> conflict-free
#pragma unroll
  for (int c = 0; c < V; ++c) stage[lane * LD + c] = vals[c];
  __syncwarp();
  const int total = rows_valid * V;
  for (int e = lane; e < total; e += 32) {
    const int r = e / V, c = e - r * V;
    gbase[e] = stage[r * LD + c];
  }
  __syncwarp();
}

// WIDE: bf16 rows of 128 columns (more than 22 view-direction features); the default is 64
template <bool FAST, bool WIDE = false>
__global__ void __launch_bounds__(K1_THREADS)
cast_ipe_kernel(const float* __restrict__ t0p, const float* __restrict__ t1p, int t_stride,
                const float* __restrict__ origins, const float* __restrict__ directions,
                const float* __restrict__ vdir_enc, const float* __restrict__ radii,
                const double* __restrict__ norm_sq, long long S, int N, unsigned long long magic, int contract_mode,
                int add_origins,
                float* __restrict__ means_out, float* __restrict__ covs_out, float* __restrict__ enc_out,
                uint16_t* __restrict__ x_out, int vd_dim, int x_cols) {
  __shared__ __align__(16) float stage_all[(K1_THREADS / 32) * K1_STAGE_FLOATS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* stage = stage_all + warp * K1_STAGE_FLOATS;
  const long long warp_s0 = (long long)blockIdx.x * K1_THREADS + warp * 32;
  if (warp_s0 >= S) return;
  const long long s = warp_s0 + lane;
  const bool valid = s < S;
  const int rows_valid = (int)min((long long)32, S - warp_s0);
  const long long sc = valid ? s : S - 1;
  int b, j;
  if (N > 1) split_sample(sc, N, magic, b, j); else { b = (int)sc; j = 0; }

  float n_global = 0.f;
  if (contract_mode == 0) n_global = (float)sqrt(*norm_sq);

  const float t0 = t0p[(long long)b * t_stride + j];
  const float t1 = t1p[(long long)b * t_stride + j];
  float d[3] = {directions[b * 3 + 0], directions[b * 3 + 1], directions[b * 3 + 2]};
  const float radius = radii[b];

  float t_mean, t_var, r_var, mean[3], cov[9];
  if (add_origins & 2) frustum_moments_unstable(t0, t1, radius, t_mean, t_var, r_var);
  else frustum_moments(t0, t1, radius, t_mean, t_var, r_var);
  lift_to_xyz(d, t_mean, t_var, r_var, mean, cov);
  contract_gaussian(mean, cov, contract_mode, n_global);
  if (add_origins & 1) {
#pragma unroll
    for (int i = 0; i < 3; ++i) mean[i] = mean[i] + origins[b * 3 + i];
  }

  uint32_t packed[32];
  if (FAST) {
    // model path: only the bf16 MLP rows are wanted
    ipe_features_bf16(mean, cov, packed);
  } else {
    if (means_out) warp_store_rows<3>(stage, mean, means_out + warp_s0 * 3, lane, rows_valid);
    if (covs_out) warp_store_rows<9>(stage, cov, covs_out + warp_s0 * 9, lane, rows_valid);
    if (!enc_out && !x_out) return;

    // features into the staging tile: row = lane, leading dimension 43
    float* row = stage + lane * 43;
    ipe_features(mean, cov, true, row, row + 21, 1);
    __syncwarp();
    if (enc_out) {
      float* g = enc_out + warp_s0 * 42;
      const int total = rows_valid * 42;
      for (int e = lane; e < total; e += 32) {
        const int r = e / 42, c = e - r * 42;
        g[e] = stage[r * 43 + c];
      }
    }
    if (x_out) {
#pragma unroll
      for (int c = 0; c < 21; ++c) packed[c] = pack_bf16x2(row[2 * c], row[2 * c + 1]);
    }
  }
  if (x_out) {
    // bf16 rows of x_cols (64, or 128 when more than 22 view-direction features are configured): [0,42) IPE,
    // [42,42+vd_dim) view-direction encoding of the ray, zeros up to x_cols.  Per 64-column half: each lane
    // converts its own half-row to 8 x 16-byte chunks, stored XOR-swizzled, then the warp writes the 32 half-rows
    // (128 B each, whole lines) with 16-byte coalesced stores.
    const float* vd = vdir_enc + (long long)b * vd_dim;
    auto vdv = [&](int k) { return k < vd_dim ? __ldg(vd + k) : 0.f; };
    uint4* stage4 = reinterpret_cast<uint4*>(stage);
    constexpr int chunks_per_row = WIDE ? 16 : 8;  // 16-byte chunks per bf16 row
#pragma unroll
    for (int h = 0; h < (WIDE ? 2 : 1); ++h) {
      if (h == 0) {
        if (vd_dim == 16) {  // the default (viewdir degrees 0..4): four 16-byte loads
          const float4* vd4 = reinterpret_cast<const float4*>(vd);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 v = vd4[q];
            packed[21 + 2 * q] = pack_bf16x2(v.x, v.y);
            packed[22 + 2 * q] = pack_bf16x2(v.z, v.w);
          }
          packed[29] = packed[30] = packed[31] = 0u;
        } else {
#pragma unroll
          for (int p = 21; p < 32; ++p) packed[p] = pack_bf16x2(vdv(2 * (p - 21)), vdv(2 * (p - 21) + 1));
        }
      } else {
#pragma unroll
        for (int p = 0; p < 32; ++p) packed[p] = pack_bf16x2(vdv(22 + 64 * (h - 1) + 2 * p), vdv(23 + 64 * (h - 1) + 2 * p));
      }
      __syncwarp();  // everyone is done reading the staging rows (fp32 features / the previous half)
#pragma unroll
      for (int c = 0; c < 8; ++c)
        stage4[lane * 8 + (c ^ (lane & 7))] =
            make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
      __syncwarp();
      uint4* g4 = reinterpret_cast<uint4*>(x_out) + warp_s0 * chunks_per_row + 8 * h;
      const int total = rows_valid * 8;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int q = it * 32 + lane;
        if (q < total) {
          const int r = q >> 3, c = q & 7;
          g4[(long long)r * chunks_per_row + c] = stage4[r * 8 + (c ^ (r & 7))];
        }
      }
    }
  }
}

// Σ |d t_mean|^2 over all samples (App. A1): per-thread fp32 products as the reference forms them,
// accumulated in fp64.
__global__ void __launch_bounds__(256)
frustum_norm_sq_kernel(const float* __restrict__ t0p, const float* __restrict__ t1p, int t_stride,
                       const float* __restrict__ directions, long long S, int N, unsigned long long magic,
                       double* __restrict__ out) {
  double acc = 0.0;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (long long)gridDim.x * blockDim.x) {
    int b, j;
    if (N > 1) split_sample(s, N, magic, b, j); else { b = (int)s; j = 0; }
    const float t0 = t0p[(long long)b * t_stride + j], t1 = t1p[(long long)b * t_stride + j];
    // t_mean of parameterization.py:103 only (t_var / r_var are not needed here)
    const float mu = (t0 + t1) / 2.f, hw = (t1 - t0) / 2.f;
    const float hw2 = hw * hw;
    const float t_mean = mu + (2.f * mu * hw2) / (3.f * (mu * mu) + hw2);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float m = directions[b * 3 + i] * t_mean;
      acc += (double)(m * m);
    }
  }
  acc = warp_sum(acc);
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += sm[i];
    atomicAdd(out, s);
  }
}

// The same sum for knot rows [B, N+1] with N in {32, 64, 128}: 8 lanes per ray, E = N/8 intervals per lane, so a
// knot is fetched once per lane instead of twice per sample, the direction once per lane, and one fp64 add per
// sample (the three fp32 squares of a sample are summed in fp32 first) replaces three.  (A cyclic knot assignment with
// whole-sector loads and neighbour shuffles was measured slower, 0.41 vs 0.34 ms for 4M rays: the kernel is bound by
// instruction issue — one IEEE division per interval — not by L1 wavefronts.)
template <int E>
__global__ void __launch_bounds__(RG_THREADS)
frustum_norm_sq_rg_kernel(const float* __restrict__ t_vals, const float* __restrict__ directions, int B,
                          double* __restrict__ out) {
  constexpr int N = E * RG_LANES;
  const int gl = threadIdx.x & 7;
  double acc = 0.0;
  for (long long base = (long long)blockIdx.x * RG_RAYS_PER_BLOCK; base < B; base += (long long)gridDim.x * RG_RAYS_PER_BLOCK) {
    const long long ray = base + (threadIdx.x >> 3);
    if (ray >= B) continue;
    float t[E + 1];
    rg_load_knots<E>(t_vals + ray * (N + 1), gl * E, t);
    const float d0 = __ldg(directions + ray * 3), d1 = __ldg(directions + ray * 3 + 1), d2 = __ldg(directions + ray * 3 + 2);
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const float mu = (t[i] + t[i + 1]) / 2.f, hw = (t[i + 1] - t[i]) / 2.f;
      const float hw2 = hw * hw;
      const float t_mean = mu + (2.f * mu * hw2) / (3.f * (mu * mu) + hw2);
      const float m0 = d0 * t_mean, m1 = d1 * t_mean, m2 = d2 * t_mean;
      acc += (double)(m0 * m0 + m1 * m1 + m2 * m2);
    }
  }
  acc = warp_sum(acc);
  __shared__ double sm[RG_THREADS / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < RG_THREADS / 32; ++i) s += sm[i];
    atomicAdd(out, s);
  }
}

__global__ void __launch_bounds__(256)
sum_sq_kernel(const float* __restrict__ x, long long n, double* __restrict__ out) {
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = x[i];
    acc += v * v;
  }
  acc = warp_sum(acc);
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += sm[i];
    atomicAdd(out, s);
  }
}

__global__ void __launch_bounds__(256)
contract_kernel(const float* __restrict__ x, long long n, const double* __restrict__ norm_sq, float* __restrict__ y) {
  const float nrm = (float)sqrt(*norm_sq);
  const bool ident = nrm <= 1.f;
  const float c = 2.f - 1.f / nrm;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = ident ? x[i] : c * (x[i] / nrm);
}

__global__ void __launch_bounds__(256)
gaussian_contract_kernel(const float* __restrict__ mi, const float* __restrict__ ci, const double* __restrict__ norm_sq,
                         long long S, float* __restrict__ mo, float* __restrict__ co) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const float nrm = (float)sqrt(*norm_sq);
  float mean[3], cov[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) mean[i] = mi[s * 3 + i];
#pragma unroll
  for (int i = 0; i < 9; ++i) cov[i] = ci[s * 9 + i];
  contract_gaussian(mean, cov, 0, nrm);
#pragma unroll
  for (int i = 0; i < 3; ++i) mo[s * 3 + i] = mean[i];
#pragma unroll
  for (int i = 0; i < 9; ++i) co[s * 9 + i] = cov[i];
}

__global__ void __launch_bounds__(256)
gaussian_to_xyz_kernel(const float* __restrict__ directions, const float* __restrict__ t_mean,
                       const float* __restrict__ t_var, const float* __restrict__ r_var, long long S, int N,
                       float* __restrict__ means, float* __restrict__ covs) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int b = (int)(s / N);
  float d[3] = {directions[b * 3], directions[b * 3 + 1], directions[b * 3 + 2]};
  float mean[3], cov[9];
  lift_to_xyz(d, t_mean[s], t_var[s], r_var[s], mean, cov);
#pragma unroll
  for (int i = 0; i < 3; ++i) means[s * 3 + i] = mean[i];
#pragma unroll
  for (int i = 0; i < 9; ++i) covs[s * 9 + i] = cov[i];
}

// intern/parameterization.py:49-53 (diag=True): the diagonal of the covariance only
__global__ void __launch_bounds__(256)
gaussian_to_xyz_diag_kernel(const float* __restrict__ directions, const float* __restrict__ t_mean,
                            const float* __restrict__ t_var, const float* __restrict__ r_var, long long S, int N,
                            float* __restrict__ means, float* __restrict__ cov_diag) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int b = (int)(s / N);
  const float d[3] = {directions[b * 3], directions[b * 3 + 1], directions[b * 3 + 2]};
  const float dmag = fmaxf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2], 1e-10f);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float dd = d[i] * d[i];
    means[s * 3 + i] = d[i] * t_mean[s];
    cov_diag[s * 3 + i] = t_var[s] * dd + r_var[s] * (1.f - dd / dmag);
  }
}

__global__ void __launch_bounds__(256)
ipe_kernel(const float* __restrict__ means, const float* __restrict__ covs, long long S, float* __restrict__ enc) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float mean[3], cov[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) mean[i] = means[s * 3 + i];
  if (covs) {
#pragma unroll
    for (int i = 0; i < 9; ++i) cov[i] = covs[s * 9 + i];
  }
  float out[42];
  ipe_features(mean, cov, covs != nullptr, out, out + 21, 1);
#pragma unroll
  for (int i = 0; i < 42; ++i) enc[s * 42 + i] = out[i];
}

// intern/encoding.py:79-89
__global__ void __launch_bounds__(256)
viewdir_enc_kernel(const float* __restrict__ viewdirs, int B, int min_deg, int max_deg, float* __restrict__ enc) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float x = viewdirs[b * 3], y = viewdirs[b * 3 + 1], z = viewdirs[b * 3 + 2];
  const float theta = acosf(z);
  const float phi = atanf(y / (x + 1e-6f));
  const int ns = max_deg - min_deg;
  float* o = enc + (long long)b * 4 * ns;
  for (int k = 0; k < ns; ++k) {
    const float sc = exp2f((float)(min_deg + k));
    float s, c;
    sincosf(sc * theta, &s, &c);
    o[k] = s;
    o[ns + k] = c;
    sincosf(sc * phi, &s, &c);
    o[2 * ns + k] = s;
    o[3 * ns + k] = c;
  }
}

// intern/ray.py:100-111.  The stratified-jitter uniforms come from t_rand (given draw), else from the in-kernel
// generator (rng.enabled), else the knots are deterministic.  With norm_sq != null the kernel also accumulates the
// squared Frobenius norm the reference's contract() sees for these knots (parameterization.py:25,75; App. A1), so the
// separate pre-pass over t_vals disappears: the thread of knot i owns interval [t_i, t_{i+1}].
__global__ void __launch_bounds__(256)
level0_t_kernel(const float* __restrict__ near, const float* __restrict__ far, const float* __restrict__ s_lin,
                const float* __restrict__ t_rand, RngArgs rng, const float* __restrict__ directions,
                double* __restrict__ norm_sq, float* __restrict__ t_out, int B, int N) {
  const int K = N + 1;
  const long long total = (long long)B * K;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = e < total;
  double contrib = 0.0;
  if (valid) {
    const int b = (int)(e / K), i = (int)(e - (long long)b * K);
    const float gf = g_disp(far[b]), gn = g_disp(near[b]);
    auto tv = [&](int k) {
      const float s = s_lin[k];
      return g_disp(s * gf + (1.f - s) * gn);
    };
    const bool randomized = t_rand != nullptr || rng.enabled;
    const uint32_t epoch = rng.enabled ? rng_epoch(rng) : 0u;
    // knot k of this ray with its jitter: mids = .5*(t[1:]+t[:-1]); upper = [mids, t[-1]]; lower = [t[0], mids]
    auto knot = [&](int k, float tk, float t_prev, float t_next) {
      if (!randomized) return tk;
      const float upper = (k < N) ? 0.5f * (t_next + tk) : tk;
      const float lower = (k > 0) ? 0.5f * (tk + t_prev) : tk;
      const float u = t_rand ? t_rand[(long long)b * K + k] : rng_uniform(rng, epoch, (uint32_t)b, (uint32_t)k);
      return lower + (upper - lower) * u;
    };
    const float t_im1 = i > 0 ? tv(i - 1) : 0.f, t_i = tv(i), t_ip1 = i < N ? tv(i + 1) : 0.f;
    const float mine = knot(i, t_i, t_im1, t_ip1);
    t_out[e] = mine;
    if (norm_sq && i < N) {
      const float t_ip2 = i + 1 < N ? tv(i + 2) : 0.f;
      const float next = knot(i + 1, t_ip1, t_i, t_ip2);
      // t_mean of parameterization.py:103 and |d t_mean|^2 formed like frustum_norm_sq_rg_kernel
      const float mu = (mine + next) / 2.f, hw = (next - mine) / 2.f;
      const float hw2 = hw * hw;
      const float t_mean = mu + (2.f * mu * hw2) / (3.f * (mu * mu) + hw2);
      const float m0 = directions[b * 3] * t_mean, m1 = directions[b * 3 + 1] * t_mean, m2 = directions[b * 3 + 2] * t_mean;
      contrib = (double)(m0 * m0 + m1 * m1 + m2 * m2);
    }
  }
  if (norm_sq) {
    contrib = warp_sum(contrib);
    __shared__ double sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = contrib;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < 8; ++w) s += sm[w];
      atomicAdd(norm_sq, s);
    }
  }
}

// The same for N in {32, 64, 128} with 8 lanes per ray (ray_group.cuh): lane gl owns knots gl + 8c, so the group writes 8
// consecutive floats per step, the deterministic knots of the neighbours i - 1 / i + 1 and the jittered knot i + 1 come from
// the adjacent lanes by shuffle instead of being recomputed, and the lane's E + 1 uniforms cost E / 4 + 1 Philox calls.
template <int E>
__global__ void __launch_bounds__(RG_THREADS)
level0_t_rg_kernel(const float* __restrict__ near, const float* __restrict__ far, const float* __restrict__ s_lin,
                   const float* __restrict__ t_rand, RngArgs rng, const float* __restrict__ directions,
                   double* __restrict__ norm_sq, float* __restrict__ t_out, int B) {
  constexpr int N = E * RG_LANES, K = N + 1, S = E + 1;
  const int gl = threadIdx.x & 7;
  const bool randomized = t_rand != nullptr || rng.enabled;
  const uint32_t epoch = rng.enabled ? rng_epoch(rng) : 0u;
  float sl[S];
#pragma unroll
  for (int c = 0; c < S; ++c) sl[c] = __ldg(s_lin + ((c < E) ? gl + RG_LANES * c : N));
  double acc = 0.0;
  for (long long base = (long long)blockIdx.x * RG_RAYS_PER_BLOCK; base < B; base += (long long)gridDim.x * RG_RAYS_PER_BLOCK) {
    const long long ray_raw = base + (threadIdx.x >> 3);
    const bool active = ray_raw < B;
    const long long ray = active ? ray_raw : B - 1;
    const float gf = g_disp(__ldg(far + ray)), gn = g_disp(__ldg(near + ray));
    float u[S];
    if (t_rand) {
#pragma unroll
      for (int c = 0; c < S; ++c) u[c] = __ldg(t_rand + ray * K + ((c < E) ? gl + RG_LANES * c : N));
    } else if (rng.enabled) {
      rg_draw<E>(rng, epoch, (uint32_t)ray, gl, u);
    }
    float tv[S];  // deterministic knots: slot c = knot gl + 8c, slot E = knot N (every lane computes it; lane 0 uses it)
#pragma unroll
    for (int c = 0; c < S; ++c) tv[c] = g_disp(sl[c] * gf + (1.f - sl[c]) * gn);
    float t[S];
#pragma unroll
    for (int c = 0; c < S; ++c) {
      float v = tv[c];
      if (randomized) {
        // knot i - 1: the previous lane's knot of this slot, or lane 7's knot of the previous slot; knot i + 1: the next
        // lane's, or lane 0's next slot (slot E = knot N for c = E - 1)
        const float up = __shfl_up_sync(FULL_MASK, tv[c], 1, RG_LANES);
        const float l7 = __shfl_sync(FULL_MASK, tv[c > 0 ? c - 1 : 0], RG_LANES - 1, RG_LANES);
        const float dn = __shfl_down_sync(FULL_MASK, tv[c], 1, RG_LANES);
        const float l0 = __shfl_sync(FULL_MASK, tv[c < E ? c + 1 : E], 0, RG_LANES);
        const int k = (c < E) ? gl + RG_LANES * c : N;
        const float t_prev = gl == 0 ? l7 : up, t_next = gl == RG_LANES - 1 ? l0 : dn;
        // slot E (knot N, valid in lane 0): its predecessor is knot N - 1 = lane 7's slot E - 1
        const float prev = (c == E) ? __shfl_sync(FULL_MASK, tv[E - 1], RG_LANES - 1, RG_LANES) : t_prev;
        const float upper = (k < N) ? 0.5f * (t_next + v) : v;
        const float lower = (k > 0) ? 0.5f * (v + prev) : v;
        v = lower + (upper - lower) * u[c];
      }
      t[c] = v;
    }
    if (active) {
#pragma unroll
      for (int c = 0; c < E; ++c) t_out[ray * K + gl + RG_LANES * c] = t[c];
      if (gl == 0) t_out[ray * K + N] = t[E];
    }
    if (norm_sq) {
      const float d0 = __ldg(directions + ray * 3), d1 = __ldg(directions + ray * 3 + 1), d2 = __ldg(directions + ray * 3 + 2);
#pragma unroll
      for (int c = 0; c < E; ++c) {
        const float dn = __shfl_down_sync(FULL_MASK, t[c], 1, RG_LANES);
        const float l0 = __shfl_sync(FULL_MASK, t[c + 1], 0, RG_LANES);
        const float t0 = t[c], t1 = gl == RG_LANES - 1 ? l0 : dn;
        const float mu = (t0 + t1) / 2.f, hw = (t1 - t0) / 2.f;
        const float hw2 = hw * hw;
        const float t_mean = mu + (2.f * mu * hw2) / (3.f * (mu * mu) + hw2);
        const float m0 = d0 * t_mean, m1 = d1 * t_mean, m2 = d2 * t_mean;
        if (active) acc += (double)(m0 * m0 + m1 * m1 + m2 * m2);
      }
    }
  }
  if (norm_sq) {
    acc = warp_sum(acc);
    __shared__ double sm[RG_THREADS / 32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int i = 0; i < RG_THREADS / 32; ++i) s += sm[i];
      atomicAdd(norm_sq, s);
    }
  }
}

static inline int blocks_for(long long n, int threads) { return (int)((n + threads - 1) / threads); }
static inline int capped_blocks(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace mip360

using namespace mip360;

extern "C" {

int mip360_level0_t_vals(const float* near, const float* far, const float* s_lin, const float* t_rand, float* t_out,
                         int B, int N, mip360_stream_t stream) {
  return mip360_level0_sample(near, far, s_lin, t_rand, 0, 0ull, 0u, nullptr, nullptr, nullptr, t_out, B, N, stream);
}

int mip360_level0_sample(const float* near, const float* far, const float* s_lin, const float* t_rand, int use_rng,
                         unsigned long long rng_seed, unsigned int rng_stream, const unsigned long long* rng_epoch,
                         const float* directions, double* norm_sq, float* t_out, int B, int N, mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (near && far && s_lin && t_out), "level0_sample: null pointer");
  MIP_REQUIRE(B >= 0 && N >= 1, "level0_sample: bad sizes B=%d N=%d", B, N);
  MIP_REQUIRE(!norm_sq || directions, "level0_sample: the norm needs the ray directions");
  MIP_REQUIRE(!(t_rand && use_rng), "level0_sample: either a given draw or the in-kernel generator");
  if (B == 0) return MIP360_OK;
  const long long total = (long long)B * (N + 1);
  RngArgs rng{rng_seed, rng_epoch, rng_stream, use_rng ? 1 : 0};
  cudaStream_t st = (cudaStream_t)stream;
  if (rg_supported_host(N) && N == 32)
    level0_t_rg_kernel<4><<<rg_grid(B), RG_THREADS, 0, st>>>(near, far, s_lin, t_rand, rng, directions, norm_sq, t_out, B);
  else if (rg_supported_host(N) && N == 64)
    level0_t_rg_kernel<8><<<rg_grid(B), RG_THREADS, 0, st>>>(near, far, s_lin, t_rand, rng, directions, norm_sq, t_out, B);
  else if (rg_supported_host(N) && N == 128)
    level0_t_rg_kernel<16><<<rg_grid(B), RG_THREADS, 0, st>>>(near, far, s_lin, t_rand, rng, directions, norm_sq, t_out, B);
  else
    level0_t_kernel<<<blocks_for(total, 256), 256, 0, st>>>(near, far, s_lin, t_rand, rng, directions, norm_sq, t_out, B, N);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_frustum_norm_sq(const float* t0, const float* t1, int t_stride, const float* directions, int B, int N,
                           double* norm_sq, mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (t0 && t1 && directions && norm_sq), "frustum_norm_sq: null pointer");
  MIP_REQUIRE(B >= 0 && N >= 1 && t_stride >= N, "frustum_norm_sq: bad sizes");
  if (B == 0) return MIP360_OK;
  const long long S = (long long)B * N;
  cudaStream_t st = (cudaStream_t)stream;
  if (rg_supported_host(N) && t1 == t0 + 1 && t_stride == N + 1) {  // adjacent knots of one [B, N+1] array
    if (N == 32) frustum_norm_sq_rg_kernel<4><<<rg_grid(B), RG_THREADS, 0, st>>>(t0, directions, B, norm_sq);
    else if (N == 64) frustum_norm_sq_rg_kernel<8><<<rg_grid(B), RG_THREADS, 0, st>>>(t0, directions, B, norm_sq);
    else frustum_norm_sq_rg_kernel<16><<<rg_grid(B), RG_THREADS, 0, st>>>(t0, directions, B, norm_sq);
  } else {
    frustum_norm_sq_kernel<<<capped_blocks(S, 256), 256, 0, st>>>(t0, t1, t_stride, directions, S, N, div_magic(N), norm_sq);
  }
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_cast_ipe(const float* t0, const float* t1, int t_stride, const float* origins, const float* directions,
                    const float* vdir_enc, const float* radii, const double* norm_sq, int B, int N, int contract_mode,
                    int add_origins, float* means, float* covs, float* enc, uint16_t* x_bf16, mip360_stream_t stream) {
  return mip360_cast_ipe_x(t0, t1, t_stride, origins, directions, vdir_enc, MIP360_VDIR_DIM, radii, norm_sq, B, N,
                           contract_mode, add_origins, means, covs, enc, x_bf16, MIP360_MLP_IN_PAD, stream);
}

int mip360_cast_ipe_x(const float* t0, const float* t1, int t_stride, const float* origins, const float* directions,
                      const float* vdir_enc, int vd_dim, const float* radii, const double* norm_sq, int B, int N,
                      int contract_mode, int add_origins, float* means, float* covs, float* enc, uint16_t* x_bf16,
                      int x_cols, mip360_stream_t stream) {
  MIP_REQUIRE(!x_bf16 || (vd_dim >= 0 && vd_dim % 4 == 0 && (x_cols == 64 || x_cols == 128) && 42 + vd_dim <= x_cols),
              "cast_ipe: %d view-direction features do not fit bf16 rows of %d columns (64 or 128)", vd_dim, x_cols);
  MIP_REQUIRE(B <= 0 || (t0 && t1 && directions && radii), "cast_ipe: null pointer");
  MIP_REQUIRE(B >= 0 && N >= 1 && t_stride >= N, "cast_ipe: bad sizes B=%d N=%d stride=%d", B, N, t_stride);
  MIP_REQUIRE(contract_mode >= 0 && contract_mode <= 2, "cast_ipe: contract_mode %d", contract_mode);
  MIP_REQUIRE(contract_mode != 0 || norm_sq, "cast_ipe: reference contraction needs norm_sq");
  MIP_REQUIRE(!(add_origins & 1) || origins, "cast_ipe: add_origins without origins");
  MIP_REQUIRE(!x_bf16 || vdir_enc || vd_dim == 0, "cast_ipe: x_bf16 output needs vdir_enc [B,vd_dim]");
  if (B == 0) return MIP360_OK;
  const long long S = (long long)B * N;
  if (x_bf16 && x_cols == 128)
    cast_ipe_kernel<false, true><<<blocks_for(S, K1_THREADS), K1_THREADS, 0, (cudaStream_t)stream>>>(
        t0, t1, t_stride, origins, directions, vdir_enc, radii, norm_sq, S, N, div_magic(N), contract_mode, add_origins,
        means, covs, enc, x_bf16, vd_dim, x_cols);
  else if (x_bf16 && !means && !covs && !enc)
    cast_ipe_kernel<true><<<blocks_for(S, K1_THREADS), K1_THREADS, 0, (cudaStream_t)stream>>>(
        t0, t1, t_stride, origins, directions, vdir_enc, radii, norm_sq, S, N, div_magic(N), contract_mode, add_origins,
        means, covs, enc, x_bf16, vd_dim, x_cols);
  else
    cast_ipe_kernel<false><<<blocks_for(S, K1_THREADS), K1_THREADS, 0, (cudaStream_t)stream>>>(
        t0, t1, t_stride, origins, directions, vdir_enc, radii, norm_sq, S, N, div_magic(N), contract_mode, add_origins,
        means, covs, enc, x_bf16, vd_dim, x_cols);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_gaussian_to_xyz(const float* directions, const float* t_mean, const float* t_var, const float* r_var, int B,
                           int N, float* means, float* covs, mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (directions && t_mean && t_var && r_var && means && covs), "gaussian_to_xyz: null pointer");
  if (B <= 0) return MIP360_OK;
  const long long S = (long long)B * N;
  gaussian_to_xyz_kernel<<<blocks_for(S, 256), 256, 0, (cudaStream_t)stream>>>(directions, t_mean, t_var, r_var, S, N,
                                                                               means, covs);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_gaussian_to_xyz_diag(const float* directions, const float* t_mean, const float* t_var, const float* r_var,
                                int B, int N, float* means, float* cov_diag, mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (directions && t_mean && t_var && r_var && means && cov_diag), "gaussian_to_xyz_diag: null pointer");
  if (B <= 0) return MIP360_OK;
  const long long S = (long long)B * N;
  gaussian_to_xyz_diag_kernel<<<blocks_for(S, 256), 256, 0, (cudaStream_t)stream>>>(directions, t_mean, t_var, r_var, S,
                                                                                    N, means, cov_diag);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_sum_sq(const float* x, long long n, double* out, mip360_stream_t stream) {
  MIP_REQUIRE(x && out, "sum_sq: null pointer");
  if (n <= 0) return MIP360_OK;
  sum_sq_kernel<<<capped_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, out);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_contract(const float* x, long long n, const double* norm_sq, float* y, mip360_stream_t stream) {
  MIP_REQUIRE(x && y && norm_sq, "contract: null pointer");
  if (n <= 0) return MIP360_OK;
  contract_kernel<<<capped_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, norm_sq, y);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_gaussian_contract(const float* means_in, const float* covs_in, const double* norm_sq, long long S,
                             float* means_out, float* covs_out, mip360_stream_t stream) {
  MIP_REQUIRE(means_in && covs_in && norm_sq && means_out && covs_out, "gaussian_contract: null pointer");
  if (S <= 0) return MIP360_OK;
  gaussian_contract_kernel<<<blocks_for(S, 256), 256, 0, (cudaStream_t)stream>>>(means_in, covs_in, norm_sq, S,
                                                                                 means_out, covs_out);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_ipe(const float* means, const float* covs, long long S, float* enc, mip360_stream_t stream) {
  MIP_REQUIRE(means && enc, "ipe: null pointer");
  if (S <= 0) return MIP360_OK;
  ipe_kernel<<<blocks_for(S, 256), 256, 0, (cudaStream_t)stream>>>(means, covs, S, enc);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

int mip360_viewdir_enc(const float* viewdirs, int B, int min_deg, int max_deg, float* enc, mip360_stream_t stream) {
  MIP_REQUIRE(B <= 0 || (viewdirs && enc), "viewdir_enc: null pointer");
  MIP_REQUIRE(max_deg > min_deg, "viewdir_enc: empty scale range");
  if (B <= 0) return MIP360_OK;
  viewdir_enc_kernel<<<blocks_for(B, 256), 256, 0, (cudaStream_t)stream>>>(viewdirs, B, min_deg, max_deg, enc);
  MIP_LAUNCH_CHECK();
  return MIP360_OK;
}

}  // extern "C"
