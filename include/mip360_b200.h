/*
 * mip360_b200.h — C ABI of the B200-native per-ray hot path of zhangkai0425/mipnerf360.
 *
 * The reference (/root/reference) is 100 % Python and defines no FFI; its hot path sits behind
 * Python call signatures.  This header is the boundary a reference-side binding
 * (ctypes) attaches to.  Each entry point names the reference function
 * (file:line, relative to the reference root) whose arithmetic it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (sm_100a) unless the name ends in _host;
 *   - tensors are dense row-major fp32 unless stated; "bf16" buffers are uint16_t bit patterns;
 *   - B = rays, N = samples (intervals) per ray, S = B*N samples; t-vectors have N+1 knots;
 *   - functions only enqueue work on `stream` (a cudaStream_t passed as void*); they never
 *     allocate, never synchronise and never throw.  Return 0 on success or a negative
 *     MIP360_ERR_* code; mip360_last_error() returns a static message for the calling thread;
 *   - inputs are never modified (the reference mutates several of its inputs).
 */
#ifndef MIP360_B200_H
#define MIP360_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIP360_OK 0
#define MIP360_ERR_ARG (-1)         /* null pointer, bad size, unsupported N */
#define MIP360_ERR_CUDA (-2)        /* a CUDA runtime/driver call failed */
#define MIP360_ERR_UNSUPPORTED (-3) /* shape not supported by the tcgen05 path */

#define MIP360_MAX_SAMPLES 512 /* one warp holds a ray: <= 4 intervals per lane up to N = 128, <= 16 up to N = 512 */
#define MIP360_ENC_DIM 42      /* 21 directions x {sin, cos}; intern/encoding.py:9-30 */
#define MIP360_VDIR_DIM 16     /* 4 scales x {sin,cos} x {theta,phi}; intern/encoding.py:67 */
#define MIP360_MLP_IN 58       /* model.py:39,127 */
#define MIP360_MLP_IN_PAD 64   /* bf16 row of the MLP input: 58 features + 6 zero columns = 128 B */

typedef void* mip360_stream_t;

const char* mip360_last_error(void);
int mip360_version(void);
/* number of kernel launches issued through this library since load / since the last reset */
long long mip360_launch_count(void);
void mip360_reset_launch_count(void);
/* Runtime switch between kernel variants that compute the same function (used by the tests to compare them):
 * key 0 = 8-lanes-per-ray register kernels for N in {32,64,128} (else one warp per ray),
 * key 1 = CTA-pair (cta_group::2) GEMM tiles, key 2 = short-K two-CTAs-per-SM GEMM configuration,
 * key 3 = packed (fp32x2 / bf16x2) arithmetic in the ReLU and trunk-Sigmoid GEMM epilogues, key 4 = the layer-fused
 * forward of the proposal MLP (mip360_mlp_fwd_fused_narrow; off = it reports MIP360_ERR_UNSUPPORTED and the caller runs
 * the layers one by one).  All default on. */
int mip360_set_option(int key, int value);

/* ------------------------------------------------------------------------------------------
 * K0  level-0 sampling               intern/ray.py:100-111, intern/parameterization.py:15-21
 *   t = g(s*g(far) + (1-s)*g(near)), g(x) = 1/(x+1e-6); s_lin[N+1] = linspace(0,1,N+1) is
 *   supplied by the host binding so that it is the same fp32 vector the reference uses.
 *   t_rand [B,N+1] (the torch.rand draw of ray.py:106) or NULL for the deterministic mode.
 * ------------------------------------------------------------------------------------------ */
int mip360_level0_t_vals(const float* near, const float* far, const float* s_lin, const float* t_rand,
                         float* t_out, int B, int N, mip360_stream_t stream);
/* The same with (a) the uniforms of ray.py:106 drawn inside the kernel when use_rng = 1 (Philox4x32-10 keyed by rng_seed;
 * number i of ray b: counter = (b, (i & 7) | ((i >> 5) << 3), rng_stream, *rng_epoch), output word (i >> 3) & 3,
 * u = (word >> 8) * 2^-24; rng_epoch is a device counter — may be NULL = 0 — so that a captured CUDA graph draws new
 * numbers per replay), and (b) with norm_sq != NULL the squared
 * Frobenius norm of the batch's uncontracted means (what mip360_frustum_norm_sq computes from t_out afterwards;
 * parameterization.py:25,75) ACCUMULATED into *norm_sq in the same pass (directions [B,3] required). */
int mip360_level0_sample(const float* near, const float* far, const float* s_lin, const float* t_rand, int use_rng,
                         unsigned long long rng_seed, unsigned int rng_stream, const unsigned long long* rng_epoch,
                         const float* directions, double* norm_sq, float* t_out, int B, int N, mip360_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K1  cast -> Gaussian -> contract -> IPE (+ view-direction encoding), fused
 *     intern/parameterization.py:85-136 (conical_frustum_to_gaussian, gaussian_to_xyz,
 *     gaussian_contract, contract, para_rays), intern/encoding.py:33-61 and :69-90,
 *     model.py:85-88 / 173-176 (concatenation order).
 *
 *   mip360_frustum_norm_sq: pre-pass for the reference's batch-global contraction
 *     (App. A1): *norm_sq += sum over all samples of |d * t_mean|^2.  *norm_sq must be
 *     zeroed by the caller (so that several shards can be accumulated into one scalar).
 *   t0/t1 are given as two pointers with a common row stride so that both t_vals[:, :-1] /
 *     t_vals[:, 1:] views (stride N+1) and separate contiguous t0, t1 (stride N) work.
 *   contract_mode: 0 = reference (global Frobenius norm, read from *norm_sq),
 *                  1 = per-point contraction (paper), 2 = no contraction.
 *   add_origins: bit 0 = means += origins after the contraction (para_rays, App. A2); clear = not
 *     (conical_frustum_to_gaussian itself).  Bit 1 = use the original frustum formula of
 *     parameterization.py:108-113 (stable=False) instead of the stable one; mip360_frustum_norm_sq always uses the
 *     stable t_mean, so with bit 1 the caller supplies *norm_sq (mip360_sum_sq of the uncontracted means).
 *   vdir_enc [B,16] = mip360_viewdir_enc of the rays' view directions (needed for x_bf16 only).
 *   Outputs, each may be NULL: means [S,3], covs [S,3,3], enc [S,42] (fp32 IPE of the returned
 *     mean and cov), x_bf16 [S,64] = the MLP input row: 42 IPE features, 16 view-direction
 *     features, 6 zeros, rounded to bf16.
 * ------------------------------------------------------------------------------------------ */
int mip360_frustum_norm_sq(const float* t0, const float* t1, int t_stride, const float* directions, int B, int N,
                           double* norm_sq, mip360_stream_t stream);
int mip360_cast_ipe(const float* t0, const float* t1, int t_stride, const float* origins, const float* directions,
                    const float* vdir_enc, const float* radii, const double* norm_sq, int B, int N,
                    int contract_mode, int add_origins, float* means, float* covs, float* enc, uint16_t* x_bf16,
                    mip360_stream_t stream);

/* mip360_cast_ipe for other view-direction degrees (ViewdirectionEncoding(min_deg, max_deg), encoding.py:63-90):
 * vdir_enc [B, vd_dim] with vd_dim = 4 * (max_deg - min_deg), rows of x_cols = 64 (vd_dim <= 20) or 128 (vd_dim <= 84)
 * bf16 columns: 42 IPE features, vd_dim view-direction features, zeros.  The MLP's first layer is packed to the same
 * width (K padded to x_cols). */
int mip360_cast_ipe_x(const float* t0, const float* t1, int t_stride, const float* origins, const float* directions,
                      const float* vdir_enc, int vd_dim, const float* radii, const double* norm_sq, int B, int N,
                      int contract_mode, int add_origins, float* means, float* covs, float* enc, uint16_t* x_bf16,
                      int x_cols, mip360_stream_t stream);

/* stand-alone pieces of the same arithmetic, for the reference's unfused free functions */
/* intern/parameterization.py:31-62 (diag=False): d [B,3], t_mean/t_var/r_var [B,N] */
int mip360_gaussian_to_xyz(const float* directions, const float* t_mean, const float* t_var, const float* r_var,
                           int B, int N, float* means, float* covs, mip360_stream_t stream);
/* intern/parameterization.py:49-53 (diag=True): means [S,3] and the covariance diagonal cov_diag [S,3] */
int mip360_gaussian_to_xyz_diag(const float* directions, const float* t_mean, const float* t_var, const float* r_var,
                                int B, int N, float* means, float* cov_diag, mip360_stream_t stream);
/* sum of squares of n floats accumulated into *out (double, caller-zeroed): torch.norm of parameterization.py:25 */
int mip360_sum_sq(const float* x, long long n, double* out, mip360_stream_t stream);
/* intern/parameterization.py:23-29: y = x if ||x||_F <= 1 else (2-1/n)(x/n), n^2 = *norm_sq */
int mip360_contract(const float* x, long long n, const double* norm_sq, float* y, mip360_stream_t stream);
/* intern/parameterization.py:64-83 on S samples; covs_in/out [S,3,3]; closed-form Jacobian (App. A2) */
int mip360_gaussian_contract(const float* means_in, const float* covs_in, const double* norm_sq, long long S,
                             float* means_out, float* covs_out, mip360_stream_t stream);
/* intern/encoding.py:33-61: mean [S,3], cov [S,3,3] -> enc [S,42] */
int mip360_ipe(const float* means, const float* covs, long long S, float* enc, mip360_stream_t stream);
/* intern/encoding.py:69-90: viewdirs [B,3] -> enc [B,4*n_scales], scales 2^min_deg .. 2^(max_deg-1) */
int mip360_viewdir_enc(const float* viewdirs, int B, int min_deg, int max_deg, float* enc, mip360_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4  hierarchical resampling         intern/ray.py:12-57 and :118-153
 *   mip360_blur_weights:  ray.py:137-142 (max-pool 2, avg-pool 2, + resample_padding) -> [B,N]
 *   mip360_resample_cdf:  ray.py:15-27   weights [B,N] -> cdf [B,N+1]
 *   mip360_resample_invert: ray.py:41-56 bins [B,N+1], cdf [B,N+1], u [B,M] (u_row_stride 0 =
 *     one shared row) -> samples [B,M] and, if idx != NULL, idx [B,M] = index of the last cdf knot
 *     <= u ("bin index"; bit-exact vs. searchsorted(right)-1, App. B2).
 *   mip360_resample: the whole no_grad block of ray.py:136-149 in one kernel.  u_base [M] is the
 *     stratum vector (randomized: arange(M)*(1/M); deterministic: linspace(0,1-eps,M)), jitter
 *     [B,M] the uniform_(0,1/M-eps) draw or NULL; randomized => u = min(u_base+u_base+jitter,1-eps)
 *     (the doubling is the reference's, App. A5).  M must equal N+1 (ray.py:146).
 * ------------------------------------------------------------------------------------------ */
int mip360_blur_weights(const float* weights, int B, int N, float resample_padding, float* out,
                        mip360_stream_t stream);
int mip360_resample_cdf(const float* weights, int B, int N, float* cdf, mip360_stream_t stream);
int mip360_resample_invert(const float* bins, const float* cdf, const float* u, int u_row_stride, int B, int N,
                           int M, float* samples, int32_t* idx, mip360_stream_t stream);
int mip360_resample(const float* t_vals, const float* weights, const float* u_base, const float* jitter,
                    int B, int N, float resample_padding, int blur, float* new_t, mip360_stream_t stream);
/* The same with (a) the jitter of ray.py:33 drawn inside the kernel when use_rng = 1: jitter = u01 * jitter_scale with
 * jitter_scale = 1/M - eps as torch's uniform_(0, 1/M - eps) scales it and u01 = number m of ray b from the generator
 * described at mip360_level0_sample, and (b) with norm_sq != NULL the squared Frobenius norm of the
 * uncontracted means of the NEW knots accumulated into *norm_sq (what mip360_frustum_norm_sq computes from new_t). */
int mip360_resample_sample(const float* t_vals, const float* weights, const float* u_base, const float* jitter,
                           int use_rng, unsigned long long rng_seed, unsigned int rng_stream,
                           const unsigned long long* rng_epoch, float jitter_scale, const float* directions,
                           double* norm_sq, int B, int N, float resample_padding, int blur, float* new_t,
                           mip360_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3  volume compositing              intern/ray.py:155-191, model.py:59-78, model.py:184-185
 *   head_mode 0: rgb [B,N,3] and density [B,N] are final values (volumetric_rendering as is).
 *   head_mode 1: raw [B,N,4] = (sigmoid density head, sigmoid colour head) straight from the MLP;
 *     density = softplus(raw0 + density_bias), rgb = raw123*(1+2*rgb_padding) - rgb_padding.
 *   head_mode 2: raw [B,N,4] = the heads' pre-activation sums without bias (mip360_mlp_fwd_fused_head); the bias
 *     head_bias[4] (device) and the Sigmoid of model.py:150-158 are applied here, then as head_mode 1.  Backward:
 *     g_raw = dL/d(pre-activation), the Sigmoid derivative included.
 *   weights-only variant (density_to_weight): density_mode 0 = density given, 1 = raw logits,
 *     density = softplus(raw + density_bias) (model.py:92).
 *   Backward: g_rgb [B,3], g_acc [B], g_dist [B], g_w [B,N] (any may be NULL) -> gradient w.r.t. rgb/density
 *     (head_mode 0: g_rgb_in [B,N,3], g_density [B,N]) or the head outputs (head_mode 1: g_raw
 *     [B,N,4] = dL/d raw).  density_to_weight: g_density = dL/d density (mode 0) or dL/d raw logit
 *     (mode 1).  distance is differentiable w.r.t. density (through the weights, inside its clamp
 *     range, like torch.clamp / nan_to_num); t_vals and dirs carry no gradient (they never do in the reference:
 *     resampling runs under no_grad, ray.py:136).
 *   mip360_head_grad_pack: fp32 head-output gradient [M,n_valid] -> bf16 rows of 64 (zero padded) that
 *     feed the head dgrad/wgrad GEMMs; act 2 folds the Sigmoid derivative y(1-y) of the saved head
 *     output y (model.py:150-158), act 0 copies.
 * ------------------------------------------------------------------------------------------ */
int mip360_composite_fwd(const float* rgb_or_raw, const float* density, const float* t_vals, const float* dirs,
                         int B, int N, int head_mode, float density_bias, float rgb_padding, int white_bkgd,
                         float* comp_rgb, float* distance, float* acc, float* weights, mip360_stream_t stream);
/* mip360_composite_fwd plus model.py:196 in the same pass: s_vals [B,N+1] = t_to_s(t_vals, near, far) and (optional)
 * t_shift = t_vals + 1e-6 exactly as mip360_t_to_s produces them (near, far [B]); head_bias [4] for head_mode 2. */
int mip360_composite_fwd_s(const float* rgb_or_raw, const float* density, const float* t_vals, const float* dirs,
                           int B, int N, int head_mode, float density_bias, float rgb_padding, int white_bkgd,
                           float* comp_rgb, float* distance, float* acc, float* weights, const float* near,
                           const float* far, float* s_vals, float* t_shift, const float* head_bias,
                           mip360_stream_t stream);
int mip360_composite_bwd(const float* rgb_or_raw, const float* density, const float* t_vals, const float* dirs,
                         int B, int N, int head_mode, float density_bias, float rgb_padding, int white_bkgd,
                         const float* g_rgb, const float* g_acc, const float* g_dist, const float* g_w,
                         float* g_rgb_in, float* g_density, float* g_raw, const float* head_bias,
                         mip360_stream_t stream);
int mip360_density_to_weight_fwd(const float* density, const float* t_vals, const float* dirs, int B, int N,
                                 int density_mode, float density_bias, float* weights, mip360_stream_t stream);
int mip360_density_to_weight_bwd(const float* density, const float* t_vals, const float* dirs, int B, int N,
                                 int density_mode, float density_bias, const float* g_w, float* g_density,
                                 mip360_stream_t stream);
int mip360_head_grad_pack(const float* g, const float* y, long long M, int n_valid, int act, uint16_t* out_bf16,
                          mip360_stream_t stream);
/* intern/parameterization.py:5-8 incl. the in-place eps shifts one call observes (App. A4):
 * s = (1/(t+e) - 1/(near+e)) / (1/(far+e) - 1/(near+2e)); t_shift = t + e (may be NULL) */
int mip360_t_to_s(const float* t_vals, const float* near, const float* far, int B, int K, float* s_vals,
                  float* t_shift, mip360_stream_t stream);
/* intern/parameterization.py:10-13 */
int mip360_s_to_t(const float* s_vals, const float* near, const float* far, int B, int K, float* t_vals,
                  mip360_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K5  distortion regulariser          intern/regularization.py:3-19 (O(N) form, App. A9/B4)
 *   loss = sum over rays; per_ray [B] optional; partials = workspace of >= mip360_partials_len(B)
 *   doubles.  bwd: g_w [B,N] = g_loss * dloss/dw (g_loss read from device scalar g_loss_ptr).
 * ------------------------------------------------------------------------------------------ */
int mip360_partials_len(int B);
int mip360_distortion_fwd(const float* s_vals, const float* weights, int B, int N, float* per_ray,
                          double* partials, float* loss, mip360_stream_t stream);
int mip360_distortion_bwd(const float* s_vals, const float* weights, int B, int N, const float* g_loss_ptr,
                          float* g_w, mip360_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K6  interlevel (proposal) loss      intern/distillation.py:4-51, intern/loss.py:6-21
 *   mip360_bounds_per_ray: b[r,i] = sum_j w_fine[r,j] [t0_j <= R_i and t1_j >= L_i] (App. B5).
 *   mip360_bounds_reduce: column sums over rays -> bound_total [N] (the value the reference
 *     broadcasts to every ray, App. A6); accumulates into bound_total (caller-zeroed doubles) so
 *     shards / ranks can be combined.
 *   mip360_interlevel_fwd: loss = sum relu(bnd - w_hat)^2/(w_hat+1e-6) / batch_div.  bound_mode 0:
 *     bnd = bound_total[i] (reference), 1: bnd = b[r,i] (per-ray, paper-style).
 *   mip360_interlevel_bwd: g_w_hat = g_loss * d loss / d w_hat (bounds are detached).
 * ------------------------------------------------------------------------------------------ */
int mip360_bounds_per_ray(const float* t_fine, const float* w_fine, const float* t_coarse, int B, int N,
                          float* b_out, mip360_stream_t stream);
int mip360_bounds_reduce(const float* b, int B, int N, double* bound_total, mip360_stream_t stream);
/* mip360_bounds_per_ray and mip360_bounds_reduce in one pass: b_out [B,N] (may be NULL when only the totals are
 * wanted: the per-ray values then never touch HBM) and the column sums ACCUMULATED into bound_total [N] (may be NULL). */
int mip360_bounds(const float* t_fine, const float* w_fine, const float* t_coarse, int B, int N, float* b_out,
                  double* bound_total, mip360_stream_t stream);
int mip360_interlevel_fwd(const float* w_hat, const float* b_per_ray, const double* bound_total, int B, int N,
                          int bound_mode, float batch_div, double* partials, float* loss, mip360_stream_t stream);
int mip360_interlevel_bwd(const float* w_hat, const float* b_per_ray, const double* bound_total, int B, int N,
                          int bound_mode, float batch_div, const float* g_loss_ptr, float* g_w_hat,
                          mip360_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K2  MLP GEMMs on tcgen05 / TMEM / TMA      model.py:43-53, :91 (prop_net); :131-158, :179-181 (nerf_net)
 *   All matrices bf16 row-major, fp32 accumulation in tensor memory.
 *   mip360_linear_fwd:  Y[M,N] = act(X[M,K] * W[N,K]^T + bias[N]).
 *       act: 0 none, 1 ReLU, 2 Sigmoid.  out_bf16 [M,N] and/or out_f32 [M,n_valid] (heads: only the
 *       first n_valid columns are stored, fp32).  K % 64 == 0, N % 64 == 0, N <= 256 or N % 256 == 0.
 *   mip360_linear_dgrad: dX[M,K] = (dY[M,N] * Wt[K,N]^T) .* act'(Yprev[M,K]) with Wt = W^T stored
 *       [K,N] row-major, act' from the saved *output* of the previous layer (ReLU: y>0, Sigmoid: y(1-y)).
 *   mip360_linear_wgrad: dW[N,K] (+)= dY[M,N]^T * X[M,K] (fp32, split over M with atomic adds, so dW
 *       must be zeroed by the caller when accumulate == 0 is not used); db[N] += column sums of dY.
 *   mip360_cast_weight: fp32 W[N,K] -> bf16 Wb[Npad,Kpad] (zero padded) and optionally Wt[Kpad,Npad].
 * ------------------------------------------------------------------------------------------ */
int mip360_linear_fwd(const uint16_t* X, const uint16_t* W, const float* bias, int M, int N, int K, int act,
                      uint16_t* out_bf16, float* out_f32, int n_valid, mip360_stream_t stream);
/* The last trunk layer with the (<= 4 wide) head folded into its epilogue (model.py:147-158,180-181): Y as above
 * (out_bf16 may be NULL: inference does not need the trunk output) and head_out[M,4] += act(...) * head_w4 with head_w4
 * fp32 [N,4] (column-interleaved head weights).  head_out is ACCUMULATED into (the column tiles of a row arrive from
 * different CTAs): zero it first; bias and head activation are applied by the consumer (mip360_composite_fwd_s,
 * head_mode 2). */
int mip360_linear_fwd_head(const uint16_t* X, const uint16_t* W, const float* bias, int M, int N, int K, int act,
                           uint16_t* out_bf16, const float* head_w4, float* head_out, mip360_stream_t stream);
/* Backward of a fused head in one pass over the saved trunk output Y [M,N] (bf16; act = the trunk layer's activation):
 * g [M,4] = dL/d(head pre-activation), head_w4 [N,4] as above.  Writes dZ [M,N] bf16 = (g head_w4^T) .* act'(Y), the
 * gradient entering the last trunk layer, and ACCUMULATES dWh[h, c] += sum_r g[r,h] Y[r,c] (row pitch ldw floats) and
 * dbh[h] += sum_r g[r,h] (may be NULL).  Replaces mip360_head_grad_pack + the head's wgrad and dgrad GEMMs. */
int mip360_head_bwd(const float* g, const float* head_w4, const uint16_t* Y, int M, int N, int act, uint16_t* dZ,
                    float* dWh, int ldw, float* dbh, mip360_stream_t stream);
int mip360_linear_dgrad(const uint16_t* dY, const uint16_t* Wt, const uint16_t* Yprev, int M, int N, int K, int act,
                        uint16_t* dX, mip360_stream_t stream);
int mip360_linear_wgrad(const uint16_t* dY, const uint16_t* X, int M, int N, int K, float* dW, float* db,
                        mip360_stream_t stream);
int mip360_cast_weight(const float* W, int N, int K, int Npad, int Kpad, uint16_t* Wb, uint16_t* Wt,
                       mip360_stream_t stream);
/* One layer of an MLP as the GEMM kernels consume it (all device pointers, widths zero padded to 64/128/256k). */
typedef struct mip360_layer {
  const uint16_t* W;  /* bf16 [n_pad, k_pad] */
  const uint16_t* Wt; /* bf16 [k_pad, n_pad] = W^T, needed by the backward pass (NULL for inference) */
  const float* bias;  /* fp32 [n_pad] */
  int n_pad, k_pad;
  int act;            /* activation applied to this layer's output: 0 none, 1 ReLU, 2 Sigmoid */
} mip360_layer;
/* Whole-MLP forward (model.py:91 / :179-181): x bf16 [M, trunk[0].k_pad] -> head outputs out fp32 [M, n_valid].
 *   head: the (merged) output layer, n_pad = 64.  acts: n_act_bufs = n_trunk buffers [M, n_pad_l] to keep every
 *   trunk activation for the backward pass, or 2 ping-pong buffers [M, max n_pad] for inference.
 * Whole-MLP backward: g_out = dL/d out [M, n_valid], out = the saved head outputs (needed when head->act == 2).
 *   dW[l] fp32 [n_pad_l, k_pad_l] and db[l] [n_pad_l] for l < n_trunk, dW[n_trunk] [64, k_pad] / db[n_trunk] [64]
 *   for the head, all ACCUMULATED into (split-K atomics).  dz_head [M,64], dz0, dz1 [M, max n_pad]: bf16 scratch. */
int mip360_mlp_fwd(const uint16_t* x, int M, const mip360_layer* trunk, int n_trunk, const mip360_layer* head,
                   int n_valid, uint16_t* const* acts, int n_act_bufs, float* out, mip360_stream_t stream);
/* mip360_mlp_fwd for the proposal net's shape — x [M,64] -> 4 layers of width 256 (ReLU / Sigmoid) -> head [64 padded, 256]
 * without activation — as ONE persistent kernel: a 128-row tile's activations stay in shared memory (written by the epilogue
 * in the UMMA operand layout) and tensor memory, weights are streamed from L2.  acts: NULL (inference) or 4 buffers
 * [M,256] that receive the trunk activations for the backward pass.  Bit-identical to the layer-by-layer path.  Other
 * shapes: MIP360_ERR_UNSUPPORTED. */
int mip360_mlp_fwd_fused_narrow(const uint16_t* x, int M, const mip360_layer* trunk, int n_trunk, const mip360_layer* head,
                                int n_valid, uint16_t* const* acts, float* out, mip360_stream_t stream);
/* mip360_mlp_fwd with the head folded into the last trunk layer (mip360_linear_fwd_head): head_w4 fp32 [k_pad, 4].
 * `out` [M,4] receives the head's PRE-activation sums WITHOUT bias (zeroed inside).  With n_act_bufs == 2 the last trunk
 * activation is not written at all. */
int mip360_mlp_fwd_fused_head(const uint16_t* x, int M, const mip360_layer* trunk, int n_trunk, const float* head_w4,
                              uint16_t* const* acts, int n_act_bufs, float* out, mip360_stream_t stream);
/* Backward of mip360_mlp_fwd_fused_head: g_out [M,4] = dL/d(out) (pre-activation sums), mip360_head_bwd for the head,
 * then the trunk as in mip360_mlp_bwd.  dW[n_trunk] is the head's [64, k_pad] gradient (rows 0..3 written), db[n_trunk] [64]. */
int mip360_mlp_bwd_fused_head(const float* g_out, const uint16_t* x, int M, const mip360_layer* trunk, int n_trunk,
                              const float* head_w4, uint16_t* const* acts, float* const* dW, float* const* db,
                              uint16_t* dz0, uint16_t* dz1, mip360_stream_t stream);
int mip360_mlp_bwd(const float* g_out, const float* out, const uint16_t* x, int M, const mip360_layer* trunk, int n_trunk,
                   const mip360_layer* head, int n_valid, uint16_t* const* acts, float* const* dW, float* const* db,
                   uint16_t* dz_head, uint16_t* dz0, uint16_t* dz1, mip360_stream_t stream);
/* number of SMs the persistent GEMM grids are sized for (148 on B200) */
int mip360_sm_count(void);

/* ------------------------------------------------------------------------------------------
 * Ray generation on the device (SURVEY §8f rank 1 — the step immediately before the hot path)
 *   dataset.py:109-145 (pinhole camera model, radii from the row-neighbour direction distance) and, with
 *   ndc = 1, dataset.py:364-387 + intern/ray.py:59-79 (NDC origins/directions at plane ndc_near, radii from
 *   both NDC-origin neighbours).  c2w [n_img, c2w_rows >= 3, 4] row-major camera-to-world matrices.
 *   Outputs are flattened like dataset.py:147-152: ray index = (img*H + y)*W + x; origins, directions,
 *   viewdirs [n,3]; radii, near, far [n,1].
 * ------------------------------------------------------------------------------------------ */
int mip360_generate_rays(const float* c2w, int c2w_rows, int n_img, int H, int W, float focal, float near, float far,
                         int ndc, float ndc_near, float* origins, float* directions, float* viewdirs, float* radii,
                         float* near_out, float* far_out, mip360_stream_t stream);
/* The same for the slab [ray_begin, ray_begin + ray_count) of the flattened ray index only (outputs hold ray_count
 * rays): a render chunk loop (model.py:262-264) or a rank of a ray-partitioned render generates exactly the rays it
 * is about to consume, so a frame's rays are never materialised or uploaded. */
int mip360_generate_rays_range(const float* c2w, int c2w_rows, int n_img, int H, int W, float focal, float near,
                               float far, int ndc, float ndc_near, long long ray_begin, long long ray_count,
                               float* origins, float* directions, float* viewdirs, float* radii, float* near_out,
                               float* far_out, mip360_stream_t stream);

/* intern/utils.py:17-21 (to8b, used by model.render_image, model.py:270): uint8 = 255 * clip(nan_to_num(x), 0, 1),
 * truncated like NumPy's astype(uint8); n elements (the image leaves the device as 3 B/pixel). */
int mip360_to8b(const float* x, long long n, uint8_t* out, mip360_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Depth / normal visualisation of a rendered frame    intern/pose.py:112-212
 *   mip360_normals_scaling: pose.py:130-136 — over the non-NaN pixels of depth [H,W]: stats[0] count,
 *       [1..3] mean of (x = column, y = row, depth), [4..6] their population variances,
 *       [7] scaling = sqrt(((var x + var y)/2) / var depth).  fp64, two passes, deterministic.
 *       partials: mip360_vis_partials_len() doubles of workspace.
 *   mip360_visualize_normals: pose.py:112-121 + :137-145 — normals of stats[7] * depth through the 3x3
 *       blur/edge convolutions (zero fill outside the frame, as scipy 'same'), shaded (n + 1)/2 with
 *       NaN -> 1, blended to white by acc (optional).  Writes vis [H,W,3] fp32 and/or vis8 [H,W,3] uint8
 *       (= to8b(vis), utils.py:17-21).
 *   mip360_depth_range: pose.py:180-194 — range[0] = near, range[1] = far; a value whose auto_* flag is set
 *       is replaced by the acc-weighted quantile of depth the reference reads off its argsort + cumsum:
 *       near = first depth (ascending, NaN last) whose cumulative acc >= ignore_frac * total, minus eps;
 *       far = last depth whose cumulative acc <= (1 - ignore_frac) * total, plus eps.  Found without a sort
 *       by a 32-step bisection over the order-preserving integer image of the depths; acc enters in exact
 *       integer units of 2^-24.  work: mip360_vis_work_len() unsigned 64-bit words.
 *   mip360_visualize_depth: pose.py:196-212 — curve (0: -log(x+eps), 1: x, 2: 1/(x+eps), 3: log(x+eps)) applied
 *       to depth, near, far; modulus > 0: value = mod(curved, modulus)/modulus, else clip((curved - min)/|far - near|);
 *       colour = lut[trunc(value * n_lut)] (lut [n_lut,3] fp32, listed-colour-map indexing) or, with lut NULL,
 *       the sinebow map (pose.py:123-126); blended to white by acc (NaN depth -> acc 0).
 * ------------------------------------------------------------------------------------------ */
int mip360_vis_partials_len(void);
int mip360_vis_work_len(void);
int mip360_normals_scaling(const float* depth, int H, int W, double* partials, double* stats, mip360_stream_t stream);
int mip360_visualize_normals(const float* depth, const float* acc, const double* stats, int H, int W, float* vis,
                             uint8_t* vis8, mip360_stream_t stream);
int mip360_depth_range(const float* depth, const float* acc, long long n, double ignore_frac, float near, float far,
                       int auto_near, int auto_far, unsigned long long* work, float* range, mip360_stream_t stream);
int mip360_visualize_depth(const float* depth, const float* acc, const float* range, int curve, float modulus,
                           const float* lut, int n_lut, long long n, float* vis, uint8_t* vis8, mip360_stream_t stream);

/* fused AdamW over one flat fp32 parameter tensor (train.py:38,63,81 — "next" row f3 of SURVEY §8):
 * decoupled weight decay, bias correction from `step` (1-based), optional bf16 re-cast of the weights */
int mip360_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                 float eps, float weight_decay, int step, mip360_stream_t stream);

/* One launch per net: [AdamW ->] refresh of the GEMM operands ("fused multi-tensor AdamW with the
 * bf16 weight re-cast folded in").  Each table entry describes one layer of a packed MLP: up to two source blocks
 * of fp32 rows (nn.Linear weights [rows, K] stacked into one padded [n_pad, k_pad] matrix — the NeRF heads
 * final_density / final_color share one 64-row head), their biases, and the destinations: bf16 Wb [n_pad, k_pad], its
 * transpose Wt [k_pad, n_pad] (may be NULL) and the padded fp32 bias [n_pad] (may be NULL).  tile_begin = first 32x32
 * tile of the entry in the launch (entries sorted, total_tiles = sum of (n_pad/32)*(k_pad/32)).
 * do_adam = 1: the sources lie inside the flat parameter buffer p; the element at p[i] is updated from g[i], m[i], v[i]
 *   exactly as mip360_adamw does, then cast.  hyper_dev (device, may be NULL) = {lr, 1 - beta1^step, sqrt(1 - beta2^step)}
 *   overrides lr / step, so that a captured CUDA graph can be replayed with this step's values.  zero_grad = 1 clears g[i]
 *   after use (the next backward pass accumulates into a zeroed buffer without a separate memset).
 * do_adam = 0: cast only (after load_state_dict or an external optimiser changed the fp32 parameters). */
typedef struct mip360_pack_entry {
  const float* w_src[2]; /* fp32 [rows[i], K] row-major; w_src[1] NULL when rows[1] == 0 */
  const float* b_src[2]; /* fp32 [rows[i]] */
  int rows[2];
  int K;
  int n_pad, k_pad;      /* multiples of 32 */
  int tile_begin;
  uint16_t* Wb;
  uint16_t* Wt;
  float* bias;
  float* w4;             /* NULL, or fp32 [k_pad, 4]: rows 0..3 of the padded matrix (bf16-rounded), column-interleaved —
                            the head weights as mip360_linear_fwd_head reads them */
} mip360_pack_entry;
int mip360_adamw_pack(const mip360_pack_entry* entries, int n_entries, int total_tiles, float* p, float* g, float* m,
                      float* v, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                      const float* hyper_dev, int do_adam, int zero_grad, mip360_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MIP360_B200_H */
