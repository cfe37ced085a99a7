"""Host-side operators over the C ABI (include/mip360_b200.h): output allocation, argument checking and
torch.autograd.Function wrappers for the fwd/bwd kernel pairs.  torch is plumbing only (device
memory, streams, autograd bookkeeping); every arithmetic step runs in libmip360_b200.so.

Reference functions replaced (paths relative to the reference root) are cited per operator.
"""
from __future__ import annotations

import functools

import torch

from . import _lib
from ._lib import call, check_cuda, f32c, ptr

EPS32 = float(torch.finfo(torch.float32).eps)

CONTRACT_REFERENCE, CONTRACT_PER_POINT, CONTRACT_NONE = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2


def _empty(shape, like, dtype=torch.float32):
    return torch.empty(shape, device=like.device, dtype=dtype)


# ------------------------------------------------------------------------------------------------
# K0 / K1
# ------------------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------------
# random draws: generated inside the consuming kernels (csrc/common.cuh), keyed by torch's seed
# ------------------------------------------------------------------------------------------------
class _Rng:
    torch_seed = None
    seed = 0
    calls = 0
    epochs = {}


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return x ^ (x >> 31)


def rng_epoch(device):
    """Device-resident replay counter of the in-kernel generator (one uint64 per device)."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    ep = _Rng.epochs.get(key)
    if ep is None:
        ep = _Rng.epochs[key] = torch.zeros(1, device=torch.device("cuda", key), dtype=torch.int64)
    return ep


def rng_advance(device):
    """Bump the replay counter: as the first node of a captured CUDA graph it makes every replay draw fresh numbers."""
    rng_epoch(device).add_(1)


def manual_seed(seed):
    """torch.manual_seed(seed) and a restart of the in-kernel generator's call counter (re-seeding torch with the SAME
    value is not visible through torch.initial_seed(), so exact restarts go through this function)."""
    torch.manual_seed(seed)
    _Rng.torch_seed, _Rng.seed, _Rng.calls = torch.initial_seed(), _splitmix64(torch.initial_seed() & 0xFFFFFFFFFFFFFFFF), 0


def rng_next(device):
    """(seed, stream id, epoch tensor) for the next randomized kernel call.  The key follows torch.manual_seed
    (torch.initial_seed()), every call gets its own stream id, so runs are reproducible per seed and call order."""
    ts = torch.initial_seed()
    if ts != _Rng.torch_seed:
        _Rng.torch_seed, _Rng.seed, _Rng.calls = ts, _splitmix64(ts & 0xFFFFFFFFFFFFFFFF), 0
    _Rng.calls += 1
    return _Rng.seed, _Rng.calls & 0xFFFFFFFF, rng_epoch(device)


_CONSTS = {}


def _const_vector(key, make, device):
    """Small constant device vectors (the linspace / arange rows of ray.py:31-38,100), built once per device with the
    reference's torch ops and reused: three launches less per forward.  Nothing is cached while a CUDA graph is being
    captured (such a tensor would live in the graph's private pool)."""
    k = (key, device.type, device.index)
    v = _CONSTS.get(k)
    if v is None:
        v = make()
        if not (device.type == "cuda" and torch.cuda.is_current_stream_capturing()):
            _CONSTS[k] = v
    return v


def level0_t_vals(near, far, num_samples, randomized, t_rand=None, directions=None, norm_sq=None, rng=None):
    """intern/ray.py:100-111.  near/far [B,1]; returns t_vals [B,N+1].  randomized: the uniforms of ray.py:106 are the
    given t_rand [B,N+1] or, by default, drawn inside the kernel (rng = (seed, stream id, epoch tensor) to pin them).
    norm_sq (zeroed fp64 [1]) + directions: also accumulate the batch's squared contraction norm (App. A1) into it."""
    near, far = f32c(near), f32c(far)
    check_cuda(near, far)
    B = near.shape[0]
    s_lin = _const_vector(("s_lin", num_samples), lambda: torch.linspace(0.0, 1, num_samples + 1, device=near.device),
                          near.device)
    use_rng, seed, stream_id, epoch = 0, 0, 0, None
    if not randomized:
        t_rand = None
    elif t_rand is not None:
        t_rand = f32c(t_rand)
    else:
        seed, stream_id, epoch = rng if rng is not None else rng_next(near.device)
        use_rng = 1
    t = _empty((B, num_samples + 1), near)
    directions = f32c(directions) if norm_sq is not None else None
    call("mip360_level0_sample", ptr(near), ptr(far), ptr(s_lin), ptr(t_rand), use_rng, seed, stream_id, ptr(epoch),
         ptr(directions), ptr(norm_sq), ptr(t), B, num_samples)
    return t


def viewdir_enc(viewdirs, min_deg=0, max_deg=4):
    """intern/encoding.py:69-90."""
    v = f32c(viewdirs)
    check_cuda(v)
    B = v.shape[0]
    out = _empty((B, 4 * (max_deg - min_deg)), v)
    call("mip360_viewdir_enc", ptr(v), B, int(min_deg), int(max_deg), ptr(out))
    return out


def frustum_norm_sq(t0, t1, t_stride, directions, B, N, out=None):
    """Squared Frobenius norm the reference's contract() sees (intern/parameterization.py:25,75; App. A1).
    Returns a device double scalar (accumulates into `out` if given)."""
    if out is None:
        out = torch.zeros(1, device=directions.device, dtype=torch.float64)
    call("mip360_frustum_norm_sq", t0, t1, t_stride, ptr(directions), B, N, ptr(out))
    return out


def cast_ipe(t_vals, origins, directions, radii, vdir_enc=None, *, t0=None, t1=None, contract_mode=CONTRACT_REFERENCE,
             add_origins=True, norm_sq=None, want_means=False, want_covs=False, want_enc=False, want_x=False,
             stable=True):
    """Fused cast -> Gaussian -> contract -> IPE (intern/parameterization.py:85-136, intern/encoding.py:33-61,
    model.py:85-88).  Either t_vals [B,N+1] or separate t0, t1 [B,N].  Returns dict of requested outputs.
    stable=False selects the original frustum formula (parameterization.py:108-113)."""
    directions, radii = f32c(directions), f32c(radii)
    origins = f32c(origins) if origins is not None else None
    if t_vals is not None:
        t_vals = f32c(t_vals)
        check_cuda(t_vals)
        B, N = t_vals.shape[0], t_vals.shape[1] - 1
        p0, p1, stride = t_vals.data_ptr(), t_vals.data_ptr() + 4, N + 1
        keep = (t_vals,)
    else:
        t0, t1 = f32c(t0), f32c(t1)
        check_cuda(t0, t1)
        B, N = t0.shape
        p0, p1, stride = t0.data_ptr(), t1.data_ptr(), N
        keep = (t0, t1)
    check_cuda(directions, radii, origins, vdir_enc)
    dev = directions
    if B == 0:
        return dict(means=_empty((0, N, 3), dev) if want_means else None,
                    covs=_empty((0, N, 3, 3), dev) if want_covs else None,
                    enc=_empty((0, N, 42), dev) if want_enc else None,
                    x=_empty((0, 64 if vdir_enc is None or vdir_enc.shape[-1] <= 22 else 128), dev, torch.bfloat16)
                    if want_x else None, norm_sq=norm_sq)
    flags = int(bool(add_origins)) | (0 if stable else 2)
    if contract_mode == CONTRACT_REFERENCE and norm_sq is None:
        if stable:
            norm_sq = frustum_norm_sq(p0, p1, stride, directions, B, N)
        else:  # the pre-pass kernel evaluates the stable t_mean: take the norm of the uncontracted means instead
            raw = _empty((B, N, 3), dev)
            call("mip360_cast_ipe", p0, p1, stride, None, ptr(directions), None, ptr(radii), None, B, N, CONTRACT_NONE, 2,
                 ptr(raw), None, None, None)
            norm_sq = sum_sq(raw)
    out = {}
    means = _empty((B, N, 3), dev) if want_means else None
    covs = _empty((B, N, 3, 3), dev) if want_covs else None
    enc = _empty((B, N, 42), dev) if want_enc else None
    if want_x and vdir_enc is None:
        raise _lib.Mip360Error("cast_ipe: the bf16 MLP input needs vdir_enc")
    vd_dim = 0 if vdir_enc is None else int(vdir_enc.shape[-1])
    x_cols = 64 if 42 + vd_dim <= 64 else 128  # bf16 row width = the first MLP layer's padded K
    if want_x and (vd_dim % 4 or 42 + vd_dim > 128):
        raise _lib.Mip360Error(f"cast_ipe: {vd_dim} view-direction features do not fit a 128-column MLP input row")
    x = _empty((B * N, x_cols), dev, torch.bfloat16) if want_x else None
    call("mip360_cast_ipe_x", p0, p1, stride, ptr(origins), ptr(directions), ptr(f32c(vdir_enc) if vdir_enc is not None else None),
         vd_dim, ptr(radii), ptr(norm_sq), B, N, int(contract_mode), flags, ptr(means), ptr(covs), ptr(enc), ptr(x), x_cols)
    del keep
    out.update(means=means, covs=covs, enc=enc, x=x, norm_sq=norm_sq)
    return out


def gaussian_to_xyz(d, t_mean, t_var, r_var, diag=False):
    """intern/parameterization.py:31-62: (means [B,N,3], covs [B,N,3,3]) or, with diag=True, the diagonal [B,N,3]."""
    d, t_mean, t_var, r_var = f32c(d), f32c(t_mean), f32c(t_var), f32c(r_var)
    check_cuda(d, t_mean, t_var, r_var)
    B, N = t_mean.shape
    means = _empty((B, N, 3), d)
    covs = _empty((B, N, 3) if diag else (B, N, 3, 3), d)
    call("mip360_gaussian_to_xyz_diag" if diag else "mip360_gaussian_to_xyz", ptr(d), ptr(t_mean), ptr(t_var), ptr(r_var),
         B, N, ptr(means), ptr(covs))
    return means, covs


def sum_sq(x):
    x = f32c(x)
    check_cuda(x)
    out = torch.zeros(1, device=x.device, dtype=torch.float64)
    call("mip360_sum_sq", ptr(x), x.numel(), ptr(out))
    return out


def contract(x, norm_sq=None):
    """intern/parameterization.py:23-29 (norm over the whole tensor, App. A1)."""
    x = f32c(x)
    check_cuda(x)
    if norm_sq is None:
        norm_sq = sum_sq(x)
    y = torch.empty_like(x)
    call("mip360_contract", ptr(x), x.numel(), ptr(norm_sq), ptr(y))
    return y


def gaussian_contract(mean, cov, norm_sq=None):
    """intern/parameterization.py:64-83 with the closed-form Jacobian (App. A2/B1)."""
    mean, cov = f32c(mean), f32c(cov)
    check_cuda(mean, cov)
    if norm_sq is None:
        norm_sq = sum_sq(mean)
    mo, co = torch.empty_like(mean), torch.empty_like(cov)
    call("mip360_gaussian_contract", ptr(mean), ptr(cov), ptr(norm_sq), mean.numel() // 3, ptr(mo), ptr(co))
    return mo, co


def ipe(mean, cov):
    """intern/encoding.py:33-61; cov None gives the plain positional encoding branch."""
    mean = f32c(mean)
    cov = f32c(cov) if cov is not None else None
    check_cuda(mean, cov)
    enc = _empty(mean.shape[:-1] + (42,), mean)
    call("mip360_ipe", ptr(mean), ptr(cov), mean.numel() // 3, ptr(enc))
    return enc


# ------------------------------------------------------------------------------------------------
# K4
# ------------------------------------------------------------------------------------------------
def blur_weights(weights, resample_padding):
    """intern/ray.py:137-142."""
    w = f32c(weights)
    check_cuda(w)
    out = torch.empty_like(w)
    call("mip360_blur_weights", ptr(w), w.shape[0], w.shape[1], float(resample_padding), ptr(out))
    return out


def resample_cdf(weights):
    """intern/ray.py:15-27."""
    w = f32c(weights)
    check_cuda(w)
    B, N = w.shape
    cdf = _empty((B, N + 1), w)
    call("mip360_resample_cdf", ptr(w), B, N, ptr(cdf))
    return cdf


def resample_invert(bins, cdf, u, return_idx=False):
    """intern/ray.py:41-56; u [B,M] or [M]."""
    bins, cdf, u = f32c(bins), f32c(cdf), f32c(u)
    check_cuda(bins, cdf, u)
    B, K = cdf.shape
    M = u.shape[-1]
    stride = M if u.dim() == 2 and u.shape[0] == B else 0
    samples = _empty((B, M), bins)
    idx = _empty((B, M), bins, torch.int32) if return_idx else None
    call("mip360_resample_invert", ptr(bins), ptr(cdf), ptr(u), stride, B, K - 1, M, ptr(samples), ptr(idx))
    return (samples, idx) if return_idx else samples


def pdf_u_base(num_samples, randomized, device):
    """The per-stratum part of intern/ray.py:31-38, formed with the same torch ops as the reference so that
    it is the same fp32 vector."""
    device = torch.device(device)
    if randomized:
        s = 1 / num_samples
        return _const_vector(("u_rand", num_samples), lambda: torch.arange(num_samples, device=device) * s, device)
    return _const_vector(("u_det", num_samples), lambda: torch.linspace(0.0, 1.0 - EPS32, num_samples, device=device), device)


@functools.lru_cache(maxsize=None)
def jitter_scale(num_samples):
    """Upper end of uniform_(0, 1/M - eps) (intern/ray.py:33) as the fp32 value torch scales its [0,1) draw by."""
    return float(torch.tensor(1 / num_samples - EPS32, dtype=torch.float32))


def draw_jitter(B, num_samples, device):
    """intern/ray.py:33: uniform_(0, 1/M - eps)."""
    s = 1 / num_samples
    return torch.empty(B, num_samples, device=device).uniform_(to=(s - EPS32))


def resample(t_vals, weights, randomized, resample_padding, jitter=None, blur=True, directions=None, norm_sq=None,
             rng=None):
    """The no_grad block of intern/ray.py:136-149 (blur=True) or intern/ray.py:12-57 alone (blur=False).
    randomized: the jitter of ray.py:33 is the given [B,N+1] draw or, by default, generated inside the kernel.
    norm_sq (zeroed fp64 [1]) + directions: also accumulate the squared contraction norm of the NEW knots into it."""
    t_vals, weights = f32c(t_vals.detach()), f32c(weights.detach())
    check_cuda(t_vals, weights)
    B, N = weights.shape
    u_base = f32c(pdf_u_base(N + 1, randomized, t_vals.device))
    use_rng, seed, stream_id, epoch = 0, 0, 0, None
    if not randomized:
        jitter = None
    elif jitter is not None:
        jitter = f32c(jitter)
    else:
        seed, stream_id, epoch = rng if rng is not None else rng_next(t_vals.device)
        use_rng = 1
    new_t = _empty((B, N + 1), t_vals)
    directions = f32c(directions) if norm_sq is not None else None
    call("mip360_resample_sample", ptr(t_vals), ptr(weights), ptr(u_base), ptr(jitter), use_rng, seed, stream_id, ptr(epoch),
         jitter_scale(N + 1), ptr(directions), ptr(norm_sq), B, N, float(resample_padding), int(bool(blur)), ptr(new_t))
    return new_t


# ------------------------------------------------------------------------------------------------
# K3
# ------------------------------------------------------------------------------------------------
class _Composite(torch.autograd.Function):
    """intern/ray.py:155-191 (+ model.py:184-185 when head_mode=1)."""

    @staticmethod
    def forward(ctx, rgb_or_raw, density, t_vals, dirs, head_mode, density_bias, rgb_padding, white_bkgd, s_out=None,
                head_bias=None):
        ctx.set_materialize_grads(False)  # unused outputs (distance, acc in training) arrive as None, not zeros
        B, N = t_vals.shape[0], t_vals.shape[1] - 1
        comp, dist, acc = _empty((B, 3), t_vals), _empty((B,), t_vals), _empty((B,), t_vals)
        w = _empty((B, N), t_vals)
        near, far, s_vals, t_shift = s_out if s_out is not None else (None, None, None, None)
        call("mip360_composite_fwd_s", ptr(rgb_or_raw), ptr(density), ptr(t_vals), ptr(dirs), B, N, head_mode,
             density_bias, rgb_padding, int(white_bkgd), ptr(comp), ptr(dist), ptr(acc), ptr(w), ptr(near), ptr(far),
             ptr(s_vals), ptr(t_shift), ptr(head_bias))
        ctx.head_bias = head_bias
        ctx.save_for_backward(rgb_or_raw, density, t_vals, dirs)
        ctx.cfg = (head_mode, density_bias, rgb_padding, int(white_bkgd))
        return comp, dist, acc, w

    @staticmethod
    def backward(ctx, g_comp, g_dist, g_acc, g_w):
        rgb_or_raw, density, t_vals, dirs = ctx.saved_tensors
        head_mode, density_bias, rgb_padding, white = ctx.cfg
        B, N = t_vals.shape[0], t_vals.shape[1] - 1
        g_comp = f32c(g_comp) if g_comp is not None else None
        g_acc = f32c(g_acc) if g_acc is not None else None
        g_dist = f32c(g_dist) if g_dist is not None else None
        g_w = f32c(g_w) if g_w is not None else None
        if head_mode >= 1:
            g_raw = torch.empty_like(rgb_or_raw)
            g_rgb_in = g_density = None
        else:
            g_raw = None
            g_rgb_in, g_density = torch.empty_like(rgb_or_raw), torch.empty_like(density)
        call("mip360_composite_bwd", ptr(rgb_or_raw), ptr(density), ptr(t_vals), ptr(dirs), B, N, head_mode,
             density_bias, rgb_padding, white, ptr(g_comp), ptr(g_acc), ptr(g_dist), ptr(g_w), ptr(g_rgb_in),
             ptr(g_density), ptr(g_raw), ptr(ctx.head_bias))
        if head_mode >= 1:
            return g_raw, None, None, None, None, None, None, None, None, None
        return g_rgb_in, g_density, None, None, None, None, None, None, None, None


def _no_grad_inputs(who, **tensors):
    for name, t in tensors.items():
        if t.requires_grad:
            raise _lib.Mip360Error(f"{who}: {name} must not require grad (rays and sample positions carry no gradient "
                                   "in the reference: resampling runs under no_grad, ray.py:136)")


def composite(rgb, density, t_vals, dirs, white_bkgd):
    """volumetric_rendering(rgb [B,N,3], density [B,N,1] or [B,N], t_vals, dirs, white_bkgd).  comp_rgb, distance,
    acc and weights are differentiable w.r.t. rgb and density."""
    _no_grad_inputs("composite", t_vals=t_vals, dirs=dirs)
    rgb, t_vals, dirs = f32c(rgb), f32c(t_vals), f32c(dirs)
    density = f32c(density.reshape(density.shape[0], density.shape[1]))
    check_cuda(rgb, density, t_vals, dirs)
    return _Composite.apply(rgb, density, t_vals, dirs, 0, 0.0, 0.0, bool(white_bkgd))


def composite_heads(raw, t_vals, dirs, density_bias, rgb_padding, white_bkgd, near=None, far=None, head_bias=None):
    """model.py:184-186 fused: raw [B,N,4] = (density head, colour head) post-sigmoid outputs of the MLP — or, with
    head_bias [>=4] (device), the heads' pre-activation sums without bias as the fused-head MLP produces them
    (mlp.PackedMLP.fuse_head): bias and Sigmoid (model.py:150-158) are then applied here.
    With near / far the same launch also produces model.py:196's s_vals = t_to_s(t_vals, near, far) and the shifted
    t_vals the reference returns (App. A4): -> (comp_rgb, distance, acc, weights, s_vals, t_shift)."""
    raw, t_vals, dirs = f32c(raw), f32c(t_vals), f32c(dirs)
    check_cuda(raw, t_vals, dirs)
    mode = 1
    if head_bias is not None:
        check_cuda(head_bias)
        mode, head_bias = 2, head_bias.detach()
    if near is None:
        return _Composite.apply(raw, None, t_vals, dirs, mode, float(density_bias), float(rgb_padding), bool(white_bkgd),
                                None, head_bias)
    near, far = f32c(near), f32c(far)
    check_cuda(near, far)
    s_vals, t_shift = torch.empty_like(t_vals), torch.empty_like(t_vals)
    out = _Composite.apply(raw, None, t_vals, dirs, mode, float(density_bias), float(rgb_padding), bool(white_bkgd),
                           (near, far, s_vals, t_shift), head_bias)
    return out + (s_vals, t_shift)


class _DensityToWeight(torch.autograd.Function):
    """model.py:59-78 (+ model.py:92 when density_mode=1)."""

    @staticmethod
    def forward(ctx, density, t_vals, dirs, density_mode, density_bias):
        B, N = density.shape
        w = _empty((B, N), density)
        call("mip360_density_to_weight_fwd", ptr(density), ptr(t_vals), ptr(dirs), B, N, density_mode, density_bias,
             ptr(w))
        ctx.save_for_backward(density, t_vals, dirs)
        ctx.cfg = (density_mode, density_bias)
        return w

    @staticmethod
    def backward(ctx, g_w):
        density, t_vals, dirs = ctx.saved_tensors
        density_mode, density_bias = ctx.cfg
        B, N = density.shape
        g = torch.empty_like(density)
        call("mip360_density_to_weight_bwd", ptr(density), ptr(t_vals), ptr(dirs), B, N, density_mode, density_bias,
             ptr(f32c(g_w)), ptr(g))
        return g, None, None, None, None


def density_to_weight(t_vals, density, dirs, raw_logits=False, density_bias=0.0):
    density = f32c(density.reshape(density.shape[0], density.shape[1]))
    t_vals, dirs = f32c(t_vals), f32c(dirs)
    check_cuda(density, t_vals, dirs)
    return _DensityToWeight.apply(density, t_vals, dirs, 1 if raw_logits else 0, float(density_bias))


def t_to_s(t_vals, near, far):
    """intern/parameterization.py:5-8; returns (s_vals, t_vals + 1e-6) — the second is what the reference's
    t_vals argument holds after the call (App. A4)."""
    t_vals, near, far = f32c(t_vals), f32c(near), f32c(far)
    check_cuda(t_vals, near, far)
    B, K = t_vals.shape
    s, ts = torch.empty_like(t_vals), torch.empty_like(t_vals)
    call("mip360_t_to_s", ptr(t_vals), ptr(near), ptr(far), B, K, ptr(s), ptr(ts))
    return s, ts


def s_to_t(s_vals, near, far):
    """intern/parameterization.py:10-13."""
    s_vals, near, far = f32c(s_vals), f32c(near), f32c(far)
    check_cuda(s_vals, near, far)
    B, K = s_vals.shape
    t = torch.empty_like(s_vals)
    call("mip360_s_to_t", ptr(s_vals), ptr(near), ptr(far), B, K, ptr(t))
    return t


# ------------------------------------------------------------------------------------------------
# K5 / K6
# ------------------------------------------------------------------------------------------------
def _partials(dev):
    return torch.empty(_lib.load().mip360_partials_len(0), device=dev, dtype=torch.float64)


class _Distortion(torch.autograd.Function):
    """intern/regularization.py:3-19 (sum over the batch, both (i,j) orders, App. A9)."""

    @staticmethod
    def forward(ctx, s_vals, weights):
        B, N = weights.shape
        loss = _empty((), weights)
        call("mip360_distortion_fwd", ptr(s_vals), ptr(weights), B, N, None, ptr(_partials(weights.device)), ptr(loss))
        ctx.save_for_backward(s_vals, weights)
        return loss

    @staticmethod
    def backward(ctx, g):
        s_vals, weights = ctx.saved_tensors
        B, N = weights.shape
        g_w = torch.empty_like(weights)
        call("mip360_distortion_bwd", ptr(s_vals), ptr(weights), B, N, ptr(f32c(g)), ptr(g_w))
        return None, g_w


def distortion_loss(s_vals, weights):
    if s_vals.requires_grad:
        # the reference's s_vals never carries a gradient (resampling runs under no_grad, ray.py:136); refuse
        # instead of silently returning a zero gradient for it
        raise _lib.Mip360Error("distortion_loss: s_vals must not require grad (only the weights are differentiated)")
    s_vals, weights = f32c(s_vals), f32c(weights)
    check_cuda(s_vals, weights)
    return _Distortion.apply(s_vals.detach(), weights)


def distortion_per_ray(s_vals, weights):
    s_vals, weights = f32c(s_vals), f32c(weights)
    check_cuda(s_vals, weights)
    B, N = weights.shape
    per_ray, loss = _empty((B,), weights), _empty((), weights)
    call("mip360_distortion_fwd", ptr(s_vals), ptr(weights), B, N, ptr(per_ray), ptr(_partials(weights.device)),
         ptr(loss))
    return per_ray


def bounds_per_ray(t_fine, w_fine, t_coarse):
    t_fine, w_fine, t_coarse = f32c(t_fine.detach()), f32c(w_fine.detach()), f32c(t_coarse.detach())
    check_cuda(t_fine, w_fine, t_coarse)
    B, N = w_fine.shape
    b = _empty((B, N), w_fine)
    call("mip360_bounds_per_ray", ptr(t_fine), ptr(w_fine), ptr(t_coarse), B, N, ptr(b))
    return b


def bounds_batch_total(t_fine, w_fine, t_coarse, out=None):
    """Batch totals of the proposal bounds per coarse interval (intern/distillation.py:25-29, App. A6) straight from the
    knots and fine weights: the per-ray values never touch HBM.  fp64 [N], accumulated into `out` if given."""
    t_fine, w_fine, t_coarse = f32c(t_fine.detach()), f32c(w_fine.detach()), f32c(t_coarse.detach())
    check_cuda(t_fine, w_fine, t_coarse)
    B, N = w_fine.shape
    if out is None:
        out = torch.zeros(N, device=w_fine.device, dtype=torch.float64)
    call("mip360_bounds", ptr(t_fine), ptr(w_fine), ptr(t_coarse), B, N, None, ptr(out))
    return out


def bounds_total(b_per_ray, out=None):
    """Column sums over rays (fp64): the value intern/distillation.py:25-29 broadcasts to every ray (App. A6)."""
    B, N = b_per_ray.shape
    if out is None:
        out = torch.zeros(N, device=b_per_ray.device, dtype=torch.float64)
    call("mip360_bounds_reduce", ptr(b_per_ray), B, N, ptr(out))
    return out


class _Interlevel(torch.autograd.Function):
    """intern/distillation.py:35-51 given detached bounds."""

    @staticmethod
    def forward(ctx, w_hat, b_per_ray, bound_total, bound_mode, batch_div):
        B, N = w_hat.shape
        loss = _empty((), w_hat)
        call("mip360_interlevel_fwd", ptr(w_hat), ptr(b_per_ray), ptr(bound_total), B, N, bound_mode, batch_div,
             ptr(_partials(w_hat.device)), ptr(loss))
        ctx.save_for_backward(w_hat, b_per_ray, bound_total)
        ctx.cfg = (bound_mode, batch_div)
        return loss

    @staticmethod
    def backward(ctx, g):
        w_hat, b_per_ray, bound_total = ctx.saved_tensors
        bound_mode, batch_div = ctx.cfg
        B, N = w_hat.shape
        g_w = torch.empty_like(w_hat)
        call("mip360_interlevel_bwd", ptr(w_hat), ptr(b_per_ray), ptr(bound_total), B, N, bound_mode, batch_div,
             ptr(f32c(g)), ptr(g_w))
        return g_w, None, None, None, None


def interlevel_loss(w_hat, b_per_ray=None, bound_total=None, per_ray_bounds=False, batch_div=None):
    w_hat = f32c(w_hat)
    check_cuda(w_hat)
    if batch_div is None:
        batch_div = float(w_hat.shape[0])
    return _Interlevel.apply(w_hat, b_per_ray, bound_total, 1 if per_ray_bounds else 0, float(batch_div))


# ------------------------------------------------------------------------------------------------
# K2
# ------------------------------------------------------------------------------------------------
def _pad64(n):
    return (n + 63) // 64 * 64


def cast_weight(W, n_pad=None, k_pad=None, transposed=True):
    """fp32 [N,K] -> bf16 [Npad,Kpad] and its transpose [Kpad,Npad] (zero padded)."""
    W = f32c(W.detach())
    check_cuda(W)
    N, K = W.shape
    n_pad, k_pad = n_pad or _pad64(N), k_pad or _pad64(K)
    Wb = torch.empty((n_pad, k_pad), device=W.device, dtype=torch.bfloat16)
    Wt = torch.empty((k_pad, n_pad), device=W.device, dtype=torch.bfloat16) if transposed else None
    call("mip360_cast_weight", ptr(W), N, K, n_pad, k_pad, ptr(Wb), ptr(Wt))
    return Wb, Wt


def linear_fwd(x, Wb, bias, act, out_f32_cols=0, want_bf16=True):
    """Y = act(x Wb^T + bias): x bf16 [M,K], Wb bf16 [N,K], bias fp32 [N]."""
    M, K = x.shape
    N = Wb.shape[0]
    y = torch.empty((M, N), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    yf = torch.empty((M, out_f32_cols), device=x.device, dtype=torch.float32) if out_f32_cols else None
    call("mip360_linear_fwd", ptr(x), ptr(Wb), ptr(bias), M, N, K, act, ptr(y), ptr(yf), out_f32_cols)
    return y, yf


def linear_fwd_head(x, Wb, bias, act, head_w4, head_out, want_bf16=True):
    """The last trunk layer with the head folded into its epilogue: y = act(x Wb^T + bias) (bf16, optional) and
    head_out [M,4] += y head_w4 (fp32, accumulated: zero it first)."""
    M, K = x.shape
    N = Wb.shape[0]
    y = torch.empty((M, N), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    call("mip360_linear_fwd_head", ptr(x), ptr(Wb), ptr(bias), M, N, K, act, ptr(y), ptr(head_w4), ptr(head_out))
    return y


def head_bwd(g, head_w4, y, act, dWh, dbh):
    """Backward of a fused head in one pass over the saved trunk output y [M,N]: returns dZ [M,N] bf16 and accumulates
    the head's weight / bias gradients into dWh [>=4, N] / dbh [>=4]."""
    g = f32c(g)
    M, N = y.shape
    dz = torch.empty((M, N), device=y.device, dtype=torch.bfloat16)
    call("mip360_head_bwd", ptr(g), ptr(head_w4), ptr(y), M, N, act, ptr(dz), ptr(dWh), dWh.shape[1], ptr(dbh))
    return dz


def linear_dgrad(dY, Wt, y_prev, act, out=None):
    """dX = (dY Wt^T) .* act'(y_prev): dY bf16 [M,N], Wt bf16 [K,N], y_prev bf16 [M,K]."""
    M, N = dY.shape
    K = Wt.shape[0]
    dX = out if out is not None else torch.empty((M, K), device=dY.device, dtype=torch.bfloat16)
    call("mip360_linear_dgrad", ptr(dY), ptr(Wt), ptr(y_prev), M, N, K, act, ptr(dX))
    return dX


def linear_wgrad(dY, x, dW=None, db=None, want_db=True):
    """dW[N,K] += dY^T x, db[N] += colsum(dY): dY bf16 [M,N], x bf16 [M,K]; fp32 outputs (zeroed if new)."""
    M, N = dY.shape
    K = x.shape[1]
    if dW is None:
        dW = torch.zeros((N, K), device=dY.device, dtype=torch.float32)
    if db is None and want_db:
        db = torch.zeros((N,), device=dY.device, dtype=torch.float32)
    call("mip360_linear_wgrad", ptr(dY), ptr(x), M, N, K, ptr(dW), ptr(db))
    return dW, db


def head_grad_pack(g, y, act):
    """fp32 head gradient [M,nv] -> bf16 [M,64] rows with the head activation derivative folded in."""
    g = f32c(g)
    M, nv = g.shape
    out = torch.empty((M, 64), device=g.device, dtype=torch.bfloat16)
    call("mip360_head_grad_pack", ptr(g), ptr(y), M, nv, act, ptr(out))
    return out


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step):
    call("mip360_adamw", ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), float(lr), float(beta1), float(beta2), float(eps),
         float(weight_decay), int(step))


def adamw_pack_step(packed_mlp, p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, hyper_dev=None, zero_grad=False):
    """AdamW over the flat buffers of one net and, in the same launch, the refresh of its bf16 GEMM operands
    (mlp.PackedMLP).  hyper_dev: device tensor [lr, 1 - beta1^step, sqrt(1 - beta2^step)] for graph replays."""
    tab, n_entries, n_tiles = packed_mlp.table()
    call("mip360_adamw_pack", tab, n_entries, n_tiles, ptr(p), ptr(g), ptr(m), ptr(v), float(lr), float(beta1),
         float(beta2), float(eps), float(weight_decay), int(step), ptr(hyper_dev), 1, int(bool(zero_grad)))
    packed_mlp.mark_fresh()


# ------------------------------------------------------------------------------------------------
# ray generation
# ------------------------------------------------------------------------------------------------
def generate_rays(cam_to_world, h, w, focal, near, far, ndc=False, ndc_near=1.0, ray_begin=0, ray_count=None):
    """dataset.py:109-145 (pinhole) / dataset.py:364-387 (LLFF NDC), flattened as dataset.py:147-152.
    cam_to_world [n, >=3, 4] device tensor -> Rays of [n*h*w, c] device tensors; nothing is built on the host.
    ray_begin / ray_count select a slab of the flattened ray index (a render chunk, a rank's partition)."""
    from mipnerf360_b200.intern.ray import Rays
    c2w = f32c(cam_to_world)
    check_cuda(c2w)
    if c2w.dim() == 2:
        c2w = c2w[None]
    total = c2w.shape[0] * h * w
    n = total - ray_begin if ray_count is None else int(ray_count)
    o, d, v = _empty((n, 3), c2w), _empty((n, 3), c2w), _empty((n, 3), c2w)
    r, nr, fr = _empty((n, 1), c2w), _empty((n, 1), c2w), _empty((n, 1), c2w)
    call("mip360_generate_rays_range", ptr(c2w), c2w.shape[1], c2w.shape[0], int(h), int(w), float(focal), float(near),
         float(far), int(bool(ndc)), float(ndc_near), int(ray_begin), n, ptr(o), ptr(d), ptr(v), ptr(r), ptr(nr), ptr(fr))
    return Rays(o, d, v, r, nr, fr)


def to8b(img):
    """intern/utils.py:17-21 on the device: float image -> uint8 (255 * clip(nan_to_num(x), 0, 1), truncated)."""
    x = f32c(img)
    check_cuda(x)
    out = torch.empty(x.shape, device=x.device, dtype=torch.uint8)
    call("mip360_to8b", ptr(x), x.numel(), ptr(out))
    return out


# ---------------------------------------------------------------------------------------------------
# depth / normal visualisation of a rendered frame (intern/pose.py:112-212)
# ---------------------------------------------------------------------------------------------------
CURVES = {"neg_log": 0, "identity": 1, "inverse": 2, "log": 3}


def normals_scaling(depth):
    """pose.py:130-136: fp64 stats [count, mean x/y/z, var x/y/z, scaling] over the non-NaN pixels of depth [H,W]."""
    d = f32c(depth)
    check_cuda(d)
    H, W = d.shape
    lib = _lib.load()
    partials = torch.empty(lib.mip360_vis_partials_len(), device=d.device, dtype=torch.float64)
    stats = torch.zeros(8, device=d.device, dtype=torch.float64)
    call("mip360_normals_scaling", ptr(d), H, W, ptr(partials), ptr(stats))
    return stats


def visualize_normals(depth, acc=None, stats=None, as_uint8=False):
    """pose.py:128-147 on the device: [H,W,3] fp32 shading of the fake normals of depth, or its to8b image."""
    d = f32c(depth)
    check_cuda(d)
    H, W = d.shape
    a = None if acc is None else f32c(acc)
    if stats is None:
        stats = normals_scaling(d)
    out = torch.empty((H, W, 3), device=d.device, dtype=torch.uint8 if as_uint8 else torch.float32)
    call("mip360_visualize_normals", ptr(d), ptr(a), ptr(stats), H, W, None if as_uint8 else ptr(out),
         ptr(out) if as_uint8 else None)
    return out


def depth_range(depth, acc=None, near=None, far=None, ignore_frac=0.0):
    """pose.py:180-194: device tensor [near, far]; falsy near / far are replaced by the acc-weighted depth quantiles."""
    d = f32c(depth)
    check_cuda(d)
    a = None if acc is None else f32c(acc)
    auto_near, auto_far = not near, not far  # `near = near or ...` in the reference: 0 and None both mean automatic
    lib = _lib.load()
    work = torch.empty(lib.mip360_vis_work_len(), device=d.device, dtype=torch.int64)
    rng = torch.empty(2, device=d.device, dtype=torch.float32)
    call("mip360_depth_range", ptr(d), ptr(a), d.numel(), float(ignore_frac), float(near or 0.0), float(far or 0.0),
         int(auto_near), int(auto_far), ptr(work), ptr(rng))
    return rng


def visualize_depth(depth, acc=None, near=None, far=None, ignore_frac=0.0, curve="neg_log", modulus=0.0, lut=None,
                    as_uint8=False):
    """pose.py:149-212 on the device.  lut: [n,3] fp32 colour table (listed-colour-map indexing); None with
    modulus > 0 selects the sinebow map."""
    d = f32c(depth)
    check_cuda(d)
    a = None if acc is None else f32c(acc)
    rng = depth_range(d, a, near, far, ignore_frac)
    table = None if lut is None else f32c(lut)[:, :3].contiguous()
    out = torch.empty(tuple(d.shape) + (3,), device=d.device, dtype=torch.uint8 if as_uint8 else torch.float32)
    call("mip360_visualize_depth", ptr(d), ptr(a), ptr(rng), CURVES[curve], float(modulus), ptr(table),
         0 if table is None else table.shape[0], d.numel(), None if as_uint8 else ptr(out), ptr(out) if as_uint8 else None)
    return out
