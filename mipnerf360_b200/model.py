"""Mirror of model.py: prop_net, nerf_net and mipNeRF360 with the reference's constructor arguments,
forward signatures, return tuples, train/eval flag plumbing (App. A7) and state_dict keys,
running on the fused sm_100a path:

    level-0 sampling / resampling  ->  fused cast+contract+IPE (bf16 rows)  ->  tcgen05 MLP
    ->  fused head activations + compositing  ->  t_to_s

The fp32 nn.Linear parameters are the master weights (checkpoints round-trip with the reference);
MLP arithmetic is bf16 x bf16 -> fp32 (mlp.py).  Inputs are never mutated (App. A4).
"""
import weakref

import torch
import torch.distributed as dist
import torch.nn as nn

from mipnerf360_b200 import mlp as _mlp
from mipnerf360_b200 import ops
from mipnerf360_b200.intern.encoding import PositionalEncoding, ViewdirectionEncoding
from mipnerf360_b200.intern.ray import namedtuple_map


def _kaiming_init(model):
    """model.py:8-12."""
    for module in model.modules():
        if isinstance(module, nn.Linear):
            nn.init.kaiming_uniform_(module.weight)


_VDIR_CACHE = {}


def _viewdir_features(viewdirs_encoding, viewdirs):
    """View-direction encoding of a ray batch, computed once per batch object: the six forwards of one training
    iteration (train.py:54-71) and the two nets of one model(rays) call see the same `rays.viewdirs` tensor."""
    key = (viewdirs_encoding.min_deg, viewdirs_encoding.max_deg)
    hit = _VDIR_CACHE.get(key)
    if hit is not None and hit[0]() is viewdirs and hit[1] == viewdirs._version and not torch.cuda.is_current_stream_capturing():
        return hit[2]
    enc = viewdirs_encoding(viewdirs)
    if not torch.cuda.is_current_stream_capturing():  # tensors of a graph's private pool must not outlive a capture
        _VDIR_CACHE[key] = (weakref.ref(viewdirs), viewdirs._version, enc)
    return enc


def _encode(rays, t_vals, viewdirs_encoding, contract_mode, batch_group=None, norm_sq=None):
    """cast -> Gaussian -> contract -> IPE ++ view-direction encoding, straight to bf16 MLP rows
    (model.py:82-88 / 169-176 without materialising means, covs or the [B,N,58] fp32 tensor).

    batch_group: None (default) = this call is the whole batch, like the reference.  A torch.distributed process
    group = the call is one ray shard of a data-parallel batch: the reference's batch-global contraction norm
    (App. A1) is summed over the group so that the sharded forward equals the unsharded one.  Only
    train.Trainer sets it; model(rays) / render_image never issue a collective."""
    vd = _viewdir_features(viewdirs_encoding, rays.viewdirs)  # [B, 4 * (max_deg - min_deg)]
    # norm_sq: the batch's squared contraction norm, already accumulated by the kernel that produced t_vals
    if contract_mode == ops.CONTRACT_REFERENCE and batch_group is not None:
        if norm_sq is None:
            t = ops.f32c(t_vals)
            B, N = t.shape[0], t.shape[1] - 1
            norm_sq = ops.frustum_norm_sq(t.data_ptr(), t.data_ptr() + 4, N + 1, ops.f32c(rays.directions), B, N)
        dist.all_reduce(norm_sq, op=dist.ReduceOp.SUM, group=batch_group)
    return ops.cast_ipe(t_vals, rays.origins, rays.directions, rays.radii, vd, contract_mode=contract_mode,
                        norm_sq=norm_sq, want_x=True)["x"]


def _norm_buffer(self, device):
    """A zeroed fp64 scalar for the batch's squared contraction norm (App. A1), or None when the contraction mode does
    not need it."""
    if self.contract_mode != ops.CONTRACT_REFERENCE:
        return None
    return torch.zeros(1, device=device, dtype=torch.float64)


class prop_net(nn.Module):
    def __init__(self, randomized=False, num_samples=128, hidden_proposal=256, density_bias=-1, viewdir_min_deg=0,
                 viewdir_max_deg=4, device=torch.device("cuda")):
        super().__init__()
        self.randomized = randomized
        self.num_samples = num_samples
        self.hidden_proposal = hidden_proposal
        self.density_bias = density_bias
        self.viewdir_min_deg = viewdir_min_deg
        self.viewdir_max_deg = viewdir_max_deg
        self.device = device
        self.contract_mode = ops.CONTRACT_REFERENCE
        self.batch_group = None  # see _encode; set by train.Trainer for ray-sharded training only

        self.positional_encoding = PositionalEncoding()
        self.viewdirs_encoding = ViewdirectionEncoding(self.viewdir_min_deg, self.viewdir_max_deg)
        self.input_size = 21 * 2 + (self.viewdir_max_deg - self.viewdir_min_deg) * 2 * 2
        self.density_activation = nn.Softplus()

        # model.py:43-53 — same module order, so state_dict keys are prop_net.model.{0,2,4,6,8}.*
        self.model = nn.Sequential(
            nn.Linear(self.input_size, self.hidden_proposal), nn.ReLU(True),
            nn.Linear(self.hidden_proposal, self.hidden_proposal), nn.ReLU(True),
            nn.Linear(self.hidden_proposal, self.hidden_proposal), nn.ReLU(True),
            nn.Linear(self.hidden_proposal, self.hidden_proposal), nn.Sigmoid(),
            nn.Linear(self.hidden_proposal, 1))
        _kaiming_init(self)
        self.to(device)
        self._packed = _mlp.pack_prop(self.model)

    _norm_buffer = _norm_buffer

    def density_to_weight(self, t_vals, density, dirs):
        """model.py:59-78."""
        return ops.density_to_weight(t_vals, density, dirs)

    def forward(self, rays):
        """model.py:80-94 -> (t_vals [B,N+1], weights [B,N])."""
        B = rays.origins.shape[0]
        norm_sq = self._norm_buffer(rays.near.device)
        t_vals = ops.level0_t_vals(rays.near, rays.far, self.num_samples, self.randomized, directions=rays.directions,
                                   norm_sq=norm_sq)  # sampling, its random draw and the contraction norm: one launch
        x = _encode(rays, t_vals, self.viewdirs_encoding, self.contract_mode, self.batch_group, norm_sq)
        raw = _mlp.mlp_apply(self._packed, x)  # [B*N, 1] logits
        weights = ops.density_to_weight(t_vals, raw.view(B, self.num_samples), rays.directions, raw_logits=True,
                                        density_bias=self.density_bias)
        return t_vals, weights


class nerf_net(nn.Module):
    def __init__(self, randomized=False, num_samples=128, hidden_nerf=1024, density_bias=-1, rgb_padding=0.001,
                 resample_padding=0.01, white_bkgd=False, viewdir_min_deg=0, viewdir_max_deg=4,
                 device=torch.device("cuda")):
        super().__init__()
        self.randomized = randomized
        self.num_samples = num_samples
        self.hidden_nerf = hidden_nerf
        self.density_bias = density_bias
        self.rgb_padding = rgb_padding
        self.resample_padding = resample_padding
        self.white_bkgd = white_bkgd
        self.viewdir_min_deg = viewdir_min_deg
        self.viewdir_max_deg = viewdir_max_deg
        self.device = device
        self.contract_mode = ops.CONTRACT_REFERENCE
        self.batch_group = None  # see _encode

        self.positional_encoding = PositionalEncoding()
        self.viewdirs_encoding = ViewdirectionEncoding(self.viewdir_min_deg, self.viewdir_max_deg)
        self.input_size = 21 * 2 + (self.viewdir_max_deg - self.viewdir_min_deg) * 2 * 2
        self.density_activation = nn.Softplus()

        # model.py:131-158 — keys nerf_net.model.{0,2,...,14}.*, final_density.0.*, final_color.0.*
        layers = [nn.Linear(self.input_size, self.hidden_nerf), nn.ReLU(True)]
        for _ in range(6):
            layers += [nn.Linear(self.hidden_nerf, self.hidden_nerf), nn.ReLU(True)]
        layers += [nn.Linear(self.hidden_nerf, self.hidden_nerf), nn.Sigmoid()]
        self.model = nn.Sequential(*layers)
        self.final_density = nn.Sequential(nn.Linear(self.hidden_nerf, 1), nn.Sigmoid())
        self.final_color = nn.Sequential(nn.Linear(self.hidden_nerf, 3), nn.Sigmoid())
        _kaiming_init(self)
        self.to(device)
        self._packed = _mlp.pack_nerf(self.model, self.final_density, self.final_color)
        # final_density / final_color (model.py:150-158) ride in the epilogue of the last trunk GEMM; their bias and
        # Sigmoid are applied by the compositing kernel
        self._packed.fuse_head = True

    _norm_buffer = _norm_buffer

    def forward(self, rays, t_vals, coarse_weights):
        """model.py:163-200 -> (rgb [B,3], dist [B], acc [B], t_vals [B,N+1], weights [B,N], s_vals [B,N+1])."""
        B = rays.origins.shape[0]
        norm_sq = self._norm_buffer(rays.near.device)
        new_t = ops.resample(t_vals, coarse_weights, self.randomized, self.resample_padding, directions=rays.directions,
                             norm_sq=norm_sq)
        N = new_t.shape[1] - 1
        x = _encode(rays, new_t, self.viewdirs_encoding, self.contract_mode, self.batch_group, norm_sq)
        raw = _mlp.mlp_apply(self._packed, x)  # [B*N, 4] = (density head, colour head)
        # heads' bias + Sigmoid (fused-head MLP: raw holds pre-activation sums), compositing and model.py:196's t_to_s
        # in one launch
        head_bias = self._packed.head_bias() if self._packed.fuse_head else None
        comp_rgb, distance, acc, weights, s_vals, t_shift = ops.composite_heads(
            raw.view(B, N, 4), new_t, rays.directions, self.density_bias, self.rgb_padding, self.white_bkgd,
            near=rays.near, far=rays.far, head_bias=head_bias)
        # model.py:193-196: stashed for the distillation / regularisation losses; the reference's t_vals comes
        # back shifted by +1e-6 because t_to_s -> g() adds eps in place (App. A4)
        self.fine_weights = weights
        self.t_vals = t_shift
        self.s_vals = s_vals
        return comp_rgb, distance, acc, self.t_vals, self.fine_weights, self.s_vals


class mipNeRF360(nn.Module):
    def __init__(self, randomized=False, num_samples=128, hidden_proposal=256, hidden_nerf=1024, density_bias=-1,
                 rgb_padding=0.001, resample_padding=0.01, white_bkgd=False, viewdir_min_deg=0, viewdir_max_deg=4,
                 device=torch.device("cuda")):
        super().__init__()
        self.randomized = randomized
        self.num_samples = num_samples
        self.hidden_proposal = hidden_proposal
        self.hidden_nerf = hidden_nerf
        self.density_bias = density_bias
        self.rgb_padding = rgb_padding
        self.resample_padding = resample_padding
        self.white_bkgd = white_bkgd
        self.viewdir_min_deg = viewdir_min_deg
        self.viewdir_max_deg = viewdir_max_deg
        self.device = device
        self.init_randomized = randomized

        self.prop_net = prop_net(randomized=self.randomized, num_samples=self.num_samples,
                                 hidden_proposal=self.hidden_proposal, density_bias=self.density_bias,
                                 viewdir_min_deg=self.viewdir_min_deg, viewdir_max_deg=self.viewdir_max_deg,
                                 device=self.device)
        self.nerf_net = nerf_net(randomized=self.randomized, num_samples=self.num_samples,
                                 hidden_nerf=self.hidden_nerf, density_bias=self.density_bias,
                                 rgb_padding=self.rgb_padding, resample_padding=self.resample_padding,
                                 white_bkgd=self.white_bkgd, viewdir_min_deg=self.viewdir_min_deg,
                                 viewdir_max_deg=self.viewdir_max_deg, device=self.device)
        self.to(device)

    def forward(self, rays):
        """model.py:247-252."""
        t_hat, w_hat = self.prop_net.forward(rays)
        final_rgbs, final_dist, final_accs, _, _, _ = self.nerf_net.forward(rays, t_vals=t_hat, coarse_weights=w_hat)
        return final_rgbs, final_dist, final_accs

    def render_image(self, rays, height, width, chunks=4096):
        """model.py:254-274: chunked inference.  Results stay on the device until the end (one D2H copy per
        output instead of three per chunk plus a print)."""
        from mipnerf360_b200.render import render_rays  # chunk loop (one CUDA-graph replay per full chunk)
        rgb, dists, accs = render_rays(self, rays, chunks)
        rgbs = ops.to8b(rgb.reshape(height, width, 3)).cpu().numpy()  # 3 B/pixel over PCIe
        return rgbs, dists.reshape(height, width).cpu().numpy(), accs.reshape(height, width).cpu().numpy()

    def train(self, mode=True):
        """model.py:276-279."""
        self.randomized = self.init_randomized
        super().train(mode)
        return self

    def eval(self):
        """model.py:281-283 (the sub-nets keep their own flags, App. A7)."""
        self.randomized = False
        return super().eval()
