"""Mirror of intern/distillation.py: the interlevel (proposal) loss with the reference's batch-coupled,
fine->coarse bound (App. A6), as O(N) per-ray warp kernels instead of the N-iteration masked gather."""
from mipnerf360_b200 import ops


def bounds(t_vals_fine, fine_weights, t_vals_coarse):
    """distillation.py:4-33: bound for coarse interval i = total over ALL rays, broadcast to every ray; detached."""
    b = ops.bounds_per_ray(t_vals_fine, fine_weights, t_vals_coarse)
    total = ops.bounds_total(b)
    return total.float()[None, :].expand_as(b)


def loss_prop(coarse_weights, bounds):
    """distillation.py:35-51 for an explicit [B,N] bounds tensor."""
    return ops.interlevel_loss(coarse_weights, b_per_ray=ops.f32c(bounds.detach()), per_ray_bounds=True)
