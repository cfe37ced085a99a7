"""Mirror of intern/loss.py."""
import torch

from mipnerf360_b200 import ops


def Loss_prop(t, w, t_hat, w_hat):
    """loss.py:6-21: bounds from the (detached) fine level, loss on the coarse weights."""
    total = ops.bounds_batch_total(t, w, t_hat)
    return ops.interlevel_loss(w_hat, bound_total=total)


def Loss_nerf(input, target):
    """loss.py:23-40: mse summed over channels / batch; loss = -psnr + 30.  B x 3 elements: stays in torch
    (SURVEY §2.1 'Photometric loss')."""
    batch_size = input.shape[0]
    mse_loss = ((input[..., :3] - target[..., :3]) ** 2).sum() / batch_size
    psnr = mse_to_psnr(mse_loss)
    mse_loss = -mse_to_psnr(mse_loss) + 30
    return mse_loss, psnr


def Loss_dist(s_vals, weights):
    """loss.py:42-53."""
    return ops.distortion_loss(s_vals, weights)


def mse_to_psnr(mse):
    """loss.py:56-58."""
    return -10.0 * torch.log10(mse)
