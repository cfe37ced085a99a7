"""Mirror of intern/regularization.py: the distortion regulariser, O(N) per ray (App. A9)."""
from mipnerf360_b200 import ops


def loss_dist(s_vals, weights):
    """regularization.py:3-19: summed over the batch, both (i,j) orders."""
    return ops.distortion_loss(s_vals, weights)
