"""Mirror of intern/ray.py (hot-path functions; convert_to_ndc is dataset preprocessing, out of scope)."""
from collections import namedtuple

from mipnerf360_b200 import ops

Rays = namedtuple('Rays', ('origins', 'directions', 'viewdirs', 'radii', 'near', 'far'))  # ray.py:6


def namedtuple_map(fn, tup):
    """ray.py:8-10."""
    return type(tup)(*map(fn, tup))


def sorted_piecewise_constant_pdf(bins, weights, num_samples, randomized=True):
    """ray.py:12-57.  num_samples must be N+1 (the only value the reference passes, ray.py:146)."""
    if num_samples != weights.shape[-1] + 1:
        raise ValueError("sorted_piecewise_constant_pdf: num_samples must equal bins.shape[-1] (ray.py:146)")
    return ops.resample(bins, weights, randomized, 0.0, blur=False)


def sample_along_rays(origins, directions, radii, num_samples, near, far, randomized):
    """ray.py:81-116 -> t_vals [B,N+1], (means [B,N,3], covs [B,N,3,3])."""
    t_vals = ops.level0_t_vals(near, far, num_samples, randomized)
    out = ops.cast_ipe(t_vals, origins, directions, radii, want_means=True, want_covs=True)
    return t_vals, (out["means"], out["covs"])


def resample_along_rays(origins, directions, radii, t_vals, weights, randomized, resample_padding):
    """ray.py:118-153."""
    new_t = ops.resample(t_vals, weights, randomized, resample_padding, blur=True)
    out = ops.cast_ipe(new_t, origins, directions, radii, want_means=True, want_covs=True)
    return new_t, (out["means"], out["covs"])


def volumetric_rendering(rgb, density, t_vals, dirs, white_bkgd):
    """ray.py:155-191 -> comp_rgb [B,3], distance [B], acc [B], weights [B,N]."""
    return ops.composite(rgb, density, t_vals, dirs, white_bkgd)
