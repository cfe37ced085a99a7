"""Mirror of intern/parameterization.py: same names and argument meaning, bodies run on sm_100a kernels.

Differences from the reference, all deliberate: nothing is mutated in place, and the
per-sample autograd Jacobian loop (parameterization.py:77-79) is the closed form inside the kernel.
"""
import torch

from mipnerf360_b200 import ops


def t_to_s(t_vals, near, far):
    """parameterization.py:5-8.  Returns s_vals; reproduces the eps shifts one reference call observes."""
    return ops.t_to_s(t_vals, near, far)[0]


def s_to_t(s_vals, near, far):
    """parameterization.py:10-13."""
    return ops.s_to_t(s_vals, near, far)


def g(x):
    """parameterization.py:15-21: 1/(x+1e-6).  Pure (the reference adds eps to x in place); two elementwise
    ops kept in torch — inside the hot path this arithmetic lives in the sampling and t<->s kernels."""
    return 1.0 / (x + 1e-6)


def contract(x):
    """parameterization.py:23-29: Frobenius norm over the WHOLE tensor (App. A1)."""
    return ops.contract(x)


def gaussian_to_xyz(d, t_mean, t_var, r_var, diag=False):
    """parameterization.py:31-62: full covariance [B,N,3,3], or its diagonal [B,N,3] with diag=True."""
    return ops.gaussian_to_xyz(d, t_mean, t_var, r_var, diag=diag)


def gaussian_contract(mean, cov):
    """parameterization.py:64-83."""
    return ops.gaussian_contract(mean, cov)


_DIAG_MSG = ("diag=True: the reference itself fails here — gaussian_contract multiplies the [B,N,3] diagonal by the "
             "[B,N,3,3] Jacobians (parameterization.py:80-81, 'size of tensor a must match'); use gaussian_to_xyz(diag=True)")


def conical_frustum_to_gaussian(d, t0, t1, base_radius, diag, stable=True):
    """parameterization.py:85-117, contraction included; stable=False selects the original formula (:108-113)."""
    if diag:
        raise RuntimeError(_DIAG_MSG)
    out = ops.cast_ipe(None, None, d, base_radius, t0=t0, t1=t1, add_origins=False, want_means=True, want_covs=True,
                       stable=stable)
    return out["means"], out["covs"]


def para_rays(t_vals, origins, directions, radii, diag=False):
    """parameterization.py:119-136: origins are added AFTER the contraction (App. A2)."""
    if diag:
        raise RuntimeError(_DIAG_MSG)
    out = ops.cast_ipe(t_vals, origins, directions, radii, want_means=True, want_covs=True)
    return out["means"], out["covs"]
