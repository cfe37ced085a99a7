"""mipnerf360_b200 — B200-native per-ray hot path of zhangkai0425/mipnerf360.

Layout mirrors the reference so its entry points drop in:
    mipnerf360_b200.model                       <- model.py   (mipNeRF360, prop_net, nerf_net)
    mipnerf360_b200.intern.{ray,encoding,parameterization,distillation,regularization,loss,utils}
    mipnerf360_b200.ops / mlp                   host-side operators over the C ABI
    mipnerf360_b200.csrc / lib                  hand-written sm_100a kernels, libmip360_b200.so
"""
__all__ = ["ops", "mlp", "model", "intern"]
