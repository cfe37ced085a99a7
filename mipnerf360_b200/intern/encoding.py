"""Mirror of intern/encoding.py: PositionalEncoding (single-scale IPE on 21 directions, App. A3) and
ViewdirectionEncoding, as nn.Modules with the reference's constructor/forward signatures."""
import torch
import torch.nn as nn

from mipnerf360_b200 import ops


class PositionalEncoding(nn.Module):
    """encoding.py:5-61.  The 21x3 basis lives in the kernel's constant memory; `P` is kept as a plain
    attribute (not a buffer) exactly like the reference, so state_dict keys are unchanged."""

    def __init__(self):
        super().__init__()
        a, b, c, d, e = 0.8506508, 0.5257311, 0.809017, 0.5, 0.309017
        self.P = torch.tensor([[a, 0, b], [c, d, e], [b, a, 0], [1, 0, 0], [c, d, -e], [a, 0, -b], [e, c, -d],
                               [0, b, -a], [d, e, -c], [0, 1, 0], [-b, a, 0], [-e, c, -d], [0, b, a], [-e, c, d],
                               [e, c, d], [d, e, c], [d, -e, c], [0, 0, 1], [-d, e, c], [-c, d, e], [-c, d, -e]],
                              requires_grad=False)

    def forward(self, mean, cov):
        """mean [B,N,3], cov [B,N,3,3] or None -> [B,N,42] (encoding.py:33-61)."""
        return ops.ipe(mean, cov)


class ViewdirectionEncoding(nn.Module):
    """encoding.py:63-90."""

    def __init__(self, viewdir_min_deg, viewdir_max_deg):
        super().__init__()
        self.min_deg, self.max_deg = viewdir_min_deg, viewdir_max_deg
        self.scales = torch.tensor([2 ** i for i in range(viewdir_min_deg, viewdir_max_deg)], dtype=torch.float32,
                                   requires_grad=False)

    def forward(self, viewdirs):
        return ops.viewdir_enc(viewdirs, self.min_deg, self.max_deg)
