"""The proposal / NeRF MLPs (model.py:43-53, :131-158) as chains of tcgen05 GEMM launches.

Master weights stay fp32 nn.Parameters under the reference's state_dict keys; bf16
operand copies (zero-padded to multiples of 64, plus transposes for dgrad) are derived on the device
and cached per parameter version, so AdamW on the fp32 parameters works unchanged.
"""
from __future__ import annotations

import torch

from . import ops
from .ops import ACT_NONE, ACT_RELU, ACT_SIGMOID


def pad_width(n):
    """Layer widths the tcgen05 tiles support: 64, 128 or a multiple of 256 (zero padded)."""
    if n <= 64:
        return 64
    if n <= 128:
        return 128
    return (n + 255) // 256 * 256


class PackedMLP:
    """bf16 operand cache for a trunk of Linear layers and one (possibly merged) head.

    trunk: list of (nn.Linear, act);  heads: list of nn.Linear whose rows are stacked into one padded
    [64, K] head; head_act is the activation applied to the head outputs (Sigmoid for nerf_net's
    final_density / final_color, none for prop_net's last Linear)."""

    def __init__(self, trunk, heads, head_act):
        self.trunk = trunk
        self.heads = heads
        self.head_act = head_act
        self.n_valid = sum(h.out_features for h in heads)
        self._key = None
        self._packed = None

    def invalidate(self):
        """Force a re-cast of the bf16 operands (for optimisers that update the weights behind autograd's back)."""
        self._key = None

    def params(self):
        ps = []
        for lin, _ in self.trunk:
            ps += [lin.weight, lin.bias]
        for h in self.heads:
            ps += [h.weight, h.bias]
        return ps

    def packed(self):
        """(list of (Wb, Wt, bias), (Wb_head, Wt_head, bias_head)), refreshed when any parameter changed."""
        key = tuple((p.data_ptr(), p._version) for p in self.params())
        if key != self._key:
            layers = []
            k_pad = 64  # the encoded input rows are 64 bf16 wide (58 features + 6 zeros)
            for lin, _ in self.trunk:
                n_pad = pad_width(lin.out_features)
                Wb, Wt = ops.cast_weight(lin.weight, n_pad=n_pad, k_pad=k_pad)
                bias = torch.zeros(n_pad, device=Wb.device, dtype=torch.float32)
                bias[: lin.out_features] = lin.bias.detach()
                layers.append((Wb, Wt, bias))
                k_pad = n_pad
            Wh = torch.cat([h.weight.detach() for h in self.heads], 0)
            bh = torch.cat([h.bias.detach() for h in self.heads], 0).float()
            Wb, Wt = ops.cast_weight(Wh, n_pad=64, k_pad=k_pad)
            bias = torch.zeros(64, device=Wh.device, dtype=torch.float32)
            bias[: bh.numel()] = bh
            self._packed = (layers, (Wb, Wt, bias))
            self._key = key
        return self._packed


class _MLPFunction(torch.autograd.Function):
    """x bf16 [M,64] -> head outputs fp32 [M, n_valid]; gradients for every weight and bias (fp32).
    The input needs no gradient (nothing upstream of the encodings is trainable)."""

    @staticmethod
    def forward(ctx, x, mlp, *params):
        layers, (Wh, Wht, bh) = mlp.packed()
        need_grad = any(ctx.needs_input_grad[2:])  # grad mode is off inside forward(); autograd tells us here
        acts = [a for _, a in mlp.trunk]
        saved = [x]
        h = x
        for (Wb, _, bias), act in zip(layers, acts):
            h, _ = ops.linear_fwd(h, Wb, bias, act)
            if need_grad:
                saved.append(h)
        _, out = ops.linear_fwd(h, Wh, bh, mlp.head_act, out_f32_cols=mlp.n_valid, want_bf16=False)
        if need_grad:
            ctx.mlp = mlp
            ctx.acts = acts
            ctx.save_for_backward(out, *saved)
        return out

    @staticmethod
    def backward(ctx, g_out):
        mlp, acts = ctx.mlp, ctx.acts
        out, *saved = ctx.saved_tensors  # saved[0] = x, saved[l] = output of trunk layer l
        layers, (Wh, Wht, bh) = mlp.packed()
        L = len(layers)
        # head: dZ_head (bf16, padded to 64 columns) with the head activation derivative folded in
        dzh = ops.head_grad_pack(g_out, out if mlp.head_act == ACT_SIGMOID else None, mlp.head_act)
        dWh, dbh = ops.linear_wgrad(dzh, saved[L])
        grads_head = []
        r = 0
        for h in mlp.heads:
            n = h.out_features
            grads_head += [dWh[r:r + n, : h.in_features], dbh[r:r + n]]
            r += n
        # into the trunk: derivative of the last trunk activation from its saved output
        dz = ops.linear_dgrad(dzh, Wht, saved[L], acts[L - 1])
        grads_trunk = [None] * (2 * L)
        for l in range(L, 0, -1):  # trunk layer l maps saved[l-1] -> saved[l]
            lin = mlp.trunk[l - 1][0]
            gw, gb = lin.weight.grad, lin.bias.grad
            if (gw is not None and gb is not None and gw.is_contiguous() and gb.is_contiguous()
                    and gw.dtype == torch.float32 and tuple(gw.shape) == (dz.shape[1], saved[l - 1].shape[1])):
                # unpadded layer with preallocated .grad (e.g. the flat buffers of train.FlatAdamW): the split-K
                # kernel accumulates straight into it; returning None tells autograd there is nothing left to add
                ops.linear_wgrad(dz, saved[l - 1], dW=gw, db=gb)
            else:
                dW, db = ops.linear_wgrad(dz, saved[l - 1])
                grads_trunk[2 * (l - 1)] = dW[: lin.out_features, : lin.in_features]
                grads_trunk[2 * (l - 1) + 1] = db[: lin.out_features]
            if l > 1:
                dz = ops.linear_dgrad(dz, layers[l - 1][1], saved[l - 1], acts[l - 2])
        return (None, None, *grads_trunk, *grads_head)


def mlp_apply(mlp: PackedMLP, x):
    """Run the packed MLP on bf16 rows x [M,64]; differentiable w.r.t. the fp32 master parameters."""
    return _MLPFunction.apply(x, mlp, *mlp.params())


def pack_prop(model_seq):
    """prop_net.model (model.py:43-53): Linear/ReLU x3, Linear/Sigmoid, Linear(hidden,1)."""
    lins = [m for m in model_seq if isinstance(m, torch.nn.Linear)]
    trunk = [(lins[0], ACT_RELU), (lins[1], ACT_RELU), (lins[2], ACT_RELU), (lins[3], ACT_SIGMOID)]
    return PackedMLP(trunk, [lins[4]], ACT_NONE)


def pack_nerf(model_seq, final_density, final_color):
    """nerf_net.model / final_density / final_color (model.py:131-158)."""
    lins = [m for m in model_seq if isinstance(m, torch.nn.Linear)]
    trunk = [(l, ACT_RELU) for l in lins[:-1]] + [(lins[-1], ACT_SIGMOID)]
    return PackedMLP(trunk, [final_density[0], final_color[0]], ACT_SIGMOID)
