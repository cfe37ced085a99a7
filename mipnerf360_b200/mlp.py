"""The proposal / NeRF MLPs (model.py:43-53, :131-158) as chains of tcgen05 GEMM launches.

Master weights stay fp32 nn.Parameters under the reference's state_dict keys; bf16
operand copies (zero-padded to multiples of 64, plus transposes for dgrad) are derived on the device
and cached per parameter version, so AdamW on the fp32 parameters works unchanged.
"""
from __future__ import annotations

import torch

import ctypes

from . import _lib, ops
from .ops import ACT_NONE, ACT_RELU, ACT_SIGMOID


def _pad64(n):
    return (n + 63) // 64 * 64


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
    final_density / final_color, none for prop_net's last Linear).

    The bf16 copies (Wb, its transpose Wt, padded fp32 bias per layer) are allocated ONCE and refreshed in place by
    one kernel launch (mip360_adamw_pack), so pointers baked into TMA descriptors, `struct mip360_layer` tables or
    captured CUDA graphs stay valid across optimiser steps."""

    def __init__(self, trunk, heads, head_act):
        self.trunk = trunk
        self.heads = heads
        self.head_act = head_act
        self.n_valid = sum(h.out_features for h in heads)
        if len(heads) > 2:
            raise ValueError("at most two head Linear modules can share the padded head")
        # Fold the head into the epilogue of the last trunk layer (mip360_linear_fwd_head): the MLP then returns the head's
        # PRE-activation sums without bias, and the consumer applies bias + head activation (ops.composite_heads with
        # head_bias=).  Set by the owner of the MLP (nerf_net) when its consumer does that.
        self.fuse_head = False
        self._key = None        # (data_ptr, version) of every parameter at the last refresh
        self._ptr_key = None    # data_ptr of every parameter the device table was built for
        self._packed = None
        # Opt-in (set by train.FlatAdamW): backward adds every dW/db straight into the preallocated .grad tensors
        # and returns None to autograd.  Off by default, so torch.autograd.grad(), tensor hooks and DDP see the
        # gradients of every layer through autograd.
        self.direct_grad = False
        # With direct_grad: called as grad_hook(l) right after the gradients of trunk layer l (and, for
        # l == len(trunk) - 1, of the heads) are complete — the data-parallel trainer all-reduces that bucket
        # on a side stream while the remaining dgrad/wgrad GEMMs run (train.py:61-64,79-82 across ranks).
        self.grad_hook = None

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

    # -- buffers and device table --------------------------------------------------------------------------------
    def _allocate(self, dev):
        layers = []
        k_pad = _pad64(self.trunk[0][0].in_features)  # the encoded input rows are 64 bf16 wide (58 features + zeros)
        for lin, _ in self.trunk:
            n_pad = pad_width(lin.out_features)
            layers.append((torch.empty((n_pad, k_pad), device=dev, dtype=torch.bfloat16),
                           torch.empty((k_pad, n_pad), device=dev, dtype=torch.bfloat16),
                           torch.empty(n_pad, device=dev, dtype=torch.float32)))
            k_pad = n_pad
        head = (torch.empty((64, k_pad), device=dev, dtype=torch.bfloat16),
                torch.empty((k_pad, 64), device=dev, dtype=torch.bfloat16),
                torch.empty(64, device=dev, dtype=torch.float32))
        self._packed = (layers, head)
        self.head_w4 = torch.zeros((k_pad, 4), device=dev, dtype=torch.float32)  # rows 0..3 of the head, column-interleaved
        acts = [a for _, a in self.trunk]
        arr = (_lib.Layer * len(layers))()
        for i, ((W_, Wt_, b_), a) in enumerate(zip(layers, acts)):
            arr[i] = _lib.Layer(W_.data_ptr(), Wt_.data_ptr(), b_.data_ptr(), W_.shape[0], W_.shape[1], a)
        Wb, Wt, bias = head
        self._cstructs = (arr, _lib.Layer(Wb.data_ptr(), Wt.data_ptr(), bias.data_ptr(), 64, Wb.shape[1], self.head_act))

    def _build_table(self):
        """struct mip360_pack_entry per layer, uploaded to the device (rebuilt only when a parameter moved)."""
        layers, head = self._packed
        groups = [[lin] for lin, _ in self.trunk] + [list(self.heads)]
        bufs = list(layers) + [head]
        entries = (_lib.PackEntry * len(bufs))()
        tile = 0
        for i, (lins, (Wb, Wt, bias)) in enumerate(zip(groups, bufs)):
            e = entries[i]
            for j, lin in enumerate(lins):
                if lin.weight.dtype != torch.float32 or not lin.weight.is_contiguous() or not lin.bias.is_contiguous():
                    raise _lib.Mip360Error("PackedMLP: fp32 contiguous master weights expected")
                e.w_src[j], e.b_src[j], e.rows[j] = lin.weight.data_ptr(), lin.bias.data_ptr(), lin.out_features
            e.K = lins[0].in_features
            e.n_pad, e.k_pad, e.tile_begin = Wb.shape[0], Wb.shape[1], tile
            e.Wb, e.Wt, e.bias = Wb.data_ptr(), Wt.data_ptr(), bias.data_ptr()
            e.w4 = self.head_w4.data_ptr() if i == len(bufs) - 1 else None
            tile += (Wb.shape[0] // 32) * (Wb.shape[1] // 32)
        raw = torch.frombuffer(bytearray(bytes(entries)), dtype=torch.uint8)
        self._table = raw.to(layers[0][0].device)
        self._n_entries, self._n_tiles = len(bufs), tile

    def table(self):
        """(device pointer of the entry table, entries, tiles) for mip360_adamw_pack."""
        ps = self.params()
        if self._packed is None:
            self._allocate(ps[0].device)
        ptr_key = tuple(p.data_ptr() for p in ps)
        if ptr_key != self._ptr_key:
            self._build_table()
            self._ptr_key = ptr_key
        return self._table.data_ptr(), self._n_entries, self._n_tiles

    def mark_fresh(self):
        """The operands were just refreshed by a fused optimiser step (mip360_adamw_pack with do_adam = 1)."""
        self._key = tuple((p.data_ptr(), p._version) for p in self.params())

    def packed(self):
        """(list of (Wb, Wt, bias), (Wb_head, Wt_head, bias_head)), refreshed when any parameter changed."""
        key = tuple((p.data_ptr(), p._version) for p in self.params())
        if key != self._key:
            tab, n_entries, n_tiles = self.table()
            _lib.call("mip360_adamw_pack", tab, n_entries, n_tiles, None, None, None, None, 0.0, 0.0, 0.0, 0.0, 0.0, 0,
                      None, 0, 0)
            self._key = key
        return self._packed

    def cstructs(self):
        self.packed()
        return self._cstructs

    def narrow_shape(self):
        """The shape mip360_mlp_fwd_fused_narrow covers: 64 -> 256 x 4 -> head without activation (the proposal net)."""
        layers, head = self.packed()
        return (len(layers) == 4 and self.head_act == ACT_NONE and layers[0][0].shape == (256, 64)
                and all(W.shape == (256, 256) for W, _, _ in layers[1:]) and head[0].shape == (64, 256))

    def head_bias(self):
        """Padded fp32 head bias [64] (first n_valid entries real), refreshed with the operands."""
        return self.packed()[1][2]


class _MLPFunction(torch.autograd.Function):
    """x bf16 [M,64] -> head outputs fp32 [M, n_valid]; gradients for every weight and bias (fp32).
    The input needs no gradient (nothing upstream of the encodings is trainable)."""

    @staticmethod
    def forward(ctx, x, mlp, need_grad, *params):
        # need_grad comes from mlp_apply: ctx.needs_input_grad does not see torch.no_grad(), so reading it here
        # would keep every activation buffer alive on inference passes
        layers, (Wh, Wht, bh) = mlp.packed()
        acts = [a for _, a in mlp.trunk]
        M, L = x.shape[0], len(layers)
        fuse = mlp.fuse_head and mlp.n_valid == 4
        ran_narrow = False
        if not fuse and mlp.narrow_shape():
            # proposal-net shape: ONE kernel for all layers, activations stay on chip (written to HBM only when the
            # backward pass needs them); falls through when the library reports the shape / option as unsupported
            trunk_arr, head = mlp.cstructs()
            out = torch.empty((M, mlp.n_valid), device=x.device, dtype=torch.float32)
            bufs = [torch.empty((M, Wb.shape[0]), device=x.device, dtype=torch.bfloat16) for Wb, _, _ in layers] \
                if need_grad else []
            ptrs = (ctypes.c_void_p * L)(*[b.data_ptr() for b in bufs]) if need_grad else None
            ran_narrow = _lib.call_rc("mip360_mlp_fwd_fused_narrow", x.data_ptr(), M, trunk_arr, L, ctypes.byref(head),
                                      mlp.n_valid, ptrs, out.data_ptr())
            if ran_narrow:
                mlp.last_n_act_bufs = len(bufs)
                saved = [x] + bufs
                if _lib.PROFILE is not None and need_grad:   # instrumented runs: tell the two variants apart
                    name, ints, e0, e1 = _lib.PROFILE[-1]
                    _lib.PROFILE[-1] = (name, ints + (1,), e0, e1)
        if ran_narrow:
            pass
        elif _lib.PROFILE is not None:
            # instrumented runs (bench.py's per-kernel table): one C call per GEMM so that each launch is timed
            saved = [x]
            h = x
            for l, ((Wb, _, bias), act) in enumerate(zip(layers, acts)):
                if fuse and l == L - 1:
                    out = torch.zeros((M, 4), device=x.device, dtype=torch.float32)
                    h = ops.linear_fwd_head(h, Wb, bias, act, mlp.head_w4, out, want_bf16=need_grad)
                else:
                    h, _ = ops.linear_fwd(h, Wb, bias, act)
                if need_grad:
                    saved.append(h)
            if not fuse:
                _, out = ops.linear_fwd(h, Wh, bh, mlp.head_act, out_f32_cols=mlp.n_valid, want_bf16=False)
        elif fuse:
            # product path with the head folded into the last trunk layer: one C call, no head GEMM, and on inference
            # passes no write of the last trunk activation either
            trunk_arr, _ = mlp.cstructs()
            if need_grad:
                bufs = [torch.empty((M, Wb.shape[0]), device=x.device, dtype=torch.bfloat16) for Wb, _, _ in layers]
            else:
                wmax = max(Wb.shape[0] for Wb, _, _ in layers)
                bufs = [torch.empty((M, wmax), device=x.device, dtype=torch.bfloat16) for _ in range(2)]
            out = torch.empty((M, 4), device=x.device, dtype=torch.float32)
            mlp.last_n_act_bufs = len(bufs)
            ptrs = (ctypes.c_void_p * len(bufs))(*[b.data_ptr() for b in bufs])
            _lib.call("mip360_mlp_fwd_fused_head", x.data_ptr(), M, trunk_arr, L, mlp.head_w4.data_ptr(), ptrs, len(bufs),
                      out.data_ptr())
            saved = [x] + bufs
        else:
            # product path: the whole MLP is one call into the C ABI
            trunk_arr, head = mlp.cstructs()
            out = torch.empty((M, mlp.n_valid), device=x.device, dtype=torch.float32)
            if need_grad:
                bufs = [torch.empty((M, Wb.shape[0]), device=x.device, dtype=torch.bfloat16) for Wb, _, _ in layers]
            else:
                wmax = max(Wb.shape[0] for Wb, _, _ in layers)
                bufs = [torch.empty((M, wmax), device=x.device, dtype=torch.bfloat16) for _ in range(2)]
            mlp.last_n_act_bufs = len(bufs)  # 2 = inference ping-pong, n_trunk = saved for backward
            ptrs = (ctypes.c_void_p * len(bufs))(*[b.data_ptr() for b in bufs])
            _lib.call("mip360_mlp_fwd", x.data_ptr(), M, trunk_arr, L, ctypes.byref(head), mlp.n_valid, ptrs, len(bufs),
                      out.data_ptr())
            saved = [x] + bufs
        if need_grad:
            ctx.mlp = mlp
            ctx.acts = acts
            ctx.fused = fuse
            ctx.save_for_backward(out, *saved)
        return out

    @staticmethod
    def backward(ctx, g_out):
        mlp, acts = ctx.mlp, ctx.acts
        out, *saved = ctx.saved_tensors  # saved[0] = x, saved[l] = output of trunk layer l
        layers, (Wh, Wht, bh) = mlp.packed()
        L = len(layers)
        M = saved[0].shape[0]
        dev = saved[0].device
        direct = mlp.direct_grad and all(p.grad is not None for p in mlp.params())

        def grad_targets(lin, n_pad, k_pad):
            """(dW, db, in_place): with direct_grad, the layer's preallocated .grad buffers when the GEMM can
            accumulate into them as they are (unpadded layer, e.g. views of train.FlatAdamW's flat buffer);
            otherwise fresh zero-initialised padded scratch."""
            gw, gb = lin.weight.grad, lin.bias.grad
            if (direct and gw.is_contiguous() and gb.is_contiguous() and gw.dtype == torch.float32
                    and tuple(gw.shape) == (n_pad, k_pad)):
                return gw, gb, True
            return (torch.zeros((n_pad, k_pad), device=dev, dtype=torch.float32),
                    torch.zeros((n_pad,), device=dev, dtype=torch.float32), False)

        targets = [grad_targets(mlp.trunk[l][0], layers[l][0].shape[0], layers[l][0].shape[1]) for l in range(L)]
        dWh = torch.zeros((64, Wh.shape[1]), device=dev, dtype=torch.float32)
        dbh = torch.zeros((64,), device=dev, dtype=torch.float32)

        def finish_layer(l):
            """direct mode: fold padded scratch into .grad, then tell the trainer the bucket is complete."""
            if not direct:
                return
            dW, db, in_place = targets[l]
            lin = mlp.trunk[l][0]
            if not in_place:
                lin.weight.grad.add_(dW[: lin.out_features, : lin.in_features])
                lin.bias.grad.add_(db[: lin.out_features])
            if l == L - 1:
                r = 0
                for h in mlp.heads:
                    n = h.out_features
                    h.weight.grad.add_(dWh[r:r + n, : h.in_features])
                    h.bias.grad.add_(dbh[r:r + n])
                    r += n
            if mlp.grad_hook is not None:
                mlp.grad_hook(l)

        # fused head: `out` are pre-activation sums and g_out is already dL/d(pre-activation) (the consumer applied the
        # head activation and its derivative), so the head gradient is packed without an activation derivative
        head_act = ACT_NONE if ctx.fused else mlp.head_act
        per_layer = _lib.PROFILE is not None or (direct and mlp.grad_hook is not None)
        if per_layer:
            # one C call per GEMM: instrumented runs (bench.py's per-kernel table) and the data-parallel trainer,
            # which starts the all-reduce of a layer's gradients as soon as its wgrad has been enqueued
            if ctx.fused:  # head dgrad + wgrad in one pass over the saved trunk output
                dz = ops.head_bwd(g_out, mlp.head_w4, saved[L], acts[L - 1], dWh, dbh)
            else:
                dzh = ops.head_grad_pack(g_out, out if head_act == ACT_SIGMOID else None, head_act)
                ops.linear_wgrad(dzh, saved[L], dW=dWh, db=dbh)
                dz = ops.linear_dgrad(dzh, Wht, saved[L], acts[L - 1])
            for l in range(L, 0, -1):  # trunk layer l-1 maps saved[l-1] -> saved[l]
                ops.linear_wgrad(dz, saved[l - 1], dW=targets[l - 1][0], db=targets[l - 1][1])
                finish_layer(l - 1)
                if l > 1:
                    dz = ops.linear_dgrad(dz, layers[l - 1][1], saved[l - 1], acts[l - 2])
        else:
            trunk_arr, head = mlp.cstructs()
            wmax = max(Wb.shape[0] for Wb, _, _ in layers)
            dzh = torch.empty((M, 64), device=dev, dtype=torch.bfloat16)
            dz0 = torch.empty((M, wmax), device=dev, dtype=torch.bfloat16)
            dz1 = torch.empty((M, wmax), device=dev, dtype=torch.bfloat16)
            act_ptrs = (ctypes.c_void_p * L)(*[a.data_ptr() for a in saved[1:]])
            dW_ptrs = (ctypes.c_void_p * (L + 1))(*([t[0].data_ptr() for t in targets] + [dWh.data_ptr()]))
            db_ptrs = (ctypes.c_void_p * (L + 1))(*([t[1].data_ptr() for t in targets] + [dbh.data_ptr()]))
            g = ops.f32c(g_out)
            if ctx.fused:  # g_out is dL/d(head pre-activation): head dgrad + wgrad in one pass, then the trunk
                _lib.call("mip360_mlp_bwd_fused_head", g.data_ptr(), saved[0].data_ptr(), M, trunk_arr, L,
                          mlp.head_w4.data_ptr(), act_ptrs, dW_ptrs, db_ptrs, dz0.data_ptr(), dz1.data_ptr())
            else:
                _lib.call("mip360_mlp_bwd", g.data_ptr(), out.data_ptr(), saved[0].data_ptr(), M, trunk_arr, L,
                          ctypes.byref(head), mlp.n_valid, act_ptrs, dW_ptrs, db_ptrs, dzh.data_ptr(), dz0.data_ptr(),
                          dz1.data_ptr())
            for l in range(L - 1, -1, -1):
                finish_layer(l)
        if direct:  # everything has been added to .grad; autograd has nothing left to accumulate
            return (None, None, None) + (None,) * (2 * L + 2 * len(mlp.heads))
        grads_trunk = []
        for (dW, db, _), (lin, _) in zip(targets, mlp.trunk):
            grads_trunk += [dW[: lin.out_features, : lin.in_features], db[: lin.out_features]]
        grads_head = []
        r = 0
        for h in mlp.heads:
            n = h.out_features
            grads_head += [dWh[r:r + n, : h.in_features], dbh[r:r + n]]
            r += n
        return (None, None, None, *grads_trunk, *grads_head)


def mlp_apply(mlp: PackedMLP, x):
    """Run the packed MLP on bf16 rows x [M,64]; differentiable w.r.t. the fp32 master parameters.
    Under torch.no_grad() (render_image, eval, the detached forwards of train.py:55,68-70) only two ping-pong
    activation buffers are allocated and nothing is saved."""
    params = mlp.params()
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    return _MLPFunction.apply(x, mlp, need_grad, *params)


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
