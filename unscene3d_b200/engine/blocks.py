"""One autograd node per residual block.

The reference's BasicBlock (models/modules/resnet_block.py:24-64: conv1 - norm1 - relu - conv2 - norm2 - (+ shortcut, optionally
through a 1x1 conv + norm) - relu) runs through the module surface as ~8 nn.Module calls, 4-6 torch.autograd.Function nodes and
as many SparseTensor wrappers, although it launches only ~10 kernels; 23 blocks hold 53 of the 62 convolutions of
Res16UNet34C and the step is bound by exactly that host-side Python.  FusedBasicBlockFunction issues the same kernels in the
same order (bit-identical results) from one forward and one hand-written backward.  PyTorch still supplies memory, streams
and the autograd graph around the block.
"""
from __future__ import annotations

import torch

from .._lib import check, lib
from . import functional as Fn
from .coords import _stream


import os

_enabled = {"on": os.environ.get("US3D_FUSED_BLOCKS", "1") == "1"}


def set_fused_blocks(on: bool):
    """Switch between one autograd node per residual block (default) and the module-by-module route (same kernels)."""
    _enabled["on"] = bool(on)


class _Plan:
    """Per-call constants of a block: neighbour tables and the BatchNorm modules (running statistics)."""

    __slots__ = ("fwd1", "bwd1", "flip1", "fwd2", "bwd2", "flip2", "fwdd", "bwdd", "flipd", "norm1", "norm2", "normd")


def _w3(kernel, table):
    return kernel.detach().contiguous().view(table.kvol, kernel.shape[-2], kernel.shape[-1])


def _conv_forward(x, kernel, table, flip, bn=None):
    w3 = _w3(kernel, table)
    cin, cout = w3.shape[1], w3.shape[2]
    mode = Fn.get_precision()
    wf = Fn.packed_weights(kernel, w3, flip, mode, True)[0] if (mode != 0 and cin > 4) else None
    return Fn.spconv_gather(x, table, w3, cin, cout, False, False, None, wpack=wf, bn=bn)


def _conv_norm_stats(x, kernel, table, flip, norm):
    """y = conv(x) and the statistics norm will normalise y with: (y, mean, invstd, batch statistics?).  In training mode the
    convolution's epilogue produces them (no separate pass over y)."""
    req = norm.stats_request()
    y = _conv_forward(x, kernel, table, flip, req)
    if req is not None:
        return y, req.mean, req.invstd, True
    m, s, t = norm.statistics(y)
    return y, m, s, t


def _planes(t):
    return getattr(t, "_us3d_planes", None)


def _restore(t, planes):
    if planes is not None and planes[2] == t._version:
        t._us3d_planes = planes


class FusedBasicBlockFunction(torch.autograd.Function):
    """out = relu(bn2(conv2(relu(bn1(conv1(x))))) + shortcut(x)),  shortcut = identity or bn_d(conv_d(x))."""

    @staticmethod
    def forward(ctx, x, k1, g1, b1, k2, g2, b2, kd, gd, bd, plan: _Plan):
        x = Fn._rows(x)
        y1, m1, s1, t1 = _conv_norm_stats(x, k1, plan.fwd1, plan.flip1, plan.norm1)
        a1, g1c = Fn.bn_apply_raw(y1, g1, b1, None, m1, s1, True)
        y2, m2, s2, t2 = _conv_norm_stats(a1, k2, plan.fwd2, plan.flip2, plan.norm2)
        yd = md = sd = gdc = None
        td = False
        if kd is not None:
            yd, md, sd, td = _conv_norm_stats(x, kd, plan.fwdd, plan.flipd, plan.normd)
            res, gdc = Fn.bn_apply_raw(yd, gd, bd, None, md, sd, False)
        else:
            res = x
        out, g2c = Fn.bn_apply_raw(y2, g2, b2, res, m2, s2, True)
        ctx.save_for_backward(x, y1, a1, y2, out, yd, m1, s1, m2, s2, md, sd, g1c, g2c, gdc, k1, k2, kd)
        ctx.plan, ctx.training = plan, (bool(t1), bool(t2), bool(td))
        ctx.planes = (_planes(x), _planes(a1))  # the weight gradients re-use the forward's bf16 planes
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y1, a1, y2, out, yd, m1, s1, m2, s2, md, sd, g1c, g2c, gdc, k1, k2, kd = ctx.saved_tensors
        plan = ctx.plan
        t1, t2, td = ctx.training
        _restore(x, ctx.planes[0])
        _restore(a1, ctx.planes[1])
        dout = Fn._rows(dout)
        # norm2 (+ residual, ReLU)
        dy2, dres, dg2, db2 = Fn.bn_backward_raw(dout, y2, out, m2, s2, g2c, True, t2, True)
        w32 = _w3(k2, plan.fwd2)
        da1 = Fn.conv_input_gradient(dy2, k2, w32, plan.bwd2, plan.flip2)
        dk2 = Fn.spconv_wgrad(a1, plan.fwd2, dy2, w32.shape[1], w32.shape[2]).view(k2.shape)
        # norm1 (ReLU)
        dy1, _, dg1, db1 = Fn.bn_backward_raw(da1, y1, a1, m1, s1, g1c, True, t1, False)
        w31 = _w3(k1, plan.fwd1)
        dx = Fn.conv_input_gradient(dy1, k1, w31, plan.bwd1, plan.flip1) if ctx.needs_input_grad[0] else None
        dk1 = Fn.spconv_wgrad(x, plan.fwd1, dy1, w31.shape[1], w31.shape[2]).view(k1.shape)
        dkd = dgd = dbd = None
        if kd is not None:
            dyd, _, dgd, dbd = Fn.bn_backward_raw(dres, yd, None, md, sd, gdc, False, td, False)
            w3d = _w3(kd, plan.fwdd)
            dkd = Fn.spconv_wgrad(x, plan.fwdd, dyd, w3d.shape[1], w3d.shape[2]).view(kd.shape)
            dres = Fn.conv_input_gradient(dyd, kd, w3d, plan.bwdd, plan.flipd) if ctx.needs_input_grad[0] else None
        if dx is not None:
            check(lib.us3d_add(dx.data_ptr(), dres.data_ptr(), dx.data_ptr(), dx.numel(), _stream()))
        return dx, dk1, dg1, db1, dk2, dg2, db2, dkd, dgd, dbd, None


def fused_basic_block(block, x):
    """Runs a BasicBlock (conv1/norm1/conv2/norm2[/downsample = Sequential(conv, norm)]) as one autograd node; returns the
    output SparseTensor, or None when the block is not of that shape (the caller then takes the module-by-module route)."""
    from .tensor import MinkowskiBatchNorm, SparseTensor, _ConvBase

    if not _enabled["on"]:
        return None
    c1, n1, c2, n2 = (getattr(block, a, None) for a in ("conv1", "norm1", "conv2", "norm2"))
    if not (isinstance(c1, _ConvBase) and isinstance(c2, _ConvBase) and isinstance(n1, MinkowskiBatchNorm) and isinstance(n2, MinkowskiBatchNorm)):
        return None
    if hasattr(block, "conv3") or c1.bias is not None or c2.bias is not None or not x.F.is_cuda or x.F.dtype != torch.float32:
        return None
    ds = block.downsample
    cd = nd = None
    if ds is not None:
        if not (isinstance(ds, torch.nn.Sequential) and len(ds) == 2 and isinstance(ds[0], _ConvBase) and isinstance(ds[1], MinkowskiBatchNorm)
                and ds[0].bias is None):
            return None
        cd, nd = ds[0], ds[1]
    for n in (n1, n2, nd):
        if n is not None and (n.bn.weight is None or (n.bn.training and n.bn.track_running_stats and n.bn.momentum is None)):
            return None
    cm, key = x.coordinate_manager, x.coordinate_map_key
    plan = _Plan()
    key1, plan.fwd1, plan.bwd1, plan.flip1 = c1.tables(cm, key)
    key2, plan.fwd2, plan.bwd2, plan.flip2 = c2.tables(cm, key1)
    plan.norm1, plan.norm2, plan.normd = n1, n2, nd
    plan.fwdd = plan.bwdd = None
    plan.flipd = False
    if cd is not None:
        keyd, plan.fwdd, plan.bwdd, plan.flipd = cd.tables(cm, key)
        if keyd != key2:
            return None
    elif key2 != key or c2.kernel.shape[-1] != x.F.shape[1]:
        return None
    out = FusedBasicBlockFunction.apply(x.F, c1.kernel, n1.bn.weight, n1.bn.bias, c2.kernel, n2.bn.weight, n2.bn.bias,
                                        None if cd is None else cd.kernel, None if nd is None else nd.bn.weight,
                                        None if nd is None else nd.bn.bias, plan)
    return SparseTensor(out, coordinate_map_key=key2, coordinate_manager=cm)


class FusedConvNormReLUFunction(torch.autograd.Function):
    """out = relu(bn(conv(x))) — the stem and the strided / transposed transition layers between the stages
    (models/res16unet.py:226-289: `convXpYs2 -> bnX -> relu`, `convtrX -> bntrX -> relu`)."""

    @staticmethod
    def forward(ctx, x, k, g, b, plan):
        x = Fn._rows(x)
        fwd, bwd_getter, flip, norm = plan
        y, m, s, t = _conv_norm_stats(x, k, fwd, flip, norm)
        out, gc = Fn.bn_apply_raw(y, g, b, None, m, s, True)
        ctx.save_for_backward(x, y, out, m, s, gc, k)
        ctx.plan, ctx.training, ctx.planes = plan, bool(t), _planes(x)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y, out, m, s, gc, k = ctx.saved_tensors
        fwd, bwd_getter, flip, _ = ctx.plan
        _restore(x, ctx.planes)
        dy, _, dg, db = Fn.bn_backward_raw(Fn._rows(dout), y, out, m, s, gc, True, ctx.training, False)
        w3 = _w3(k, fwd)
        dx = Fn.conv_input_gradient(dy, k, w3, bwd_getter, flip) if ctx.needs_input_grad[0] else None
        dk = Fn.spconv_wgrad(x, fwd, dy, w3.shape[1], w3.shape[2]).view(k.shape)
        return dx, dk, dg, db, None


def fused_conv_norm_relu(conv, norm, x):
    """relu(norm(conv(x))) as one autograd node, or None when the modules are not a bias-free convolution followed by an
    affine MinkowskiBatchNorm (the caller then takes the module-by-module route)."""
    from .tensor import MinkowskiBatchNorm, SparseTensor, _ConvBase

    if not _enabled["on"] or not (isinstance(conv, _ConvBase) and isinstance(norm, MinkowskiBatchNorm)):
        return None
    bn = norm.bn
    if conv.bias is not None or bn.weight is None or (bn.training and bn.track_running_stats and bn.momentum is None):
        return None
    if not x.F.is_cuda or x.F.dtype != torch.float32:
        return None
    cm = x.coordinate_manager
    out_key, fwd, bwd_getter, flip = conv.tables(cm, x.coordinate_map_key)
    out = FusedConvNormReLUFunction.apply(x.F, conv.kernel, bn.weight, bn.bias, (fwd, bwd_getter, flip, norm))
    return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=cm)
