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

import numpy as np

from .._lib import OP_BN_APPLY, OP_BN_BACKWARD, OP_CONV, OP_WGRAD, check, lib
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
    """The kernel as a contiguous [kvol, cin, cout] view, kept on the parameter while its storage is unchanged (detach +
    contiguous + view are three tensor ops, 372 times per step)."""
    c = getattr(kernel, "_us3d_w3", None)
    ptr = kernel.data_ptr()
    if c is not None and c[0] == ptr and c[1] == table.kvol:
        return c[2]
    w3 = kernel.detach().contiguous().view(table.kvol, kernel.shape[-2], kernel.shape[-1])
    if w3.data_ptr() == ptr:  # a real view of the parameter (not a contiguous copy): safe to keep
        try:
            kernel._us3d_w3 = (ptr, table.kvol, w3)
        except Exception:  # pragma: no cover
            pass
    return w3


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



# ---------------------------------------------------------------------------------------------------------------------------
# Launch lists: the launches of a block's forward / backward pass resolved into us3d_op_t records and issued by ONE C call
# (us3d_run_ops, csrc/executor.cu).  Same kernels, same arguments, same order as the call-by-call route below — which stays
# the fallback for everything a list does not cover (exact-fp32 mode, shapes off the tensor-core path, gradients wider than 256
# channels) and the reference for the bit-identity test (tests/test_gpu_ops.py::test_launch_lists_equal_the_call_by_call_route).
_lists = {"on": os.environ.get("US3D_LAUNCH_LISTS", "1") == "1"}
_Z = (0,) * 16


class _Ops:
    """A launch list under assembly: flat int64 records [kind, p[16], v[10]] + two floats per op (us3d_run_ops_flat) — Python ints
    appended to a list and converted ONCE (a ctypes structure per op costs as much host time as the call it replaces)."""

    __slots__ = ("a", "f", "n", "keep")

    def __init__(self):
        self.a, self.f, self.n = [], [], 0
        # Every tensor an op of the list reads or writes stays alive until the list has been LAUNCHED: a temporary dropped during
        # assembly would hand its memory back to the allocator, and a kernel launched right away into the recycled block (a lazy
        # weight pack, the zero fill of a new gradient arena, a plane split) would be overwritten by the deferred op later.
        self.keep = []

    def hold(self, *tensors):
        self.keep.extend(tensors)
        return tensors[0]

    def emit(self, kind, p, v, f0=0.0, f1=0.0):
        a = self.a
        a.append(kind)
        a.extend(p)
        a.extend(_Z[:16 - len(p)])
        a.extend(v)
        a.extend(_Z[:10 - len(v)])
        self.f.append(f0)
        self.f.append(f1)
        self.n += 1


def set_launch_lists(on: bool):
    _lists["on"] = bool(on)


class _Fallback(Exception):
    """Raised while a list is being assembled (nothing launched yet except plane splits of existing tensors): take the other route."""


_pending_meta = []  # timer records of the list being assembled (bench.py's roofline leg pairs them with the library's events)


def _run(ops):
    t = Fn._timer
    if t is not None:
        t.meta.extend(_pending_meta)
    _pending_meta.clear()
    a, f = np.array(ops.a, dtype=np.int64), np.array(ops.f, dtype=np.float64)
    check(lib.us3d_run_ops_flat(a.ctypes.data, f.ctypes.data, ops.n, _stream()))


def _meta(kind, n_in, table, cin, cout, path):
    if Fn._timer is not None:
        _pending_meta.append((kind, n_in, table.n_rows, table.kvol, cin, cout, table.pairs(), path))


def _planes_of(t, mode):
    """(hi, lo) of an EXISTING tensor (cached, or split now — the tensor is complete in stream order)."""
    if t.data_ptr() % 16 != 0 or Fn._ld(t) % 4 != 0 or t.shape[1] % 8 != 0:
        raise _Fallback
    return Fn.bf16_planes(t, mode == 3)


def _new_rows(ops, n, c, dev, mode):
    """fp32 rows + the bf16 planes the pass that fills them will also write (attached like bn_apply_raw does)."""
    y = torch.empty((n, c), dtype=torch.float32, device=dev)
    if not (Fn._want_planes(c) and n > 0):
        raise _Fallback
    hi = torch.empty((n, c), dtype=torch.bfloat16, device=dev)
    lo = torch.empty((n, c), dtype=torch.bfloat16, device=dev) if mode == 3 else None
    y._us3d_planes = (hi, lo, y._version)
    ops.hold(y, hi, lo)
    return y, hi, lo


def _op_conv(ops, kind, planes, n_in, table, wimg, cin, cout, mode, dev, req, into=None):
    """y = gather-conv over `table` of the tensor whose planes are given (`into`: y = into += ...); returns y (appends the op, the
    timer record)."""
    if wimg is None or not Fn._tc_ok(cin, cout) or cin <= 4:
        raise _Fallback
    hi, lo = planes
    y = into if into is not None else torch.empty((table.n_rows, cout), dtype=torch.float32, device=dev)
    ops.hold(y, hi, lo, wimg)
    nbr, mask, order = table.ordered() or (table.nbr, table.mask, None)
    part = table.partition() if order is not None else None
    ws, ws_bytes = Fn._conv_workspace(dev, table.n_rows, table.kvol, cout) if order is None else (None, 0)
    ops.hold(ws)  # the per-device scratch is REPLACED when a later op needs more: this op keeps the buffer it was given
    bn = [0, 0, 0, 0, 0, 0]
    eps = mom = 0.0
    if req is not None and table.n_rows > 0:
        stats = ops.hold(torch.empty((2, cout), dtype=torch.float32, device=dev))
        sp = stats.data_ptr()
        bn = [Fn._bn_workspace(dev, cout).data_ptr(), sp, sp + 4 * cout, Fn._ptr(req.running_mean), Fn._ptr(req.running_var),
              Fn._ptr(req.num_batches_tracked)]
        eps, mom = float(req.eps), float(req.momentum if req.momentum is not None else 0.0)
        req.mean, req.invstd = stats[0], stats[1]
    _meta(kind, n_in, table, cin, cout, "mt")
    ops.emit(OP_CONV, (hi.data_ptr(), Fn._ptr(lo), nbr.data_ptr(), wimg.data_ptr(), 0, Fn._ptr(order), y.data_ptr(), Fn._ptr(mask),
                       Fn._ptr(part), Fn._ptr(ws), *bn),
             (n_in, table.n_rows, table.kvol, cin, cout, mode, cout, 0 if into is None else 1, ws_bytes), eps, mom)
    return y


def _stats(ops, kind, planes, n_in, table, kernel, flip, norm, mode, dev):
    """conv + the statistics its norm will use: (y, mean, invstd, batch statistics?, w3)."""
    w3 = _w3(kernel, table)
    cin, cout = w3.shape[1], w3.shape[2]
    wimg = Fn.packed_weights(kernel, w3, flip, mode, True)[0] if cin > 4 else None
    req = norm.stats_request()
    y = _op_conv(ops, kind, planes, n_in, table, wimg, cin, cout, mode, dev, req)
    if req is not None:
        if req.mean is None:
            raise _Fallback
        return y, req.mean, req.invstd, True
    m, s, t = norm.statistics(y)  # evaluation mode: running statistics, independent of y
    if t:
        raise _Fallback
    return y, m, s, False


def _op_bn_apply(ops, x, mean, invstd, gamma, beta, residual, relu, mode):
    n, c = x.shape
    g, b = Fn._detached(gamma), Fn._detached(beta)
    y, hi, lo = _new_rows(ops, n, c, x.device, mode)
    ops.hold(x, mean, invstd, g, b, residual)
    ops.emit(OP_BN_APPLY, (x.data_ptr(), mean.data_ptr(), invstd.data_ptr(), g.data_ptr(), b.data_ptr(), Fn._ptr(residual), y.data_ptr(),
                           hi.data_ptr(), Fn._ptr(lo)),
             (c, n, c, 0 if residual is None else Fn._ld(residual), int(relu), c))
    return y, g


def _op_bn_backward(ops, dy, x, y, mean, invstd, g, relu, training, has_res, mode):
    n, c = x.shape
    dev = x.device
    yy = y if y is not None else x
    if dy.data_ptr() % 16 != 0 or Fn._ld(dy) % 4 != 0 or x.data_ptr() % 16 != 0 or Fn._ld(x) % 4 != 0:
        raise _Fallback
    dx, hi, lo = _new_rows(ops, n, c, dev, mode)
    dres = torch.empty((n, c), dtype=torch.float32, device=dev) if has_res else None
    dgb = torch.empty((2, c), dtype=torch.float32, device=dev)
    ops.hold(dres, dgb, dy, x, yy, mean, invstd, g)
    gp = dgb.data_ptr()
    ops.emit(OP_BN_BACKWARD, (dy.data_ptr(), x.data_ptr(), yy.data_ptr(), mean.data_ptr(), invstd.data_ptr(), g.data_ptr(),
                              Fn._bn_workspace(dev, c).data_ptr(), dx.data_ptr(), Fn._ptr(dres), gp, gp + 4 * c, hi.data_ptr(), Fn._ptr(lo)),
             (Fn._ld(dy), Fn._ld(x), Fn._ld(yy), n, c, int(relu), int(training), c, c))
    return dx, dres, dgb[0], dgb[1]


def _op_dgrad(ops, dy, kernel, w3, bwd_getter, flip_dgrad, mode, into=None):
    cin, cout = w3.shape[1], w3.shape[2]
    bwd, flip = bwd_getter()
    if flip != flip_dgrad or cin <= 4:
        raise _Fallback
    chunks = Fn.packed_weights(kernel, w3, flip, mode, True)[1]
    if chunks is None or len(chunks) != 1:
        raise _Fallback
    if into is not None and (into.shape != (bwd.n_rows, cin) or not into.is_contiguous()):
        raise _Fallback
    return _op_conv(ops, "dgrad", dy._us3d_planes[:2], dy.shape[0], bwd, chunks[0][2], cout, cin, mode, dy.device, None, into)


def _op_wgrad(ops, x_planes, n_in, table, dy, cin, cout, mode, kshape):
    if not Fn._wgrad_tc_ok(cin, cout) or cin == 3 or Fn._wgrad_order["mode"] == "permute":
        raise _Fallback
    dw = Fn._zero_arena.take(table.kvol * cin * cout, dy.device)
    nbr, mask, order = (table.ordered() if table.kvol <= 8 else None) or (table.nbr, table.mask, None)
    dh, dl = dy._us3d_planes[:2]
    ops.hold(dw, dh, dl, x_planes[0], x_planes[1])
    _meta("wgrad", n_in, table, cin, cout, "wgrad-tc")
    ops.emit(OP_WGRAD, (x_planes[0].data_ptr(), Fn._ptr(x_planes[1]), dh.data_ptr(), Fn._ptr(dl), nbr.data_ptr(), dw.data_ptr(), Fn._ptr(mask),
                        Fn._ptr(order)),
             (table.n_rows, table.kvol, cin, cout, mode))
    return dw.view(kshape)


class FusedBasicBlockFunction(torch.autograd.Function):
    """out = relu(bn2(conv2(relu(bn1(conv1(x))))) + shortcut(x)),  shortcut = identity or bn_d(conv_d(x))."""

    @staticmethod
    def forward(ctx, x, k1, g1, b1, k2, g2, b2, kd, gd, bd, plan: _Plan):
        x = Fn._rows(x)
        res = None
        mode = Fn.get_precision()
        if _lists["on"] and mode != 0 and Fn._bn_fuse["on"]:
            try:
                res = FusedBasicBlockFunction._forward_list(x, k1, g1, b1, k2, g2, b2, kd, gd, bd, plan, mode)
            except _Fallback:
                _pending_meta.clear()
                res = None
        if res is None:
            res = FusedBasicBlockFunction._forward_calls(x, k1, g1, b1, k2, g2, b2, kd, gd, bd, plan)
        y1, a1, y2, out, yd, m1, s1, m2, s2, md, sd, g1c, g2c, gdc, t1, t2, td = res
        ctx.save_for_backward(x, y1, a1, y2, out, yd, m1, s1, m2, s2, md, sd, g1c, g2c, gdc, k1, k2, kd)
        ctx.plan, ctx.training = plan, (bool(t1), bool(t2), bool(td))
        ctx.planes = (_planes(x), _planes(a1))  # the weight gradients re-use the forward's bf16 planes
        return out

    @staticmethod
    def _forward_list(x, k1, g1, b1, k2, g2, b2, kd, gd, bd, plan, mode):
        ops = _Ops()
        res = FusedBasicBlockFunction._forward_ops(ops, x, k1, g1, b1, k2, g2, b2, kd, gd, bd, plan, mode)
        _run(ops)
        return res

    @staticmethod
    def _forward_ops(ops, x, k1, g1, b1, k2, g2, b2, kd, gd, bd, plan, mode):
        """Appends the block's forward launches to `ops` (nothing is launched here); returns what forward saves."""
        dev, n_in = x.device, x.shape[0]
        xp = _planes_of(x, mode)
        y1, m1, s1, t1 = _stats(ops, "fwd", xp, n_in, plan.fwd1, k1, plan.flip1, plan.norm1, mode, dev)
        a1, g1c = _op_bn_apply(ops, y1, m1, s1, g1, b1, None, True, mode)
        y2, m2, s2, t2 = _stats(ops, "fwd", a1._us3d_planes[:2], a1.shape[0], plan.fwd2, k2, plan.flip2, plan.norm2, mode, dev)
        yd = md = sd = gdc = None
        td = False
        if kd is not None:
            yd, md, sd, td = _stats(ops, "fwd", xp, n_in, plan.fwdd, kd, plan.flipd, plan.normd, mode, dev)
            res, gdc = _op_bn_apply(ops, yd, md, sd, gd, bd, None, False, mode)
        else:
            res = x
        out, g2c = _op_bn_apply(ops, y2, m2, s2, g2, b2, res, True, mode)
        return y1, a1, y2, out, yd, m1, s1, m2, s2, md, sd, g1c, g2c, gdc, t1, t2, td

    @staticmethod
    def _forward_calls(x, k1, g1, b1, k2, g2, b2, kd, gd, bd, plan):
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
        return y1, a1, y2, out, yd, m1, s1, m2, s2, md, sd, g1c, g2c, gdc, t1, t2, td

    @staticmethod
    def backward(ctx, dout):
        x, y1, a1, y2, out, yd, m1, s1, m2, s2, md, sd, g1c, g2c, gdc, k1, k2, kd = ctx.saved_tensors
        plan = ctx.plan
        _restore(x, ctx.planes[0])
        _restore(a1, ctx.planes[1])
        dout = Fn._rows(dout)
        mode = Fn.get_precision()
        if _lists["on"] and mode != 0:
            try:
                return FusedBasicBlockFunction._backward_list(ctx, dout, mode)
            except _Fallback:
                _pending_meta.clear()
        return FusedBasicBlockFunction._backward_calls(ctx, dout)

    @staticmethod
    def _backward_list(ctx, dout, mode):
        ops = _Ops()
        grads = FusedBasicBlockFunction._backward_ops(ops, ctx.saved_tensors, ctx.plan, ctx.training, dout, ctx.needs_input_grad[0], mode)
        _run(ops)
        return grads + (None,)

    @staticmethod
    def _backward_ops(ops, saved, plan, training, dout, need_dx, mode):
        """Appends the block's backward launches to `ops`; returns (dx, dk1, dg1, db1, dk2, dg2, db2, dkd, dgd, dbd)."""
        x, y1, a1, y2, out, yd, m1, s1, m2, s2, md, sd, g1c, g2c, gdc, k1, k2, kd = saved
        t1, t2, td = training
        xp, a1p = _planes_of(x, mode), _planes_of(a1, mode)
        # norm2 (+ residual, ReLU) -> conv2
        dy2, dres, dg2, db2 = _op_bn_backward(ops, dout, y2, out, m2, s2, g2c, True, t2, True, mode)
        w32 = _w3(k2, plan.fwd2)
        da1 = _op_dgrad(ops, dy2, k2, w32, plan.bwd2, plan.flip2, mode)
        dk2 = _op_wgrad(ops, a1p, a1.shape[0], plan.fwd2, dy2, w32.shape[1], w32.shape[2], mode, k2.shape)
        # norm1 (ReLU) -> conv1
        dy1, _, dg1, db1 = _op_bn_backward(ops, da1, y1, a1, m1, s1, g1c, True, t1, False, mode)
        w31 = _w3(k1, plan.fwd1)
        # dx = dX(conv1) + gradient of the shortcut: the second term is accumulated by the gather kernel's epilogue (y += ...), not
        # by a separate pass over both tensors — identity shortcut: into the residual gradient norm2's backward has written;
        # convolutional shortcut: its input gradient lands on top of conv1's
        dx = (_op_dgrad(ops, dy1, k1, w31, plan.bwd1, plan.flip1, mode, into=dres if kd is None else None)) if need_dx else None
        dk1 = _op_wgrad(ops, xp, x.shape[0], plan.fwd1, dy1, w31.shape[1], w31.shape[2], mode, k1.shape)
        dkd = dgd = dbd = None
        if kd is not None:
            dyd, _, dgd, dbd = _op_bn_backward(ops, dres, yd, None, md, sd, gdc, False, td, False, mode)
            w3d = _w3(kd, plan.fwdd)
            dkd = _op_wgrad(ops, xp, x.shape[0], plan.fwdd, dyd, w3d.shape[1], w3d.shape[2], mode, kd.shape)
            if need_dx:
                _op_dgrad(ops, dyd, kd, w3d, plan.bwdd, plan.flipd, mode, into=dx)
        return dx, dk1, dg1, db1, dk2, dg2, db2, dkd, dgd, dbd

    @staticmethod
    def _backward_calls(ctx, dout):
        return FusedBasicBlockFunction._backward_calls_of(ctx.saved_tensors, ctx.plan, ctx.training, dout, ctx.needs_input_grad[0]) + (None,)

    @staticmethod
    def _backward_calls_of(saved, plan, training, dout, need_dx):
        x, y1, a1, y2, out, yd, m1, s1, m2, s2, md, sd, g1c, g2c, gdc, k1, k2, kd = saved
        t1, t2, td = training
        # norm2 (+ residual, ReLU)
        dy2, dres, dg2, db2 = Fn.bn_backward_raw(dout, y2, out, m2, s2, g2c, True, t2, True)
        w32 = _w3(k2, plan.fwd2)
        da1 = Fn.conv_input_gradient(dy2, k2, w32, plan.bwd2, plan.flip2)
        dk2 = Fn.spconv_wgrad(a1, plan.fwd2, dy2, w32.shape[1], w32.shape[2]).view(k2.shape)
        # norm1 (ReLU)
        dy1, _, dg1, db1 = Fn.bn_backward_raw(da1, y1, a1, m1, s1, g1c, True, t1, False)
        w31 = _w3(k1, plan.fwd1)
        dx = Fn.conv_input_gradient(dy1, k1, w31, plan.bwd1, plan.flip1) if need_dx else None
        dk1 = Fn.spconv_wgrad(x, plan.fwd1, dy1, w31.shape[1], w31.shape[2]).view(k1.shape)
        dkd = dgd = dbd = None
        if kd is not None:
            dyd, _, dgd, dbd = Fn.bn_backward_raw(dres, yd, None, md, sd, gdc, False, td, False)
            w3d = _w3(kd, plan.fwdd)
            dkd = Fn.spconv_wgrad(x, plan.fwdd, dyd, w3d.shape[1], w3d.shape[2]).view(kd.shape)
            dres = Fn.conv_input_gradient(dyd, kd, w3d, plan.bwdd, plan.flipd) if need_dx else None
        if dx is not None:
            check(lib.us3d_add(dx.data_ptr(), dres.data_ptr(), dx.data_ptr(), dx.numel(), _stream()))
        return dx, dk1, dg1, db1, dk2, dg2, db2, dkd, dgd, dbd


def _block_plan(block, cm, key, cin):
    """(plan, the 9 parameters of FusedBasicBlockFunction, output key) of a BasicBlock applied to a [*, cin] tensor on map `key`,
    or None when the block is not of the fusable shape."""
    from .tensor import MinkowskiBatchNorm, _ConvBase

    c1, n1, c2, n2 = (getattr(block, a, None) for a in ("conv1", "norm1", "conv2", "norm2"))
    if not (isinstance(c1, _ConvBase) and isinstance(c2, _ConvBase) and isinstance(n1, MinkowskiBatchNorm) and isinstance(n2, MinkowskiBatchNorm)):
        return None
    if hasattr(block, "conv3") or c1.bias is not None or c2.bias is not None:
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
    elif key2 != key or c2.kernel.shape[-1] != cin:
        return None
    params = (c1.kernel, n1.bn.weight, n1.bn.bias, c2.kernel, n2.bn.weight, n2.bn.bias, None if cd is None else cd.kernel,
              None if nd is None else nd.bn.weight, None if nd is None else nd.bn.bias)
    return plan, params, key2


def fused_basic_block(block, x):
    """Runs a BasicBlock (conv1/norm1/conv2/norm2[/downsample = Sequential(conv, norm)]) as one autograd node; returns the
    output SparseTensor, or None when the block is not of that shape (the caller then takes the module-by-module route)."""
    from .tensor import SparseTensor

    if not _enabled["on"] or not x.F.is_cuda or x.F.dtype != torch.float32:
        return None
    cm = x.coordinate_manager
    bp = _block_plan(block, cm, x.coordinate_map_key, x.F.shape[1])
    if bp is None:
        return None
    plan, params, key2 = bp
    out = FusedBasicBlockFunction.apply(x.F, *params, plan)
    return SparseTensor(out, coordinate_map_key=key2, coordinate_manager=cm)


_N_SAVED = 18  # tensors FusedBasicBlockFunction saves per block
_BIG_MAP_BYTES = 192 << 20


def _big_map(t):
    """A list allocates every tensor of its ops before the first of them runs and keeps them until it is launched, so a deferred
    STAGE has the temporaries of all its blocks alive at once.  On maps whose row tensors are hundreds of MB that costs more than the
    saved host time (measured: the 4 x 200k-voxel Mask3D step 103 -> 122-142 ms with stage-wide lists at the finest level, where a
    tensor is 300-400 MB and the step is device-bound anyway; the 200k-voxel backbone step 14.3 -> 13.9 ms): there the stage is
    launched block by block."""
    return t.shape[0] * t.shape[1] * 4 > _BIG_MAP_BYTES


class FusedStageFunction(torch.autograd.Function):
    """The consecutive BasicBlocks of a stage (models/res16unet.py:224-297: block1 .. block8 are Sequentials of 2-6 blocks) as ONE
    autograd node with ONE launch list each way: the same launches as block by block, 8 nodes and 16 C calls per step instead of
    23 and 46.  Falls back block by block (call-by-call route) when a list cannot be assembled."""

    @staticmethod
    def forward(ctx, x, plans, *params):
        x = Fn._rows(x)
        mode = Fn.get_precision()
        nb = len(plans)
        results, done = [], 0  # done: blocks whose launches have been issued (their BatchNorm buffers are updated: never redo them)
        if _lists["on"] and mode != 0 and Fn._bn_fuse["on"]:
            try:
                ops, cur = _Ops(), x
                big = _big_map(x)
                for i in range(nb):
                    r = FusedBasicBlockFunction._forward_ops(ops, cur, *params[9 * i:9 * i + 9], plans[i], mode)
                    results.append((cur,) + r)
                    cur = r[3]
                    if big:  # one list per block: see _big_map
                        _run(ops)
                        ops, done = _Ops(), i + 1
                if ops.n:
                    _run(ops)
                done = nb
            except _Fallback:
                _pending_meta.clear()
        del results[done:]  # blocks assembled but not launched
        cur = results[-1][4] if results else x
        for i in range(done, nb):
            r = FusedBasicBlockFunction._forward_calls(cur, *params[9 * i:9 * i + 9], plans[i])
            results.append((cur,) + r)
            cur = r[3]
        saved, trainings, planes = [], [], []
        for i, (xi, y1, a1, y2, out, yd, m1, s1, m2, s2, md, sd, g1c, g2c, gdc, t1, t2, td) in enumerate(results):
            k1, k2, kd = params[9 * i], params[9 * i + 3], params[9 * i + 6]
            saved += [xi, y1, a1, y2, out, yd, m1, s1, m2, s2, md, sd, g1c, g2c, gdc, k1, k2, kd]
            trainings.append((bool(t1), bool(t2), bool(td)))
            planes.append((_planes(xi), _planes(a1)))
        ctx.save_for_backward(*saved)
        ctx.plans, ctx.trainings, ctx.planes = plans, trainings, planes
        return results[-1][4]

    @staticmethod
    def backward(ctx, dout):
        saved, plans = ctx.saved_tensors, ctx.plans
        nb = len(plans)
        per = [saved[_N_SAVED * i:_N_SAVED * (i + 1)] for i in range(nb)]
        for i in range(nb):
            _restore(per[i][0], ctx.planes[i][0])
            _restore(per[i][2], ctx.planes[i][1])
        dout = Fn._rows(dout)
        mode = Fn.get_precision()
        grads, todo = [None] * nb, nb  # todo: blocks [0, todo) still have to run (the list route works from the last block down)
        if _lists["on"] and mode != 0:
            try:
                ops, cur = _Ops(), dout
                big = _big_map(dout)
                for i in reversed(range(nb)):
                    g = FusedBasicBlockFunction._backward_ops(ops, per[i], plans[i], ctx.trainings[i], cur, i > 0 or ctx.needs_input_grad[0], mode)
                    grads[i], cur = g, g[0]
                    if big:
                        _run(ops)
                        ops, todo = _Ops(), i
                if ops.n:
                    _run(ops)
                todo = 0
            except _Fallback:
                _pending_meta.clear()
        cur = grads[todo][0] if todo < nb else dout
        for i in reversed(range(todo)):
            g = FusedBasicBlockFunction._backward_calls_of(per[i], plans[i], ctx.trainings[i], cur, i > 0 or ctx.needs_input_grad[0])
            grads[i], cur = g, g[0]
        flat = []
        for g in grads:
            flat += list(g[1:])
        return (grads[0][0], None) + tuple(flat)


def fused_stage(blocks, x):
    """Runs a Sequential of BasicBlocks as one autograd node (FusedStageFunction); None when any of them is not of the fusable
    shape (the caller then runs the Sequential, whose blocks still fuse one by one where they can)."""
    from .tensor import SparseTensor

    if not (_enabled["on"] and _lists["on"]) or not x.F.is_cuda or x.F.dtype != torch.float32 or len(blocks) < 2:
        return None
    cm, key, cin = x.coordinate_manager, x.coordinate_map_key, x.F.shape[1]
    plans, params = [], []
    for block in blocks:
        bp = _block_plan(block, cm, key, cin)
        if bp is None:
            return None
        plan, p, key = bp
        plans.append(plan)
        params += list(p)
        cin = p[3].shape[-1]
    out = FusedStageFunction.apply(x.F, plans, *params)
    return SparseTensor(out, coordinate_map_key=key, coordinate_manager=cm)


class FusedConvNormReLUFunction(torch.autograd.Function):
    """out = relu(bn(conv(x))) — the stem and the strided / transposed transition layers between the stages
    (models/res16unet.py:226-289: `convXpYs2 -> bnX -> relu`, `convtrX -> bntrX -> relu`)."""

    @staticmethod
    def forward(ctx, x, k, g, b, plan):
        x = Fn._rows(x)
        fwd, bwd_getter, flip, norm = plan
        res = None
        mode = Fn.get_precision()
        if _lists["on"] and mode != 0 and Fn._bn_fuse["on"]:
            try:  # the two launches as one list (the 3-channel stem and anything off the tensor-core path fall back)
                ops = _Ops()
                y, m, s, t = _stats(ops, "fwd", _planes_of(x, mode), x.shape[0], fwd, k, flip, norm, mode, x.device)
                out, gc = _op_bn_apply(ops, y, m, s, g, b, None, True, mode)
                _run(ops)
                res = (y, m, s, t, out, gc)
            except _Fallback:
                _pending_meta.clear()
        if res is None:
            y, m, s, t = _conv_norm_stats(x, k, fwd, flip, norm)
            out, gc = Fn.bn_apply_raw(y, g, b, None, m, s, True)
        else:
            y, m, s, t, out, gc = res
        ctx.save_for_backward(x, y, out, m, s, gc, k)
        ctx.plan, ctx.training, ctx.planes = plan, bool(t), _planes(x)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y, out, m, s, gc, k = ctx.saved_tensors
        fwd, bwd_getter, flip, _ = ctx.plan
        _restore(x, ctx.planes)
        dout = Fn._rows(dout)
        w3 = _w3(k, fwd)
        mode = Fn.get_precision()
        if _lists["on"] and mode != 0:
            try:
                ops = _Ops()
                xp = _planes_of(x, mode)
                dy, _, dg, db = _op_bn_backward(ops, dout, y, out, m, s, gc, True, ctx.training, False, mode)
                dx = _op_dgrad(ops, dy, k, w3, bwd_getter, flip, mode) if ctx.needs_input_grad[0] else None
                dk = _op_wgrad(ops, xp, x.shape[0], fwd, dy, w3.shape[1], w3.shape[2], mode, k.shape)
                _run(ops)
                return dx, dk, dg, db, None
            except _Fallback:
                _pending_meta.clear()
        dy, _, dg, db = Fn.bn_backward_raw(dout, y, out, m, s, gc, True, ctx.training, False)
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
