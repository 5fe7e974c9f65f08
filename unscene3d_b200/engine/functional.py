"""torch.autograd.Function wrappers over the C ABI (include/us3d.h).  PyTorch supplies device
memory, streams and the autograd graph; every arithmetic pass is a libus3d kernel.

All functions take row-major fp32 CUDA tensors; a tensor whose last dimension is contiguous is passed
with its row stride as leading dimension (so column slices of a concatenation need no copy).
"""
from __future__ import annotations

import ctypes
import os

import torch

from .._lib import BnFuse, check, lib
from .coords import NeighbourTable, _stream


def _rows(t: torch.Tensor) -> torch.Tensor:
    """2-D fp32 CUDA tensor with unit column stride (copy only if needed)."""
    if not t.is_cuda:
        raise RuntimeError("unscene3d_b200 operators run on CUDA tensors only (no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    if t.ndim != 2 or t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))


def _ptr(t):
    return 0 if t is None else t.data_ptr()


class KernelTimer:
    """Per-launch timing of the convolution kernels (bench.py roofline leg).  While active, libus3d brackets every
    convolution kernel launch with CUDA events recorded on the launch stream from inside the C entry point, a few
    microseconds of host time before the kernel — so the launch gaps of a host-bound section are not counted as kernel
    time, and the kernels are timed where they run, with the cache state the step gives them.
    `summary()` returns (kind, n_in_rows, n_out_rows, kvol, cin, cout, pairs, path, ms) per launch, in launch order
    (pairs = present (input row, output row) pairs of the kernel map, path = "mt" tcgen05 / "stem" / "simt" / "wgrad-tc")."""

    def __init__(self):
        self.meta = []

    def __enter__(self):
        global _timer
        _timer = self
        _dbg().us3d_debug_profile_start()
        return self

    def __exit__(self, *exc):
        global _timer
        _timer = None

    def summary(self):
        import ctypes

        cap = len(self.meta) + 16
        meta = (ctypes.c_int * (7 * cap))()
        ms = (ctypes.c_float * cap)()
        n = _dbg().us3d_debug_profile_stop(meta, ms, cap)
        assert n == len(self.meta), f"profiler saw {n} convolution launches, the host issued {len(self.meta)}"
        return [m + (float(ms[i]),) for i, m in enumerate(self.meta)]


_timer = None
_dbg_lib = None


def _dbg():
    """Debug / profiling hooks of the library (not part of the drop-in surface)."""
    global _dbg_lib
    if _dbg_lib is None:
        import ctypes

        from .._lib import LIB_PATH

        _dbg_lib = ctypes.CDLL(LIB_PATH)
        _dbg_lib.us3d_debug_profile_start.restype = None
        _dbg_lib.us3d_debug_profile_tag.argtypes = [ctypes.c_int]
        _dbg_lib.us3d_debug_profile_tag.restype = None
        _dbg_lib.us3d_debug_profile_stop.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _dbg_lib.us3d_debug_profile_stop.restype = ctypes.c_int
    return _dbg_lib


def _timed(kind, n_in, n_out, kvol, cin, cout, launch, table=None, path="simt"):
    if _timer is not None:
        # (kind, n_in, n_out, kvol, cin, cout, kernel-map pairs, kernel path)
        _timer.meta.append((kind, n_in, n_out, kvol, cin, cout, table.pairs() if table is not None else n_out * kvol, path))
    return launch()


# Arithmetic of the sparse-conv forward / input-gradient GEMMs:
#   0  exact fp32 FMA (SIMT kernel)          1  bf16 x bf16 -> fp32 on tcgen05 (one pass)
#   3  three-term bf16 split on tcgen05 (hi*hi + lo*hi + hi*lo): fp32-faithful products, fp32 accumulation
_precision = {"mode": 3}


def set_precision(mode: int):
    assert mode in (0, 1, 3)
    _precision["mode"] = mode


def get_precision() -> int:
    return _precision["mode"]


def pack_weights(w3: torch.Tensor, transpose: bool, flip_k: bool, passes: int) -> torch.Tensor:
    """fp32 [kvol, cin, cout] -> bf16 slabs in the swizzled shared-memory image the tcgen05 kernel streams."""
    kvol, w_cin, w_cout = w3.shape
    kdim, ndim = (w_cout, w_cin) if transpose else (w_cin, w_cout)
    nbytes = lib.us3d_spconv_packed_bytes(kvol, kdim, ndim, passes)
    out = torch.empty(nbytes, dtype=torch.uint8, device=w3.device)
    check(lib.us3d_spconv_pack_weights(w3.data_ptr(), kvol, w_cin, w_cout, int(transpose), int(flip_k), passes,
                                       out.data_ptr(), _stream()))
    return out


_pack_epoch = {"n": 0}
_packed_bytes_cache = {}


def invalidate_packed_weights():
    """Forget every cached weight image.  The images are keyed on the parameter's version counter, which an optimizer
    step bumps; a loop that changes weights behind autograd's back (or a benchmark that wants to pay the packing of a
    real training step on every iteration) calls this once per step."""
    _pack_epoch["n"] += 1


def _packed_bytes(kvol, kdim, ndim, passes):
    key = (kvol, kdim, ndim, passes)
    v = _packed_bytes_cache.get(key)
    if v is None:
        v = _packed_bytes_cache[key] = int(lib.us3d_spconv_packed_bytes(kvol, kdim, ndim, passes))
    return v


def packed_weights(kernel: torch.Tensor, w3: torch.Tensor, flip_dgrad: bool, passes: int, need_dgrad: bool):
    """(forward image or None, input-gradient image(s) or None) of one convolution's weights, built by ONE call and kept on
    the parameter while its version counter is unchanged.  The input-gradient side is a list of (first output column,
    columns, image): a gradient wider than 256 channels (the 384-channel concatenations of the decoder) is produced in
    column slices of the transposed weight."""
    key = (kernel._version, _pack_epoch["n"], passes, bool(flip_dgrad), w3.data_ptr())
    cached = getattr(kernel, "_us3d_packs", None)
    if cached is not None and cached[0] == key and (cached[2] is not None or not need_dgrad):
        return cached[1], cached[2]
    kvol, cin, cout = w3.shape
    dev = w3.device
    st = _stream()
    fwd = torch.empty(_packed_bytes(kvol, cin, cout, passes), dtype=torch.uint8, device=dev) if _tc_ok(cin, cout) else None
    bwd = None
    if need_dgrad and _tc_ok(cout, cin):
        bwd_img = torch.empty(_packed_bytes(kvol, cout, cin, passes), dtype=torch.uint8, device=dev)
        bwd = [(0, cin, bwd_img)]
        check(lib.us3d_spconv_pack_pair(w3.data_ptr(), kvol, cin, cout, int(flip_dgrad), passes, _ptr(fwd), bwd_img.data_ptr(), st))
    else:
        if fwd is not None:
            check(lib.us3d_spconv_pack_pair(w3.data_ptr(), kvol, cin, cout, int(flip_dgrad), passes, fwd.data_ptr(), 0, st))
        if need_dgrad and cin > 256 and cin % 32 == 0 and _tc_ok(cout, cin // 2):
            half = cin // 2
            bwd = []
            for c0 in (0, half):
                ws = w3[:, c0:c0 + half, :].contiguous()
                img = torch.empty(_packed_bytes(kvol, cout, half, passes), dtype=torch.uint8, device=dev)
                check(lib.us3d_spconv_pack_pair(ws.data_ptr(), kvol, half, cout, int(flip_dgrad), passes, 0, img.data_ptr(), st))
                bwd.append((c0, half, img))
    try:
        kernel._us3d_packs = (key, fwd, bwd)
    except Exception:  # pragma: no cover
        pass
    return fwd, bwd


_PACK_ITEM = [("w", "<u8"), ("out", "<u8"), ("kvol", "<i4"), ("w_cin", "<i4"), ("w_cout", "<i4"), ("w_ci0", "<i4"),
              ("kdim", "<i4"), ("ndim", "<i4"), ("transpose", "<i4"), ("flip", "<i4")]  # == fused::PackItem (csrc/fused_ops.cu)
_pack_plans = {}


def pack_network(net: torch.nn.Module):
    """Re-pack the weight images of EVERY tensor-core convolution under `net` with one kernel launch and mark them current —
    what a training loop calls once per step after the optimizer update (instead of invalidate_packed_weights(), after which
    each of the 62 convolutions of Res16UNet34C packs its own pair on first use).  Images live in buffers owned by the plan
    and are overwritten in stream order.  Convolutions the plan does not cover (exact-fp32 mode, shapes the tensor-core
    kernels do not take) keep packing lazily."""
    import numpy as np

    mode = _precision["mode"]
    _pack_epoch["n"] += 1
    if mode == 0:
        return
    convs = [m for m in net.modules() if hasattr(m, "kernel_volume") and hasattr(m, "IS_TRANSPOSE") and isinstance(getattr(m, "kernel", None), torch.nn.Parameter)]
    convs = [m for m in convs if m.kernel.is_cuda and m.kernel.is_contiguous() and m.kernel.dtype == torch.float32 and m.kernel.shape[-2] > 4]
    if not convs:
        return
    sig = (mode, tuple((m.kernel.data_ptr(), tuple(m.kernel.shape), m.kernel_volume) for m in convs))
    plan = _pack_plans.get(id(net))
    if plan is None or plan["sig"] != sig:
        dev = convs[0].kernel.device
        rows, entries = [], []
        for m in convs:
            kvol, cin, cout = m.kernel_volume, m.kernel.shape[-2], m.kernel.shape[-1]
            flip = (not m.IS_TRANSPOSE) and kvol > 1 and all(s == 1 for s in m.stride) and all(k % 2 == 1 for k in m.kernel_size)
            wptr = m.kernel.data_ptr()
            fwd = None
            if _tc_ok(cin, cout):
                fwd = torch.empty(_packed_bytes(kvol, cin, cout, mode), dtype=torch.uint8, device=dev)
                rows.append((wptr, fwd.data_ptr(), kvol, cin, cout, 0, cin, cout, 0, 0))
            bwd = None
            slices = [(0, cin)] if _tc_ok(cout, cin) else ([(0, cin // 2), (cin // 2, cin // 2)] if cin > 256 and cin % 32 == 0 and _tc_ok(cout, cin // 2) else [])
            if slices:
                bwd = []
                for c0, nc in slices:
                    img = torch.empty(_packed_bytes(kvol, cout, nc, mode), dtype=torch.uint8, device=dev)
                    rows.append((wptr, img.data_ptr(), kvol, cin, cout, c0, cout, nc, 1, int(flip)))
                    bwd.append((c0, nc, img))
            if fwd is not None or bwd is not None:
                entries.append((m, bool(flip), fwd, bwd))
        table = torch.from_numpy(np.array(rows, dtype=_PACK_ITEM).view(np.uint8).copy()).to(dev)
        plan = _pack_plans[id(net)] = {"sig": sig, "table": table, "n": len(rows), "entries": entries, "mode": mode}
    check(lib.us3d_spconv_pack_many(plan["table"].data_ptr(), plan["n"], mode, _stream()))
    epoch = _pack_epoch["n"]
    for m, flip, fwd, bwd in plan["entries"]:
        k = m.kernel
        k._us3d_packs = ((k._version, epoch, mode, flip, k.data_ptr()), fwd, bwd)


def bf16_planes(x: torch.Tensor, need_lo: bool):
    """bf16 hi plane (+ lo = x - hi) of a row-major fp32 tensor, cached on the tensor object while its
    version counter is unchanged (an activation feeds the forward conv, a shortcut conv and the weight gradient)."""
    cached = getattr(x, "_us3d_planes", None)
    if cached is not None and cached[2] == x._version and (cached[1] is not None or not need_lo):
        return cached[0], cached[1]
    n, c = x.shape
    hi = torch.empty((n, c), dtype=torch.bfloat16, device=x.device)
    lo = torch.empty((n, c), dtype=torch.bfloat16, device=x.device) if need_lo else None
    check(lib.us3d_split_bf16(x.data_ptr(), _ld(x), n, c, hi.data_ptr(), _ptr(lo), _stream()))
    try:
        x._us3d_planes = (hi, lo, x._version)
    except Exception:  # pragma: no cover - tensors normally accept attributes
        pass
    return hi, lo


_conv_ws = {}
_SMALL_MAP_ROWS = 148 * 2 * 128  # == the bound inside us3d_spconv_gather_mt_workspace_bytes, without the FFI round trip


def _conv_workspace(dev, n_rows, kvol, cout):
    """Per-device scratch for the split mode of small maps (partial tiles [parts][n_rows][cout] fp32).  Kernels of one stream are
    ordered, so one buffer per device serves the single-stream execution of the module surface; it only grows."""
    if kvol <= 1 or n_rows >= _SMALL_MAP_ROWS:
        return None, 0
    need = kvol * n_rows * cout * 4
    ws = _conv_ws.get(dev.index)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 32 << 20), dtype=torch.uint8, device=dev)
        _conv_ws[dev.index] = ws
    return ws, ws.numel()


def _tc_ok(cin, cout):  # == us3d_spconv_tc_supported, without the FFI round trip
    return cin >= 16 and cin % 16 == 0 and cout >= 16 and cout % 16 == 0 and cout <= 256


class BnRequest:
    """BatchNorm statistics asked of the convolution that produces the normalised tensor: the tensor-core kernel folds the column
    sums into its epilogue (us3d_spconv_gather_mt_bn); any other path computes them with the statistics kernel afterwards.
    After spconv_gather: .mean / .invstd are fp32 [c]; the running statistics are updated like nn.BatchNorm1d's."""

    __slots__ = ("running_mean", "running_var", "momentum", "eps", "num_batches_tracked", "mean", "invstd")

    def __init__(self, running_mean, running_var, momentum, eps, num_batches_tracked):
        self.running_mean, self.running_var, self.momentum, self.eps = running_mean, running_var, momentum, eps
        self.num_batches_tracked = num_batches_tracked
        self.mean = self.invstd = None


_bn_fuse = {"on": os.environ.get("US3D_FUSED_BN_STATS", "1") == "1"}


def set_fused_bn_stats(on: bool):
    """BatchNorm statistics from the convolution's epilogue (default) or from the separate statistics kernel (same sums)."""
    _bn_fuse["on"] = bool(on)


def spconv_gather(x, table: NeighbourTable, w3, cin, cout, transpose_w, flip_k, bias=None, out=None, accumulate=False, wpack=None,
                  bn: "BnRequest" = None):
    """y[j] = sum_k x[nbr[k, j]] . W'[k]  (W' = W[k], or W[K-1-k]^T / W[k]^T for the input gradient).  `wpack` is the
    cached weight image of packed_weights(); without it the image is built here.  `bn`: see BnRequest."""
    y = _spconv_gather(x, table, w3, cin, cout, transpose_w, flip_k, bias, out, accumulate, wpack, bn)
    if bn is not None and bn.mean is None:
        bn.mean, bn.invstd = bn_batch_stats(y, bn.running_mean, bn.running_var, bn.momentum, bn.eps, bn.num_batches_tracked)
    return y


def _spconv_gather(x, table, w3, cin, cout, transpose_w, flip_k, bias, out, accumulate, wpack, bn):
    x = _rows(x)
    y = out if out is not None else torch.empty((table.n_rows, cout), dtype=torch.float32, device=x.device)
    st = _stream()
    mode = _precision["mode"]
    kind = "dgrad" if transpose_w else "fwd"
    if not transpose_w and cin <= 4 and lib.us3d_stem_conv_supported(cin, cout, table.kvol) and not accumulate:
        # stem convolution: exact fp32, warp = rows, lane = output channel
        w3c = w3 if w3.is_contiguous() else w3.contiguous()
        _timed(kind, x.shape[0], table.n_rows, table.kvol, cin, cout, lambda: check(
            lib.us3d_stem_conv_fwd(x.data_ptr(), _ld(x), table.nbr.data_ptr(), table.n_rows, table.kvol, w3c.data_ptr(), cin, cout,
                                   _ptr(bias), y.data_ptr(), _ld(y), st)), table, "stem")
        return y
    if (mode != 0 and _tc_ok(cin, cout) and x.data_ptr() % 16 == 0 and _ld(x) % 4 == 0):
        if wpack is None:
            wpack = pack_weights(w3, transpose_w, flip_k, mode)
        hi, lo = bf16_planes(x, mode == 3)
        nbr, mask, order = table.ordered() or (table.nbr, table.mask, None)
        part = table.partition() if order is not None else None
        ws, ws_bytes = _conv_workspace(x.device, table.n_rows, table.kvol, cout) if order is None else (None, 0)
        desc = None
        if bn is not None and _bn_fuse["on"] and not accumulate and table.n_rows > 0:
            stats = torch.empty((2, cout), dtype=torch.float32, device=x.device)
            desc = BnFuse(_bn_workspace(x.device, cout).data_ptr(), stats.data_ptr(), stats.data_ptr() + 4 * cout, _ptr(bn.running_mean),
                          _ptr(bn.running_var), _ptr(bn.num_batches_tracked), float(bn.eps),
                          float(bn.momentum if bn.momentum is not None else 0.0))
            bn.mean, bn.invstd = stats[0], stats[1]
        _timed(kind, x.shape[0], table.n_rows, table.kvol, cin, cout, lambda: check(
            lib.us3d_spconv_gather_mt_bn(hi.data_ptr(), _ptr(lo), x.shape[0], nbr.data_ptr(), table.n_rows, table.kvol, wpack.data_ptr(),
                                         cin, cout, mode, _ptr(bias), _ptr(order), y.data_ptr(), _ld(y), int(accumulate), _ptr(mask),
                                         _ptr(part), _ptr(ws), ws_bytes, ctypes.byref(desc) if desc is not None else None, st)),
               table, "mt")
        return y
    _timed(kind, x.shape[0], table.n_rows, table.kvol, cin, cout, lambda: check(
        lib.us3d_spconv_gather(x.data_ptr(), _ld(x), table.nbr.data_ptr(), table.n_rows, table.kvol, w3.data_ptr(), cin, cout,
                               int(transpose_w), int(flip_k), _ptr(bias), 0, y.data_ptr(), _ld(y), int(accumulate),
                               _ptr(table.mask), st)), table, "simt")
    return y


class _ZeroArena:
    """Zero-filled fp32 storage for the weight gradients of a backward pass: one allocation + one fill per ~64 MB instead of one
    torch.zeros (allocation + fill launch) per convolution.  Slices are handed out once and never reused — the gradients
    outlive the pass (.grad, optimizer) and keep their arena alive through the view."""

    def __init__(self):
        self.buf, self.off = {}, {}

    def take(self, n: int, dev: torch.device) -> torch.Tensor:
        buf, off = self.buf.get(dev.index), self.off.get(dev.index, 0)
        need = (n + 63) & ~63  # 256-byte granules
        if buf is None or off + need > buf.numel():
            buf = torch.zeros(max(need, 16 << 20), dtype=torch.float32, device=dev)
            off = 0
        self.buf[dev.index], self.off[dev.index] = buf, off + need
        return buf[off:off + n]


_zero_arena = _ZeroArena()


def spconv_wgrad(x, table: NeighbourTable, dy, cin, cout):
    x, dy = _rows(x), _rows(dy)
    dw = _zero_arena.take(table.kvol * cin * cout, x.device).view(table.kvol, cin, cout)
    st = _stream()
    mode = _precision["mode"]
    if cin == 3 and table.kvol in (1, 8, 27):
        _timed("wgrad", x.shape[0], table.n_rows, table.kvol, cin, cout, lambda: check(
            lib.us3d_stem_conv_wgrad(x.data_ptr(), _ld(x), table.nbr.data_ptr(), table.n_rows, table.kvol, dy.data_ptr(), _ld(dy),
                                     dw.data_ptr(), cin, cout, st)), table, "stem")
        return dw
    if (mode != 0 and _wgrad_tc_ok(cin, cout) and x.data_ptr() % 16 == 0 and dy.data_ptr() % 16 == 0
            and _ld(x) % 4 == 0 and _ld(dy) % 4 == 0):
        xh, xl = bf16_planes(x, mode == 3)
        dh, dl = bf16_planes(dy, mode == 3)
        # pattern order pays for the weight gradient only on k2s2 maps (one parent per fine row: 1/8 of the tile x offset
        # products remain); on k3 maps the scattered dY rows cost more than the skipped products save (measured on B200:
        # 200k voxels 96 -> 96, 0.55 -> 0.62 ms ordered, k2s2 0.171 -> 0.125 ms)
        wgrad_order = _wgrad_order["mode"]
        nbr, mask, order = (table.ordered() if (table.kvol <= 8 or wgrad_order == "permute") else None) or (table.nbr, table.mask, None)
        if order is not None and table.kvol > 8:
            # k3 maps: bring the dY planes into table order with one streaming pass (out[j] = dY[order[j]]) instead of letting
            # the kernel gather scattered dY rows
            n, c = dy.shape
            ph = torch.empty_like(dh)
            pl = torch.empty_like(dl) if dl is not None else None
            check(lib.us3d_permute_planes(dh.data_ptr(), _ptr(dl), order.data_ptr(), n, c, ph.data_ptr(), _ptr(pl), st))
            dh, dl, order = ph, pl, None
        _timed("wgrad", x.shape[0], table.n_rows, table.kvol, cin, cout, lambda: check(
            lib.us3d_spconv_wgrad_planes(xh.data_ptr(), _ptr(xl), dh.data_ptr(), _ptr(dl), nbr.data_ptr(), table.n_rows,
                                         table.kvol, dw.data_ptr(), cin, cout, mode, _ptr(mask), _ptr(order), st)), table, "wgrad-tc")
        return dw
    _timed("wgrad", x.shape[0], table.n_rows, table.kvol, cin, cout, lambda: check(
        lib.us3d_spconv_wgrad(x.data_ptr(), _ld(x), table.nbr.data_ptr(), table.n_rows, table.kvol, dy.data_ptr(), _ld(dy), 0,
                              dw.data_ptr(), cin, cout, st)), table, "simt")
    return dw


# weight gradient on pattern-ordered k3 tables: "natural" = keep the natural table (no pruning), "permute" = ordered table with
# the dY planes permuted into table order first
_wgrad_order = {"mode": os.environ.get("US3D_WGRAD_ORDER", "natural")}


def _wgrad_tc_ok(cin, cout):  # == us3d_spconv_wgrad_tc_supported, without the FFI round trip
    return cin >= 8 and cin % 8 == 0 and cout >= 16 and cout % 16 == 0 and cout <= 256


def conv_input_gradient(dy, kernel, w3, bwd_getter, flip_dgrad):
    """dX of a convolution: the same gather kernel on the transposed table with W^T (column slices for gradients wider than
    256 channels, written in place)."""
    cin, cout = w3.shape[1], w3.shape[2]
    bwd, flip = bwd_getter()
    mode = _precision["mode"]
    chunks = None
    if mode != 0 and flip == flip_dgrad and cin > 4:
        chunks = packed_weights(kernel, w3, flip, mode, True)[1]
    if chunks is None or len(chunks) == 1:
        return spconv_gather(dy, bwd, w3, cout, cin, True, flip, wpack=None if chunks is None else chunks[0][2])
    dx = torch.empty((bwd.n_rows, cin), dtype=torch.float32, device=dy.device)
    for c0, nc, img in chunks:
        spconv_gather(dy, bwd, w3[:, c0:c0 + nc, :], cout, nc, True, flip, out=dx[:, c0:c0 + nc], wpack=img)
    return dx


class SparseConvFunction(torch.autograd.Function):
    """Y = conv(X) over `fwd` ([K, n_out] neighbour table); backward uses `bwd` ([K, n_in], flip flag).
    `flip_dgrad` is the flag the backward table will carry (known from the map pair at forward time), so that the
    forward and input-gradient weight images are packed by one call."""

    @staticmethod
    def forward(ctx, x, kernel, bias, fwd: NeighbourTable, bwd_getter, flip_dgrad=False):
        x = _rows(x)
        w3 = kernel.detach().contiguous().view(fwd.kvol, kernel.shape[-2], kernel.shape[-1])
        cin, cout = w3.shape[1], w3.shape[2]
        assert x.shape[1] == cin, f"input has {x.shape[1]} channels, kernel expects {cin}"
        mode = _precision["mode"]
        wf = None
        if mode != 0 and cin > 4:
            wf, _ = packed_weights(kernel, w3, flip_dgrad, mode, bool(ctx.needs_input_grad[0]))
        y = spconv_gather(x, fwd, w3, cin, cout, False, False, None if bias is None else bias.detach().contiguous(), wpack=wf)
        ctx.save_for_backward(x, kernel)
        ctx.fwd, ctx.bwd_getter, ctx.has_bias, ctx.flip_dgrad = fwd, bwd_getter, bias is not None, flip_dgrad
        ctx.x_planes = getattr(x, "_us3d_planes", None)  # the weight gradient re-uses the forward's bf16 planes
        return y

    @staticmethod
    def backward(ctx, dy):
        x, kernel = ctx.saved_tensors
        dy = _rows(dy)
        fwd = ctx.fwd
        if ctx.x_planes is not None and ctx.x_planes[2] == x._version:
            x._us3d_planes = ctx.x_planes
        w3 = kernel.detach().contiguous().view(fwd.kvol, kernel.shape[-2], kernel.shape[-1])
        cin, cout = w3.shape[1], w3.shape[2]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = conv_input_gradient(dy, kernel, w3, ctx.bwd_getter, ctx.flip_dgrad)
        if ctx.needs_input_grad[1]:
            dw = spconv_wgrad(x, fwd, dy, cin, cout).view(kernel.shape)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(0, keepdim=True)
        return dx, dw, db, None, None, None


_bn_ws = {}


def _bn_workspace(dev, c):
    """Per-device scratch of the BatchNorm column reductions: 2*c doubles + a ticket counter, ZERO between uses (every
    kernel that accumulates into it is followed by one whose last block zeroes it again).  Kernels on one stream are
    ordered, so one buffer per device serves the single-stream execution the module surface uses."""
    ws = _bn_ws.get(dev.index)
    if ws is None or ws.numel() < 2 * c + 2:
        ws = torch.zeros(max(2 * c + 2, 2050), dtype=torch.float64, device=dev)
        _bn_ws[dev.index] = ws
    return ws


def bn_batch_stats(x, running_mean, running_var, momentum, eps, num_batches_tracked=None):
    """Batch statistics of all rows in ONE launch (no autograd: the apply function's backward carries the dependence on
    them).  Returns (mean, invstd) fp32 [c]; updates the running statistics and num_batches_tracked like nn.BatchNorm1d."""
    x = _rows(x)
    n, c = x.shape
    dev = x.device
    stats = torch.empty((2, c), dtype=torch.float32, device=dev)
    check(lib.us3d_bn_stats_fused(x.data_ptr(), _ld(x), n, c, float(eps), float(momentum if momentum is not None else 0.0),
                                  stats.data_ptr(), stats.data_ptr() + 4 * c, _ptr(running_mean), _ptr(running_var),
                                  _ptr(num_batches_tracked), _bn_workspace(dev, c).data_ptr(), _stream()))
    return stats[0], stats[1]


def _want_planes(c):
    """bf16 planes are written by the producing pass when a tensor-core convolution can consume them."""
    return _precision["mode"] != 0 and c % 16 == 0


def _detached(p: torch.Tensor) -> torch.Tensor:
    """p.detach().contiguous(), kept on the parameter while its storage is unchanged (shares the storage: values stay current)."""
    c = getattr(p, "_us3d_det", None)
    ptr = p.data_ptr()
    if c is not None and c[0] == ptr:
        return c[1]
    d = p.detach().contiguous()
    if d.data_ptr() == ptr:
        try:
            p._us3d_det = (ptr, d)
        except Exception:  # pragma: no cover
            pass
    return d


def bn_apply_raw(x, gamma, beta, residual, mean, invstd, relu):
    """One pass: y = [relu](bn(x) [+ residual]) as fp32 rows + the bf16 planes of y (cached on y for the next convolution).
    Returns (y, gamma as passed to the kernel)."""
    n, c = x.shape
    dev = x.device
    if residual is not None:
        residual = _rows(residual)
    g = _detached(gamma) if gamma is not None else torch.ones(c, device=dev)
    b = _detached(beta) if beta is not None else torch.zeros(c, device=dev)
    y = torch.empty((n, c), dtype=torch.float32, device=dev)
    hi = lo = None
    if _want_planes(c) and n > 0:
        hi = torch.empty((n, c), dtype=torch.bfloat16, device=dev)
        lo = torch.empty((n, c), dtype=torch.bfloat16, device=dev) if _precision["mode"] == 3 else None
    check(lib.us3d_bn_apply_planes(x.data_ptr(), _ld(x), n, c, mean.data_ptr(), invstd.data_ptr(), g.data_ptr(), b.data_ptr(),
                                   _ptr(residual), 0 if residual is None else _ld(residual), int(relu), y.data_ptr(), c,
                                   _ptr(hi), _ptr(lo), _stream()))
    if hi is not None:
        y._us3d_planes = (hi, lo, y._version)
    return y, g


def bn_backward_raw(dy, x, y, mean, invstd, g, relu, training, has_res):
    """BatchNorm (+ReLU mask, + residual branch) backward in one call: (dx with its bf16 planes cached, dresidual or None,
    dgamma, dbeta).  `y` is the forward output (needed for the ReLU mask only; may be None without ReLU)."""
    n, c = x.shape
    dev = x.device
    yy = y if y is not None else x
    dx = torch.empty((n, c), dtype=torch.float32, device=dev)
    dres = torch.empty((n, c), dtype=torch.float32, device=dev) if has_res else None
    dgb = torch.empty((2, c), dtype=torch.float32, device=dev)
    hi = lo = None
    if _want_planes(c) and dy.data_ptr() % 16 == 0 and _ld(dy) % 4 == 0 and x.data_ptr() % 16 == 0 and _ld(x) % 4 == 0:
        hi = torch.empty((n, c), dtype=torch.bfloat16, device=dev)
        lo = torch.empty((n, c), dtype=torch.bfloat16, device=dev) if _precision["mode"] == 3 else None
    check(lib.us3d_bn_backward_planes(dy.data_ptr(), _ld(dy), x.data_ptr(), _ld(x), yy.data_ptr(), _ld(yy), n, c, mean.data_ptr(),
                                      invstd.data_ptr(), g.data_ptr(), int(relu), int(training),
                                      _bn_workspace(dev, c).data_ptr(), dx.data_ptr(), c, _ptr(dres), c, dgb.data_ptr(),
                                      dgb.data_ptr() + 4 * c, _ptr(hi), _ptr(lo), _stream()))
    if hi is not None:
        dx._us3d_planes = (hi, lo, dx._version)
    return dx, dres, dgb[0], dgb[1]


class BatchNormApplyFunction(torch.autograd.Function):
    """y = [relu]( (x - mean) * invstd * gamma + beta [+ residual] ) with (mean, invstd) given.  `batch_stats`
    says whether they are this batch's statistics (training: backward includes the two batch terms) or constants.
    The same pass writes y's bf16 planes (cached on y for the convolution that follows); backward does the same for dx."""

    @staticmethod
    def forward(ctx, x, gamma, beta, residual, mean, invstd, batch_stats, relu):
        x = _rows(x)
        y, g = bn_apply_raw(x, gamma, beta, residual, mean, invstd, relu)
        ctx.save_for_backward(x, y if relu else None, mean, invstd, g)
        ctx.relu, ctx.training, ctx.has_res = bool(relu), bool(batch_stats), residual is not None
        ctx.affine = gamma is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, mean, invstd, g = ctx.saved_tensors
        dx, dres, dgamma, dbeta = bn_backward_raw(_rows(dy), x, y, mean, invstd, g, ctx.relu, ctx.training, ctx.has_res)
        if not ctx.affine:
            dgamma = dbeta = None
        return dx, dgamma, dbeta, dres, None, None, None, None


class BatchNormFunction:
    """Statistics + apply in one call (kept for callers that do not need the deferred form)."""

    @staticmethod
    def apply(x, gamma, beta, residual, running_mean, running_var, momentum, eps, training, relu):
        if training:
            mean, invstd = bn_batch_stats(x.detach(), running_mean, running_var, momentum, eps)
        else:
            mean = running_mean.detach().float()
            invstd = torch.rsqrt(running_var.detach().float() + eps)
        return BatchNormApplyFunction.apply(x, gamma, beta, residual, mean, invstd, training, relu)


class ReLUFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, inplace):
        xc = x if (x.is_contiguous() and x.dtype == torch.float32) else x.float().contiguous()
        if not xc.is_cuda:
            raise RuntimeError("unscene3d_b200 operators run on CUDA tensors only (no CPU fallback)")
        y = xc if (inplace and xc is x) else torch.empty_like(xc)
        check(lib.us3d_relu(xc.data_ptr(), y.data_ptr(), xc.numel(), _stream()))
        if y is x:
            ctx.mark_dirty(x)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        check(lib.us3d_relu_bwd(dy.data_ptr(), y.data_ptr(), dx.data_ptr(), dy.numel(), _stream()))
        return dx, None


class AddFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _rows(a).contiguous(), _rows(b).contiguous()
        assert a.shape == b.shape
        z = torch.empty_like(a)
        check(lib.us3d_add(a.data_ptr(), b.data_ptr(), z.data_ptr(), a.numel(), _stream()))
        return z

    @staticmethod
    def backward(ctx, dz):
        return dz, dz


class CatFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *feats):
        feats = [_rows(f) for f in feats]
        n = feats[0].shape[0]
        widths = [f.shape[1] for f in feats]
        out = torch.empty((n, sum(widths)), dtype=torch.float32, device=feats[0].device)
        st, c0 = _stream(), 0
        for f, w in zip(feats, widths):
            assert f.shape[0] == n
            dst = out[:, c0:c0 + w]
            check(lib.us3d_copy2d(f.data_ptr(), _ld(f), dst.data_ptr(), out.shape[1], n, w, st))
            c0 += w
        ctx.widths = widths
        return out

    @staticmethod
    def backward(ctx, dout):
        grads, c0 = [], 0
        for w in ctx.widths:
            grads.append(dout[:, c0:c0 + w])  # consumers accept a leading dimension: no copy
            c0 += w
        return tuple(grads)


class PoolFunction(torch.autograd.Function):
    MODES = {"avg": 0, "sum": 1, "max": 2}

    @staticmethod
    def forward(ctx, x, table: NeighbourTable, n_in, mode):
        x = _rows(x).contiguous()
        c = x.shape[1]
        y = torch.empty((table.n_rows, c), dtype=torch.float32, device=x.device)
        check(lib.us3d_pool_fwd(x.data_ptr(), c, table.nbr.data_ptr(), table.n_rows, table.kvol, mode, y.data_ptr(), _stream()))
        ctx.table, ctx.mode, ctx.n_in = table, mode, n_in
        ctx.save_for_backward(x, y) if mode == 2 else ctx.save_for_backward()
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _rows(dy).contiguous()
        c = dy.shape[1]
        x, y = ctx.saved_tensors if ctx.mode == 2 else (None, None)
        dx = torch.zeros((ctx.n_in, c), dtype=torch.float32, device=dy.device)
        t = ctx.table
        check(lib.us3d_pool_bwd(dy.data_ptr(), _ptr(x), _ptr(y), c, t.nbr.data_ptr(), t.n_rows, t.kvol, ctx.mode, dx.data_ptr(), _stream()))
        return dx, None, None, None


class SegmentMeanFunction(torch.autograd.Function):
    """torch_scatter.scatter_mean(src, index, dim=0) (models/mask3d.py:223)."""

    @staticmethod
    def forward(ctx, src, index, num_segments):
        src = _rows(src).contiguous()
        index = index.to(src.device).contiguous().long()
        n, c = src.shape
        assert index.shape[0] == n, f"scatter_mean: {index.shape[0]} indices for {n} rows"
        out = torch.zeros((num_segments, c), dtype=torch.float32, device=src.device)
        count = torch.zeros(num_segments, dtype=torch.float32, device=src.device)
        check(lib.us3d_segment_mean_fwd(src.data_ptr(), index.data_ptr(), n, c, num_segments, out.data_ptr(), count.data_ptr(), _stream()))
        ctx.save_for_backward(index, count)
        return out

    @staticmethod
    def backward(ctx, dout):
        index, count = ctx.saved_tensors
        dout = _rows(dout).contiguous()
        n, c = index.shape[0], dout.shape[1]
        dsrc = torch.empty((n, c), dtype=torch.float32, device=dout.device)
        check(lib.us3d_segment_mean_bwd(dout.data_ptr(), index.data_ptr(), count.data_ptr(), n, c, dsrc.data_ptr(), _stream()))
        return dsrc, None, None


def furthest_point_sampling(xyz: torch.Tensor, nsamples: int) -> torch.Tensor:
    """pointnet2._ext.furthest_point_sampling(points[B,N,3] float cuda contiguous, nsamples) -> int32 [B, nsamples]."""
    if not xyz.is_cuda:
        raise RuntimeError("CPU not supported")  # same contract as _ext_src/src/sampling.cpp:83-85
    xyz = xyz.float().contiguous()
    B, N, _ = xyz.shape
    idx = torch.zeros((B, nsamples), dtype=torch.int32, device=xyz.device)
    temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
    check(lib.us3d_furthest_point_sampling(xyz.data_ptr(), B, N, nsamples, temp.data_ptr(), idx.data_ptr(), _stream()))
    return idx


def fourier_posenc(xyz, gauss_b, d_out, lo=None, hi=None):
    """[sin | cos](2 pi normalise(xyz) . gauss_B[:, :d_out]) as [n, 2 d_out] rows (models/position_embedding.py:128-172)."""
    if not xyz.is_cuda:
        raise RuntimeError("unscene3d_b200 operators run on CUDA tensors only (no CPU fallback)")
    xyz = _rows(xyz)
    gb = gauss_b if (gauss_b.dtype == torch.float32 and gauss_b.stride(1) == 1) else gauss_b.float().contiguous()
    n = xyz.shape[0]
    out = torch.empty((n, 2 * d_out), dtype=torch.float32, device=xyz.device)
    if lo is not None:
        lo, hi = lo.reshape(-1).float().contiguous(), hi.reshape(-1).float().contiguous()
    check(lib.us3d_fourier_posenc(xyz.data_ptr(), n, _ld(xyz), _ptr(lo), _ptr(hi), gb.data_ptr(), gb.stride(0), int(d_out),
                                  out.data_ptr(), _stream()))
    return out


def matcher_cost(logits_sq, tgt_ts, prob_qc, labels_t, w_class, w_mask, w_dice):
    """Cost matrix [Q, T] of models/matcher.py:97-168 for one scene."""
    logits_sq = _rows(logits_sq).contiguous()
    tgt_ts = _rows(tgt_ts).contiguous()
    prob_qc = _rows(prob_qc).contiguous()
    labels_t = labels_t.to(logits_sq.device).contiguous().long()  # the reference indexes with CPU or CUDA labels alike
    S, Q = logits_sq.shape
    T = tgt_ts.shape[0]
    cost = torch.empty((Q, T), dtype=torch.float32, device=logits_sq.device)
    check(lib.us3d_matcher_cost(logits_sq.data_ptr(), S, Q, tgt_ts.data_ptr(), T, prob_qc.data_ptr(), prob_qc.shape[1],
                                labels_t.data_ptr(), float(w_class), float(w_mask), float(w_dice), cost.data_ptr(), _stream()))
    return cost


# ------------------------------------------------------------------------------------------- decoder cross-attention
class DecoderMask:
    """The decoder's attention mask as it is built, [B, K, Q] bool with True = hidden and shared by all heads
    (models/mask3d.py:338-349), i.e. BEFORE the reference's repeat_interleave(num_heads) + permute to [B*h, Q, K]
    (:358).  The kernels read it in place; `torch_layout` gives what nn.MultiheadAttention would need."""

    def __init__(self, bkq: torch.Tensor):
        self.bkq = bkq

    def torch_layout(self, num_heads: int) -> torch.Tensor:
        return self.bkq.repeat_interleave(num_heads, dim=0).permute((0, 2, 1))


def _mask_view(mask, B, H, Q, K):
    """(pointer-holding tensor, batch/head/query/key strides in elements) of a boolean attention mask given either as the
    decoder's own [B, K, Q] tensor (models/mask3d.py:338-349, before its repeat_interleave + permute) or in
    nn.MultiheadAttention's layouts [Q, K] / [B*H, Q, K] (any strides: a permuted or expanded view is read in place)."""
    if mask is None:
        return None, (0, 0, 0, 0)
    if hasattr(mask, "bkq"):
        m = mask.bkq
        if m.dtype != torch.bool or tuple(m.shape) != (B, K, Q):
            raise RuntimeError(f"masked_cross_attention: decoder mask must be bool [B,K,Q] = {(B, K, Q)}, got {m.dtype} {tuple(m.shape)}")
        m = m.view(torch.uint8)
        return m, (m.stride(0), 0, m.stride(2), m.stride(1))
    if mask.dtype != torch.bool and mask.dtype != torch.uint8:
        raise RuntimeError("masked_cross_attention: only boolean masks (True = hidden) are supported")
    if mask.dtype == torch.bool:
        mask = mask.view(torch.uint8)
    if mask.ndim == 2 and tuple(mask.shape) == (Q, K):
        return mask, (0, 0, mask.stride(0), mask.stride(1))
    if mask.ndim == 3 and tuple(mask.shape) == (B * H, Q, K):
        return mask, (H * mask.stride(0), mask.stride(0), mask.stride(1), mask.stride(2))
    raise RuntimeError(f"masked_cross_attention: mask shape {tuple(mask.shape)} is neither [Q,K] nor [B*H,Q,K]")


class MaskedCrossAttentionFunction(torch.autograd.Function):
    """softmax(scale * q k^T + mask) v per (scene, head) on projected q [Q,B,E], k / v [K,B,E] (csrc/attention.cu)."""

    @staticmethod
    def forward(ctx, q, k, v, mask, num_heads):
        if not q.is_cuda:
            raise RuntimeError("unscene3d_b200 operators run on CUDA tensors only (no CPU fallback)")
        q, k, v = q.float().contiguous(), k.float().contiguous(), v.float().contiguous()
        Q, B, E = q.shape
        K = k.shape[0]
        hd = E // num_heads
        scale = float(hd) ** -0.5
        mview, ms = _mask_view(mask, B, num_heads, Q, K)
        ws = torch.empty(int(lib.us3d_xattn_workspace_bytes(B, num_heads, Q, K, hd)) // 4, dtype=torch.float32, device=q.device)
        out = torch.empty_like(q)
        lse = torch.empty((B, num_heads, Q), dtype=torch.float32, device=q.device)
        check(lib.us3d_xattn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), _ptr(mview), *ms, B, num_heads, Q, K, hd, scale,
                                 ws.data_ptr(), out.data_ptr(), lse.data_ptr(), _stream()))
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.mask, ctx.ms, ctx.dims = mview, ms, (B, num_heads, Q, K, hd, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        B, H, Q, K, hd, scale = ctx.dims
        dout = dout.float().contiguous()
        ws = torch.empty(int(lib.us3d_xattn_workspace_bytes(B, H, Q, K, hd)) // 4, dtype=torch.float32, device=q.device)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        check(lib.us3d_xattn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), _ptr(ctx.mask), *ctx.ms, out.data_ptr(), lse.data_ptr(),
                                 dout.data_ptr(), B, H, Q, K, hd, scale, ws.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                 _stream()))
        return dq, dk, dv, None, None


def multihead_cross_attention(mha: torch.nn.MultiheadAttention, query, key, value, attn_mask=None):
    """`mha(query, key, value, attn_mask=attn_mask)[0]` for a sequence-first nn.MultiheadAttention with a boolean mask —
    same parameters (`in_proj_weight`, `in_proj_bias`, `out_proj`), so state dicts are interchangeable; the three input
    projections and the output projection are library GEMMs, the attention core is the libus3d kernel."""
    if mha.batch_first or not mha._qkv_same_embed_dim or mha.bias_k is not None or mha.add_zero_attn:
        raise RuntimeError("multihead_cross_attention: unsupported nn.MultiheadAttention configuration")
    if mha.dropout > 0.0 and mha.training:
        raise RuntimeError("multihead_cross_attention: attention dropout is not implemented (every shipped config uses 0.0)")
    E = mha.embed_dim
    w, b = mha.in_proj_weight, mha.in_proj_bias
    bq, bk, bv = (None, None, None) if b is None else (b[:E], b[E:2 * E], b[2 * E:])
    q = torch.nn.functional.linear(query, w[:E], bq)
    k = torch.nn.functional.linear(key, w[E:2 * E], bk)
    v = torch.nn.functional.linear(value, w[2 * E:], bv)
    ctx = MaskedCrossAttentionFunction.apply(q, k, v, attn_mask, mha.num_heads)
    return mha.out_proj(ctx)


# ------------------------------------------------------------------------------------------- set-criterion mask losses
class MaskLossFunction(torch.autograd.Function):
    """(loss_mask, loss_dice) of the matched pairs of one scene — models/criterion.py:22-73 applied as in loss_masks
    (:168-216): logits [S, Q] (pred_masks of the scene), targets [T_all, S] (bool / uint8 / float), matched query ids and
    target ids [T], per-pair weights or None, the scene's normaliser n.  One pass over each matched column forward, one
    backward (csrc/decoder_ops.cu)."""

    @staticmethod
    def forward(ctx, logits, targets, qidx, tidx, weights, n):
        if not logits.is_cuda:
            raise RuntimeError("unscene3d_b200 operators run on CUDA tensors only (no CPU fallback)")
        logits = logits.float().contiguous()
        S, Q = logits.shape
        if targets.dtype == torch.bool:
            targets = targets.view(torch.uint8)
        elif targets.dtype not in (torch.uint8, torch.float32):
            targets = targets.float()
        targets = targets.contiguous()
        assert targets.shape[1] == S, f"targets have {targets.shape[1]} columns, logits {S} rows"
        qidx, tidx = qidx.to(logits.device).long().contiguous(), tidx.to(logits.device).long().contiguous()
        T = qidx.shape[0]
        w = None if weights is None else weights.float().contiguous()
        stats = torch.empty((max(T, 1), 4), dtype=torch.float32, device=logits.device)
        out = torch.empty(2, dtype=torch.float32, device=logits.device)
        is_float = int(targets.dtype == torch.float32)
        check(lib.us3d_mask_loss_fwd(logits.data_ptr(), S, Q, targets.data_ptr(), is_float, qidx.data_ptr(), tidx.data_ptr(), T, _ptr(w),
                                     float(n), stats.data_ptr(), out.data_ptr(), _stream()))
        ctx.save_for_backward(logits, targets, qidx, tidx, stats)
        ctx.w, ctx.n, ctx.is_float = w, float(n), is_float
        return out

    @staticmethod
    def backward(ctx, gout):
        logits, targets, qidx, tidx, stats = ctx.saved_tensors
        S, Q = logits.shape
        gout = gout.float().contiguous()
        dlogits = torch.empty_like(logits)
        check(lib.us3d_mask_loss_bwd(logits.data_ptr(), S, Q, targets.data_ptr(), ctx.is_float, qidx.data_ptr(), tidx.data_ptr(),
                                     qidx.shape[0], _ptr(ctx.w), ctx.n, stats.data_ptr(), gout.data_ptr(), dlogits.data_ptr(), _stream()))
        return dlogits, None, None, None, None, None


def mask_losses(logits_sq, targets_ts, qidx, tidx, weights, n):
    """-> (loss_mask, loss_dice) scalars of one scene."""
    out = MaskLossFunction.apply(logits_sq, targets_ts, qidx, tidx, weights, n)
    return out[0], out[1]


# ------------------------------------------------------------------------------------------- decoder attention masks
def _count(index, n):
    """Occurrences of 0..n-1 in `index` — torch.bincount(index, minlength=n) without its host synchronisation (bincount reads the
    maximum back to size its result even when minlength is given)."""
    return torch.zeros(n, dtype=torch.int64, device=index.device).scatter_add_(0, index, torch.ones_like(index))


def _segment_pool_matrix(cm, key0, point2segment, sizes, steps):
    """CSR matrix A_L [N_L, S_total] with  avgpool^L(seglogit[point2segment])[v, :] = sum_s A_L[v, s] seglogit[s, :]  for the
    coordinate manager's pyramid below `key0`: A_0[p, seg(p)] = 1, A_{l+1}[v, :] = mean over the present children u of A_l[u, :]
    (MinkowskiAvgPooling(k2, s2) = mean over present inputs).  Built level by level from the parent maps, once per
    coordinate manager (i.e. per step) and level; returns (key_L, rowptr, col, val, S_total)."""
    cache = cm.__dict__.setdefault("_segment_pool", {})
    sig = (key0, tuple(int(p.data_ptr()) for p in point2segment))
    entry = cache.get(sig)
    if entry is None:
        dev = point2segment[0].device
        offs, tot = [], 0
        for s in sizes:
            offs.append(tot)
            tot += s
        cols = torch.cat([p.long() + o for p, o in zip(point2segment, offs)])
        n0 = cm.size(key0)
        assert cols.shape[0] == n0, "point2segment does not cover the coordinate map"
        entry = cache[sig] = {"S": tot, "levels": [(key0, torch.arange(n0, device=dev), cols, torch.ones(n0, dtype=torch.float32, device=dev))]}
    levels = entry["levels"]
    while len(levels) <= steps:
        key, rows, cols, vals = levels[-1]
        nxt = cm.stride(key, (2, 2, 2))
        parent = cm._parents[(key, nxt)].long()
        n_next = cm.size(nxt)
        count = _count(parent, n_next).float()
        prow = parent[rows]
        merged, inv = torch.unique(prow * entry["S"] + cols, return_inverse=True)
        v2 = torch.zeros(merged.shape[0], dtype=torch.float32, device=vals.device).index_add_(0, inv, vals / count[prow])
        levels.append((nxt, merged // entry["S"], merged % entry["S"], v2))
    key, rows, cols, vals = levels[steps]
    csr = entry.setdefault("csr", {})
    if steps not in csr:
        n = cm.size(key)
        rowptr = torch.zeros(n + 1, dtype=torch.int64, device=rows.device)
        rowptr[1:] = torch.cumsum(_count(rows, n), 0)   # rows are sorted (level 0: arange; above: unique keys)
        csr[steps] = (rowptr, cols.contiguous(), vals.contiguous())
    return (key, *csr[steps], entry["S"])


def prepare_segment_attention(x, point2segment, max_steps):
    """Builds the pooling matrices of every decoder level ahead of the backbone, on the coordinate stream when one is set
    (engine.set_coordinate_stream): they depend on coordinates and segment ids only, and de-duplicating their entries hands
    sizes back to the host — on the compute stream, in the middle of the decoder, each of those synchronisations would wait
    for the queued backbone kernels and cost the host its run-ahead."""
    from .coords import get_coordinate_stream

    cm, key0 = x.coordinate_manager, x.coordinate_map_key
    side = get_coordinate_stream(point2segment[0].device)
    if side is None:
        sizes = [int(p.max()) + 1 if p.numel() else 0 for p in point2segment]
        _segment_pool_matrix(cm, key0, point2segment, sizes, max_steps)
        return
    main = torch.cuda.current_stream(point2segment[0].device)
    with torch.cuda.stream(side):
        sizes = [int(p.max()) + 1 if p.numel() else 0 for p in point2segment]
        for steps in range(1, max_steps + 1):
            for t in _segment_pool_matrix(cm, key0, point2segment, sizes, steps)[1:4]:
                t.record_stream(main)
    main.wait_stream(side)


def segment_attention_masks(mask_features, seg_logits, point2segment, num_pooling_steps):
    """Boolean attention mask of one decoder round at the level `num_pooling_steps` strides below `mask_features`
    (Mask3D.mask_module, models/mask3d.py:419-446): features [N_L, Q] bool = sigmoid(avgpool^L(seg_logits[point2segment])) < 0.5,
    as a (key, tensor) pair on mask_features' coordinate manager.  One sparse product instead of the point-level gather,
    the concatenation and L pooling passes over [sum N, Q] floats."""
    cm = mask_features.coordinate_manager
    sizes = [int(s.shape[0]) for s in seg_logits]  # rows of scatter_mean's output per scene (index.max() + 1, mask3d.py:223)
    key, rowptr, col, val, s_total = _segment_pool_matrix(cm, mask_features.coordinate_map_key, point2segment, sizes, num_pooling_steps)
    seg = torch.cat([s.detach() for s in seg_logits]).float().contiguous()
    assert seg.shape[0] == s_total, f"{seg.shape[0]} segment rows, point2segment addresses {s_total}"
    n, q = rowptr.shape[0] - 1, seg.shape[1]
    bits = torch.empty((n, q), dtype=torch.uint8, device=seg.device)
    check(lib.us3d_pooled_mask_bits(rowptr.data_ptr(), col.data_ptr(), val.data_ptr(), n, seg.data_ptr(), q, bits.data_ptr(), _stream()))
    return key, bits.view(torch.bool)
