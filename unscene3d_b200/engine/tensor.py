"""SparseTensor container and the nn.Module operator surface of the CUDA backend.

Host-side mirror of the MinkowskiEngine symbols the reference imports; semantics per
Appendix A.  Everything here is bookkeeping — arithmetic is in libus3d via engine/functional.py.
"""
from __future__ import annotations

import math
from enum import Enum
from typing import List

import numpy as np
import torch
import torch.nn as nn

from . import functional as Fn
from .coords import CoordinateManager, CoordinateMapKey, _tuple


class RegionType(Enum):
    HYPER_CUBE = 0
    HYPER_CROSS = 1
    CUSTOM = 2


class MinkowskiAlgorithm(Enum):
    DEFAULT = 0
    MEMORY_EFFICIENT = 1
    SPEED_OPTIMIZED = 2


class SparseTensorQuantizationMode(Enum):
    RANDOM_SUBSAMPLE = 0
    UNWEIGHTED_AVERAGE = 1
    UNWEIGHTED_SUM = 2
    NO_QUANTIZATION = 3
    MAX_POOL = 4
    SPLAT_LINEAR_INTERPOLATION = 5


class SparseTensor:
    """ME.SparseTensor as used at trainer/trainer.py:115-117, models/mask3d.py:206-209, 425-436."""

    def __init__(self, features: torch.Tensor = None, coordinates: torch.Tensor = None, tensor_stride=1,
                 coordinate_map_key: CoordinateMapKey = None, coordinate_manager: CoordinateManager = None,
                 quantization_mode=SparseTensorQuantizationMode.RANDOM_SUBSAMPLE, minkowski_algorithm=None,
                 requires_grad=None, device=None):
        assert isinstance(features, torch.Tensor), "features must be a torch.Tensor"
        if features.ndim == 1:
            features = features[:, None]
        if device is not None:
            features = features.to(device)
        if coordinate_map_key is None:
            assert coordinates is not None, "coordinates or coordinate_map_key required"
            assert coordinates.ndim == 2 and coordinates.shape[0] == features.shape[0], "one coordinate row per feature row"
            dev = features.device
            if dev.type != "cuda":
                raise RuntimeError("unscene3d_b200 SparseTensor lives on a CUDA device: pass device='cuda' "
                                   "(there is no CPU backend; the CPU oracle is test-only)")
            coords = coordinates.to(dev)
            if coords.dtype.is_floating_point:
                coords = torch.floor(coords)
            coords = coords.to(torch.int32)
            if coordinate_manager is None:
                coordinate_manager = CoordinateManager(coords.shape[1] - 1, dev)
            # untouched input rows may be read on the coordinate stream; rows that a conversion kernel just produced on
            # the compute stream may not
            coordinate_map_key, first, _ = coordinate_manager.insert(coords, _tuple(tensor_stride, coords.shape[1] - 1),
                                                                     ready=coords is coordinates)
            if first.shape[0] != features.shape[0]:  # duplicates: keep the first row of each voxel
                features = features[first.long()]
        else:
            assert coordinate_manager is not None
            assert coordinate_manager.size(coordinate_map_key) == features.shape[0], (
                f"feature rows {features.shape[0]} != coordinate map size {coordinate_manager.size(coordinate_map_key)}")
        if requires_grad is not None:
            features.requires_grad_(requires_grad)
        self._feats = features
        self._lazy = None
        self.coordinate_map_key = coordinate_map_key
        self.coordinate_manager = coordinate_manager

    # ---- deferred BatchNorm epilogue -----------------------------------------------------------
    # MinkowskiBatchNorm computes the batch statistics at once but defers the normalisation; a following
    # `+= residual` and MinkowskiReLU are folded into ONE apply pass (y = relu(bn(x) + residual)) — the
    # reference's conv-bn-relu and block-tail patterns (models/modules/resnet_block.py:48-64) become a single
    # read of the conv output.  Touching the features any other way materialises the plain result.
    @classmethod
    def _deferred(cls, lazy, coordinate_map_key, coordinate_manager):
        self = cls.__new__(cls)
        self._feats = None
        self._lazy = lazy
        self.coordinate_map_key = coordinate_map_key
        self.coordinate_manager = coordinate_manager
        return self

    @property
    def _F(self):
        if self._feats is None:
            self._feats = self._lazy.materialize()
            self._lazy = None
        return self._feats

    @_F.setter
    def _F(self, value):
        self._feats = value
        self._lazy = None

    # ---- accessors ---------------------------------------------------------------------------
    @property
    def F(self):
        return self._F

    features = F

    @property
    def C(self):
        return self.coordinate_manager.get_coordinates(self.coordinate_map_key)

    coordinates = C

    # metadata never forces the deferred apply pass
    @property
    def device(self):
        return self._lazy.x.device if self._feats is None else self._feats.device

    @property
    def dtype(self):
        return torch.float32 if self._feats is None else self._feats.dtype

    @property
    def shape(self):
        return self._lazy.x.shape if self._feats is None else self._feats.shape

    @property
    def D(self):
        return self.coordinate_manager.D

    @property
    def tensor_stride(self):
        return self.coordinate_map_key.get_tensor_stride()

    @property
    def requires_grad(self):
        return self._F.requires_grad

    def size(self, *a):
        return self.shape if not a else self.shape[a[0]]

    def __len__(self):
        return self.shape[0]

    def _like(self, feats):
        return SparseTensor(feats, coordinate_map_key=self.coordinate_map_key, coordinate_manager=self.coordinate_manager)

    def float(self):
        return self._like(self._F.float())

    def double(self):
        return self._like(self._F.double())

    def detach(self):
        return self._like(self._F.detach())

    @property
    def decomposed_features(self) -> List[torch.Tensor]:
        return [self._F[s] for s in self.coordinate_manager.batch_slices(self.coordinate_map_key)]

    @property
    def decomposed_coordinates(self) -> List[torch.Tensor]:
        C = self.C
        return [C[s, 1:] for s in self.coordinate_manager.batch_slices(self.coordinate_map_key)]

    @property
    def decomposed_coordinates_and_features(self):
        return self.decomposed_coordinates, self.decomposed_features

    def dense(self, shape=None, min_coordinate=None, contract_stride=True):
        C = self.C.long()
        ts = torch.tensor(self.tensor_stride, dtype=torch.long, device=C.device)
        mn = C[:, 1:].min(0)[0] if min_coordinate is None else torch.as_tensor(min_coordinate, device=C.device).long().view(-1)
        idx = (C[:, 1:] - mn) // ts if contract_stride else (C[:, 1:] - mn)
        B = int(C[:, 0].max()) + 1
        sz = (idx.max(0)[0] + 1).tolist()
        out = self._F.new_zeros((B, self._F.shape[1], *sz))
        out[C[:, 0], :, idx[:, 0], idx[:, 1], idx[:, 2]] = self._F
        return out, mn[None].int(), ts.int()

    # ---- arithmetic on identical keys ---------------------------------------------------------
    def _check(self, other):
        assert isinstance(other, SparseTensor)
        assert self.coordinate_manager is other.coordinate_manager, "different coordinate managers"
        assert self.coordinate_map_key == other.coordinate_map_key, "different coordinate map keys"

    def _add(self, other):
        if isinstance(other, SparseTensor):
            self._check(other)
            if self._feats is None and self._lazy.residual is None and not self._lazy.relu and other._F.dtype == torch.float32:
                return self._lazy.with_residual(other._F)  # stays deferred: bn(x) + residual
            other = other._F
        if isinstance(other, torch.Tensor) and other.shape == self._F.shape and self._F.dtype == torch.float32:
            return Fn.AddFunction.apply(self._F, other)
        return self._F + other

    def __add__(self, other):
        r = self._add(other)
        if isinstance(r, _DeferredBN):
            return SparseTensor._deferred(r, self.coordinate_map_key, self.coordinate_manager)
        return self._like(r)

    def __iadd__(self, other):
        r = self._add(other)
        if isinstance(r, _DeferredBN):
            self._feats, self._lazy = None, r
        else:
            self._F = r
        return self

    def __sub__(self, other):
        if isinstance(other, SparseTensor):
            self._check(other)
            other = other._F
        return self._like(self._F - other)

    def __mul__(self, other):
        if isinstance(other, SparseTensor):
            self._check(other)
            other = other._F
        return self._like(self._F * other)

    def __repr__(self):
        return f"SparseTensor(F={tuple(self._F.shape)}, key={self.coordinate_map_key}, device={self._F.device})"


TensorField = SparseTensor


class _DeferredBN:
    """BatchNorm whose statistics are known and whose apply pass has not run yet."""

    __slots__ = ("x", "weight", "bias", "mean", "invstd", "batch_stats", "residual", "relu", "grad_mode")

    def __init__(self, x, weight, bias, mean, invstd, batch_stats, residual=None, relu=False, grad_mode=None):
        self.x, self.weight, self.bias, self.mean, self.invstd = x, weight, bias, mean, invstd
        self.batch_stats, self.residual, self.relu = batch_stats, residual, relu
        # the apply pass belongs to the autograd context in which BatchNorm was CALLED, not to the one in which the
        # features happen to be touched first (e.g. Mask3D reads aux[-1] inside torch.no_grad(), models/mask3d.py:205)
        self.grad_mode = torch.is_grad_enabled() if grad_mode is None else grad_mode

    def with_residual(self, residual):
        return _DeferredBN(self.x, self.weight, self.bias, self.mean, self.invstd, self.batch_stats, residual, self.relu, self.grad_mode)

    def with_relu(self):
        return _DeferredBN(self.x, self.weight, self.bias, self.mean, self.invstd, self.batch_stats, self.residual, True, self.grad_mode)

    def materialize(self):
        with torch.set_grad_enabled(self.grad_mode):
            return Fn.BatchNormApplyFunction.apply(self.x, self.weight, self.bias, self.residual, self.mean, self.invstd,
                                                   self.batch_stats, self.relu)


# ------------------------------------------------------------------------------------------------
class KernelGenerator:
    def __init__(self, kernel_size=-1, stride=1, dilation=1, is_transpose=False, region_type=RegionType.HYPER_CUBE,
                 region_offsets=None, expand_coordinates=False, axis_types=None, dimension=-1):
        assert dimension > 0
        if region_type != RegionType.HYPER_CUBE or (axis_types is not None and any(a != RegionType.HYPER_CUBE for a in axis_types)):
            raise NotImplementedError("only the HYPER_CUBE region is implemented (the only one the hot path uses)")
        self.dimension = dimension
        self.kernel_size = _tuple(kernel_size, dimension)
        self.kernel_stride = _tuple(stride, dimension)
        self.kernel_dilation = _tuple(dilation, dimension)
        self.region_type = region_type
        self.expand_coordinates = expand_coordinates
        self.kernel_volume = int(np.prod(self.kernel_size))


class MinkowskiNetwork(nn.Module):
    def __init__(self, D):
        super().__init__()
        self.D = D


class MinkowskiModuleBase(nn.Module):
    pass


class _ConvBase(MinkowskiModuleBase):
    IS_TRANSPOSE = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        assert dimension is not None and dimension > 0
        if kernel_generator is None:
            kernel_generator = KernelGenerator(kernel_size, stride, dilation, dimension=dimension)
        self.kernel_generator = kernel_generator
        self.in_channels, self.out_channels, self.dimension = in_channels, out_channels, dimension
        self.kernel_size = kernel_generator.kernel_size
        self.stride = kernel_generator.kernel_stride
        self.dilation = kernel_generator.kernel_dilation
        self.kernel_volume = kernel_generator.kernel_volume
        assert self.kernel_volume <= 27, "kernel volumes up to 3x3x3 are supported"
        self.use_mm = self.kernel_volume == 1 and all(s == 1 for s in self.stride)
        shape = (in_channels, out_channels) if self.use_mm else (self.kernel_volume, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        n = (self.out_channels if self.IS_TRANSPOSE else self.in_channels) * self.kernel_volume
        stdv = 1.0 / math.sqrt(n)
        with torch.no_grad():
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def tables(self, cm, in_key):
        """(output key, forward table, getter of (backward table, flip flag), flip flag the getter will return)."""
        ks, dil = self.kernel_size, self.dilation
        flip_dgrad = False
        if self.use_mm:
            out_key = in_key
            fwd = cm.identity_table(in_key)
            bwd_getter = lambda: (fwd, False)
        elif not self.IS_TRANSPOSE:
            out_key = cm.stride(in_key, self.stride)
            fwd = cm.forward_table(in_key, out_key, ks, dil)
            bwd_getter = lambda: cm.backward_table(in_key, out_key, ks, dil)
            flip_dgrad = in_key == out_key and all(k % 2 == 1 for k in ks)  # == the flag backward_table will return
        else:
            ts = in_key.tensor_stride
            assert all(t % s == 0 for t, s in zip(ts, self.stride)), "transposed conv below tensor stride 1"
            out_key = CoordinateMapKey(tuple(t // s for t, s in zip(ts, self.stride)), "")
            if not cm.exists(out_key):
                raise NotImplementedError("transposed convolution that generates new coordinates is not on the hot path")
            # transpose of the (fine -> coarse) map: forward reads the coarse row at fine - off_k,
            # backward reads the fine rows at coarse + off_k
            fwd = cm.backward_table(out_key, in_key, ks, dil)[0]
            bwd_getter = lambda: (cm.forward_table(out_key, in_key, ks, dil), False)
        return out_key, fwd, bwd_getter, flip_dgrad

    def forward(self, x: SparseTensor) -> SparseTensor:
        cm = x.coordinate_manager
        out_key, fwd, bwd_getter, flip_dgrad = self.tables(cm, x.coordinate_map_key)
        y = Fn.SparseConvFunction.apply(x.F, self.kernel, self.bias, fwd, bwd_getter, flip_dgrad)
        return SparseTensor(y, coordinate_map_key=out_key, coordinate_manager=cm)


class MinkowskiConvolution(_ConvBase):
    IS_TRANSPOSE = False


class MinkowskiConvolutionTranspose(_ConvBase):
    IS_TRANSPOSE = True


class _PoolBase(MinkowskiModuleBase):
    MODE = "avg"

    def __init__(self, kernel_size=-1, stride=1, dilation=1, kernel_generator=None, dimension=None):
        super().__init__()
        assert dimension is not None and dimension > 0
        if kernel_generator is None:
            kernel_generator = KernelGenerator(kernel_size, stride, dilation, dimension=dimension)
        self.kernel_generator = kernel_generator
        self.kernel_size = kernel_generator.kernel_size
        self.stride = kernel_generator.kernel_stride
        self.dilation = kernel_generator.kernel_dilation
        self.dimension = dimension
        if self.kernel_size != self.stride:
            raise NotImplementedError("pooling with kernel_size != stride is not on the hot path")

    def forward(self, x: SparseTensor) -> SparseTensor:
        cm = x.coordinate_manager
        out_key = cm.stride(x.coordinate_map_key, self.stride)
        table = cm.forward_table(x.coordinate_map_key, out_key, self.kernel_size, self.dilation)
        y = Fn.PoolFunction.apply(x.F, table, x.F.shape[0], Fn.PoolFunction.MODES[self.MODE])
        return SparseTensor(y, coordinate_map_key=out_key, coordinate_manager=cm)


class MinkowskiAvgPooling(_PoolBase):
    MODE = "avg"


class MinkowskiSumPooling(_PoolBase):
    MODE = "sum"


class MinkowskiMaxPooling(_PoolBase):
    MODE = "max"


class MinkowskiAvgUnpooling(_PoolBase):
    def forward(self, x):
        raise NotImplementedError("MinkowskiAvgUnpooling is constructed by models/modules/common.py:222 but never called")


class MinkowskiBatchNorm(nn.Module):
    """nn.BatchNorm1d parameters/buffers under `.bn` (checkpoint names), our kernels for the math."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine, track_running_stats=track_running_stats)

    def statistics(self, f: torch.Tensor):
        """(mean, invstd, batch statistics?) for the rows of `f`; updates the running statistics like nn.BatchNorm1d."""
        bn = self.bn
        use_batch = bn.training or not bn.track_running_stats
        momentum = bn.momentum
        nbt = None
        if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
            if momentum is None:  # cumulative moving average: the factor depends on the counter (host read, rare)
                bn.num_batches_tracked.add_(1)
                momentum = 1.0 / float(bn.num_batches_tracked)
            else:
                nbt = bn.num_batches_tracked  # incremented by the statistics kernel itself
        rm = bn.running_mean if (bn.track_running_stats and (bn.training or not use_batch)) else None
        rv = bn.running_var if rm is not None else None
        if use_batch:
            mean, invstd = Fn.bn_batch_stats(f.detach(), rm, rv, momentum, bn.eps, nbt)
        else:
            mean = bn.running_mean.detach().float()
            invstd = torch.rsqrt(bn.running_var.detach().float() + bn.eps)
        return mean, invstd, use_batch

    def stats_request(self):
        """The batch-statistics request a producing convolution can fold into its epilogue (Fn.BnRequest), or None when the
        statistics are not this batch's (evaluation mode) or the momentum is the cumulative average (host-side counter)."""
        bn = self.bn
        if not (bn.training or not bn.track_running_stats):
            return None
        tracking = bn.training and bn.track_running_stats
        if tracking and bn.momentum is None:
            return None
        nbt = bn.num_batches_tracked if (tracking and bn.num_batches_tracked is not None) else None
        rm = bn.running_mean if tracking else None
        return Fn.BnRequest(rm, bn.running_var if rm is not None else None, bn.momentum, bn.eps, nbt)

    def forward(self, x: SparseTensor, residual: SparseTensor = None, relu: bool = False) -> SparseTensor:
        bn = self.bn
        f = x.F
        mean, invstd, use_batch = self.statistics(f)
        lazy = _DeferredBN(f, bn.weight, bn.bias, mean, invstd, use_batch, None if residual is None else residual.F, relu)
        return SparseTensor._deferred(lazy, x.coordinate_map_key, x.coordinate_manager)


class MinkowskiInstanceNorm(nn.Module):
    def __init__(self, num_features):
        super().__init__()
        self.num_features = num_features
        self.eps = 1e-6
        self.weight = nn.Parameter(torch.ones(1, num_features))
        self.bias = nn.Parameter(torch.zeros(1, num_features))

    def forward(self, x: SparseTensor) -> SparseTensor:
        outs = []
        for f in x.decomposed_features:  # per-instance statistics through the same BN kernels
            outs.append(Fn.BatchNormFunction.apply(f, None, None, None, None, None, 0.0, self.eps, True, False))
        y = torch.cat(outs, 0) if outs else x.F
        return x._like(y * self.weight + self.bias)


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()
        self.inplace = inplace

    def forward(self, x: SparseTensor) -> SparseTensor:
        if x._feats is None and not x._lazy.relu:  # fold into the deferred BatchNorm apply
            return SparseTensor._deferred(x._lazy.with_relu(), x.coordinate_map_key, x.coordinate_manager)
        f = x.F
        inplace = self.inplace and f.is_contiguous() and f.dtype == torch.float32 and not (f.requires_grad and f.is_leaf) \
            and f._base is None
        return x._like(Fn.ReLUFunction.apply(f, inplace))


def cat(*tensors) -> SparseTensor:
    if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
        tensors = tensors[0]
    for t in tensors[1:]:
        tensors[0]._check(t)
    return tensors[0]._like(Fn.CatFunction.apply(*[t.F for t in tensors]))
