"""Layer factories of the sparse backbone — host-side mirror of the reference's
models/modules/common.py (conv :125-155, conv_tr :158-188, get_norm :20-31, pooling :191-258).

Same names, arguments and meaning; whatever module is importable as ``MinkowskiEngine`` supplies
the operators (in the product that is unscene3d_b200/shims/MinkowskiEngine → sm_100a kernels).
Only the HYPER_CUBE region in 3-D is on the hot path; the other ConvType members
exist so that configs naming them still parse, and collapse onto the cube exactly like the
reference's table (:58-67) does.
"""
from collections.abc import Sequence
from enum import Enum

import torch.nn as nn
import MinkowskiEngine as ME


class NormType(Enum):
    BATCH_NORM = 0
    INSTANCE_NORM = 1
    INSTANCE_BATCH_NORM = 2


class ConvType(Enum):
    HYPERCUBE = 0
    SPATIAL_HYPERCUBE = 1
    SPATIO_TEMPORAL_HYPERCUBE = 2
    HYPERCROSS = 3
    SPATIAL_HYPERCROSS = 4
    SPATIO_TEMPORAL_HYPERCROSS = 5
    SPATIAL_HYPERCUBE_TEMPORAL_HYPERCROSS = 6

    def __int__(self):
        return self.value


_CROSS = {ConvType.HYPERCROSS, ConvType.SPATIAL_HYPERCROSS, ConvType.SPATIO_TEMPORAL_HYPERCROSS}
_SPATIAL_ONLY = {ConvType.SPATIAL_HYPERCUBE, ConvType.SPATIAL_HYPERCROSS}


def get_norm(norm_type, n_channels, D, bn_momentum=0.1):
    if norm_type == NormType.BATCH_NORM:
        return ME.MinkowskiBatchNorm(n_channels, momentum=bn_momentum)
    if norm_type == NormType.INSTANCE_NORM:
        return ME.MinkowskiInstanceNorm(n_channels)
    if norm_type == NormType.INSTANCE_BATCH_NORM:
        return nn.Sequential(ME.MinkowskiInstanceNorm(n_channels), ME.MinkowskiBatchNorm(n_channels, momentum=bn_momentum))
    raise ValueError(f"Norm type: {norm_type} not supported")


def convert_conv_type(conv_type, kernel_size, D):
    """(ConvType, kernel_size) -> (ME.RegionType, axis_types, kernel_size as the generator wants it)."""
    assert isinstance(conv_type, ConvType), "conv_type must be of ConvType"
    region = ME.RegionType.HYPER_CROSS if conv_type in _CROSS else ME.RegionType.HYPER_CUBE
    axis_types = None
    if conv_type in _SPATIAL_ONLY:
        kernel_size = list(kernel_size[:3]) if isinstance(kernel_size, Sequence) else [kernel_size] * 3
        if D == 4:
            kernel_size.append(1)
    elif conv_type in (ConvType.SPATIO_TEMPORAL_HYPERCUBE, ConvType.SPATIO_TEMPORAL_HYPERCROSS):
        assert D == 4
    elif conv_type == ConvType.SPATIAL_HYPERCUBE_TEMPORAL_HYPERCROSS:
        axis_types = [ME.RegionType.HYPER_CUBE] * 3 + ([ME.RegionType.HYPER_CROSS] if D == 4 else [])
    return region, axis_types, kernel_size


def _generator(kernel_size, stride, dilation, conv_type, D, keep_axis_types=True):
    assert D > 0, "Dimension must be a positive integer"
    region, axis_types, kernel_size = convert_conv_type(conv_type, kernel_size, D)
    gen = ME.KernelGenerator(kernel_size, stride, dilation, region_type=region,
                             axis_types=axis_types if keep_axis_types else None, dimension=D)
    return gen, kernel_size


def conv(in_planes, out_planes, kernel_size, stride=1, dilation=1, bias=False, conv_type=ConvType.HYPERCUBE, D=-1):
    # the reference drops axis_types for plain convs (models/modules/common.py:141)
    gen, kernel_size = _generator(kernel_size, stride, dilation, conv_type, D, keep_axis_types=False)
    return ME.MinkowskiConvolution(in_channels=in_planes, out_channels=out_planes, kernel_size=kernel_size,
                                   stride=stride, dilation=dilation, bias=bias, kernel_generator=gen, dimension=D)


def conv_tr(in_planes, out_planes, kernel_size, upsample_stride=1, dilation=1, bias=False,
            conv_type=ConvType.HYPERCUBE, D=-1):
    gen, kernel_size = _generator(kernel_size, upsample_stride, dilation, conv_type, D)
    return ME.MinkowskiConvolutionTranspose(in_channels=in_planes, out_channels=out_planes, kernel_size=kernel_size,
                                            stride=upsample_stride, dilation=dilation, bias=bias,
                                            kernel_generator=gen, dimension=D)


def _pool(cls, kernel_size, stride, dilation, conv_type, D):
    gen, kernel_size = _generator(kernel_size, stride, dilation, conv_type, D)
    return cls(kernel_size=kernel_size, stride=stride, dilation=dilation, kernel_generator=gen, dimension=D)


def avg_pool(kernel_size, stride=1, dilation=1, conv_type=ConvType.HYPERCUBE, in_coords_key=None, D=-1):
    return _pool(ME.MinkowskiAvgPooling, kernel_size, stride, dilation, conv_type, D)


def avg_unpool(kernel_size, stride=1, dilation=1, conv_type=ConvType.HYPERCUBE, D=-1):
    return _pool(ME.MinkowskiAvgUnpooling, kernel_size, stride, dilation, conv_type, D)


def sum_pool(kernel_size, stride=1, dilation=1, conv_type=ConvType.HYPERCUBE, D=-1):
    return _pool(ME.MinkowskiSumPooling, kernel_size, stride, dilation, conv_type, D)
