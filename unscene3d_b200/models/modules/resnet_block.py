"""Residual blocks of the sparse backbone — mirror of the reference's
models/modules/resnet_block.py (BasicBlockBase :7-64, BottleneckBase :79-137).

Attribute names (conv1/norm1/conv2/norm2[/conv3/norm3]/downsample) are the reference's, so
checkpoints written by either implementation load into the other.
"""
import torch.nn as nn
from MinkowskiEngine import MinkowskiReLU

from .common import ConvType, NormType, conv, get_norm


def _cuda_block_core(block, x):
    from unscene3d_b200.engine.blocks import fused_basic_block

    return fused_basic_block(block, x)


class _ResidualBase(nn.Module):
    # Optional whole-block execution (unscene3d_b200.engine.blocks.fused_basic_block); returns None to fall through to the
    # reference's module-by-module sequence below.  The CPU tests that run these definitions over the oracle install a core
    # that always returns None (tests/helpers.py).
    block_core = None
    expansion = 1
    NORM_TYPE = NormType.BATCH_NORM
    # (attribute suffix, kernel size, uses the block's stride/dilation/conv_type, output multiplier)
    STAGES = ()

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None,
                 conv_type=ConvType.HYPERCUBE, bn_momentum=0.1, D=3):
        super().__init__()
        cin = inplanes
        for suffix, ksize, shaped, mult in self.STAGES:
            cout = planes * mult
            if shaped:
                layer = conv(cin, cout, kernel_size=ksize, stride=stride if suffix == self.STRIDED else 1,
                             dilation=dilation, conv_type=conv_type, D=D)
            else:
                layer = conv(cin, cout, kernel_size=ksize, D=D)
            setattr(self, f"conv{suffix}", layer)
            setattr(self, f"norm{suffix}", get_norm(self.NORM_TYPE, cout, D, bn_momentum=bn_momentum))
            cin = cout
        self.relu = MinkowskiReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        core = type(self).block_core or _cuda_block_core
        fused = core(self, x)  # one autograd node for the whole block where the engine offers it (same kernels, same order)
        if fused is not None:
            return fused
        out = x
        last = self.STAGES[-1][0]
        for suffix, _, _, _ in self.STAGES:
            out = getattr(self, f"norm{suffix}")(getattr(self, f"conv{suffix}")(out))
            if suffix != last:
                out = self.relu(out)
        shortcut = x if self.downsample is None else self.downsample(x)
        out += shortcut
        return self.relu(out)


class BasicBlockBase(_ResidualBase):
    expansion = 1
    STRIDED = "1"
    STAGES = (("1", 3, True, 1), ("2", 3, True, 1))


class BasicBlock(BasicBlockBase):
    NORM_TYPE = NormType.BATCH_NORM


class BasicBlockIN(BasicBlockBase):
    NORM_TYPE = NormType.INSTANCE_NORM


class BasicBlockINBN(BasicBlockBase):
    NORM_TYPE = NormType.INSTANCE_BATCH_NORM


class BottleneckBase(_ResidualBase):
    expansion = 4
    STRIDED = "2"
    STAGES = (("1", 1, False, 1), ("2", 3, True, 1), ("3", 1, False, 4))


class Bottleneck(BottleneckBase):
    NORM_TYPE = NormType.BATCH_NORM


class BottleneckIN(BottleneckBase):
    NORM_TYPE = NormType.INSTANCE_NORM


class BottleneckINBN(BottleneckBase):
    NORM_TYPE = NormType.INSTANCE_BATCH_NORM
