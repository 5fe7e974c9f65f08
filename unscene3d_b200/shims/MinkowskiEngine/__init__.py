"""Drop-in `MinkowskiEngine` module backed by the sm_100a kernels of unscene3d_b200.

Exposes exactly the symbols the reference imports; put
`unscene3d_b200/shims` on sys.path (unscene3d_b200.install_shims()) and the reference's
models/*.py run unchanged on this backend.
"""
from unscene3d_b200.engine import (CoordinateManager, CoordinateMapKey, KernelGenerator, MinkowskiAlgorithm,
                                   MinkowskiAvgPooling, MinkowskiAvgUnpooling, MinkowskiBatchNorm, MinkowskiConvolution,
                                   MinkowskiConvolutionTranspose, MinkowskiInstanceNorm, MinkowskiMaxPooling,
                                   MinkowskiNetwork, MinkowskiReLU, MinkowskiSumPooling, RegionType, SparseTensor,
                                   SparseTensorQuantizationMode, TensorField, cat)
from . import MinkowskiOps, MinkowskiPooling, utils

__version__ = "0.5.4+us3d"
