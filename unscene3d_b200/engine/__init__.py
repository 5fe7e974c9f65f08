"""CUDA backend behind the MinkowskiEngine-compatible operator surface."""
from .coords import (CoordinateManager, CoordinateMapKey, NeighbourTable, get_coordinate_stream, kernel_offsets, set_coordinate_stream,
                     set_row_ordering, unique_coords)
from .tensor import (KernelGenerator, MinkowskiAlgorithm, MinkowskiAvgPooling, MinkowskiAvgUnpooling, MinkowskiBatchNorm,
                     MinkowskiConvolution, MinkowskiConvolutionTranspose, MinkowskiInstanceNorm, MinkowskiMaxPooling,
                     MinkowskiNetwork, MinkowskiReLU, MinkowskiSumPooling, RegionType, SparseTensor,
                     SparseTensorQuantizationMode, TensorField, cat)
from . import functional
from .utils import batched_coordinates, sparse_collate, sparse_quantize
