"""CPU restatement of the MinkowskiEngine operator surface used by UnScene3D's hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py): pure torch/numpy on CPU, autograd supplies backward.
PARITY UNPINNED at the ME boundary — MinkowskiEngine ≈0.5.4 is an un-vendored dependency
(/root/reference/.devcontainer/Dockerfile:50-51); the semantics restated here are the ones listed
in SURVEY.md Appendix A (A.1–A.15) and each is anchored on the reference call site that relies on
it:

  SparseTensor / CoordinateManager      trainer/trainer.py:115-117, models/mask3d.py:206-209,425-436
  stride map (floor division)           models/res16unet.py:51-59 (conv(..., stride=2))
  kernel offsets / kernel maps          models/modules/common.py:137-155 (KernelGenerator, HYPER_CUBE)
  convolution  Y[o]=sum_k X[o+off_k]W[k]  models/modules/common.py:146-155
  transposed convolution (k2,s2)        models/modules/common.py:179-188, models/res16unet.py:126-204
  batch norm / relu / cat / +=          models/modules/common.py:20-22, models/res16unet.py:222,259,
                                        models/modules/resnet_block.py:61
  avg / sum / max pooling               models/mask3d.py:131,213,432
  sparse_quantize / sparse_collate      datasets/utils.py:266-287,403-432

Algorithm for convolution is the one ME's CPU path uses: per kernel offset, gather the input rows
of the kernel map, one dense GEMM, scatter-add into the output rows.
"""
from __future__ import annotations

import math
from enum import Enum
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

# --------------------------------------------------------------------------------------------
# coordinate keys
# --------------------------------------------------------------------------------------------
_AXIS_BITS = 18
_AXIS_BIAS = 1 << (_AXIS_BITS - 1)


def pack_keys(coords: np.ndarray) -> np.ndarray:
    """(b, x, y, z) int rows -> one int64 key; order-preserving per field."""
    c = coords.astype(np.int64)
    if c.shape[0]:
        assert c[:, 0].min() >= 0 and c[:, 0].max() < (1 << 9), "batch index out of range"
        assert c[:, 1:].min() >= -_AXIS_BIAS and c[:, 1:].max() < _AXIS_BIAS, "coordinate out of range"
    return (
        (c[:, 0] << (3 * _AXIS_BITS))
        | ((c[:, 1] + _AXIS_BIAS) << (2 * _AXIS_BITS))
        | ((c[:, 2] + _AXIS_BIAS) << _AXIS_BITS)
        | (c[:, 3] + _AXIS_BIAS)
    )


def _unique_first_occurrence(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (first_idx, inverse): `first_idx` ascending row indices of the first occurrence of
    every distinct key; `inverse[i]` = position of keys[i] within that list.  (Appendix A.13)"""
    _, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")  # distinct-key id -> rank by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    return first[order], rank[inv.reshape(-1)]


class _KeyIndex:
    """Sorted-key lookup table: key -> row index (or -1)."""

    def __init__(self, keys: np.ndarray):
        self.order = np.argsort(keys, kind="stable")
        self.sorted = keys[self.order]

    def lookup(self, q: np.ndarray) -> np.ndarray:
        if self.sorted.shape[0] == 0:
            return np.full(q.shape, -1, dtype=np.int64)
        pos = np.searchsorted(self.sorted, q)
        pos = np.minimum(pos, self.sorted.shape[0] - 1)
        hit = self.sorted[pos] == q
        return np.where(hit, self.order[pos], -1)


def _tuple3(v, D=3) -> Tuple[int, ...]:
    if isinstance(v, (list, tuple)):
        assert len(v) == D, f"expected {D} entries, got {v}"
        return tuple(int(a) for a in v)
    if isinstance(v, torch.Tensor) or isinstance(v, np.ndarray):
        return tuple(int(a) for a in v)
    return (int(v),) * D


# --------------------------------------------------------------------------------------------
# enums and small public types
# --------------------------------------------------------------------------------------------
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


class CoordinateMapKey:
    def __init__(self, tensor_stride: Sequence[int], string_id: str = ""):
        self.tensor_stride = tuple(int(s) for s in tensor_stride)
        self.string_id = string_id

    def get_tensor_stride(self):
        return list(self.tensor_stride)

    def get_key(self):
        return (list(self.tensor_stride), self.string_id)

    def __hash__(self):
        return hash((self.tensor_stride, self.string_id))

    def __eq__(self, other):
        return (
            isinstance(other, CoordinateMapKey)
            and self.tensor_stride == other.tensor_stride
            and self.string_id == other.string_id
        )

    def __repr__(self):
        return f"CoordinateMapKey(stride={list(self.tensor_stride)}, id='{self.string_id}')"


def kernel_offsets(kernel_size: Sequence[int], tensor_stride: Sequence[int], dilation: Sequence[int]) -> np.ndarray:
    """HYPER_CUBE offsets, axis 0 fastest (Appendix A.4).  odd k: -(k//2)..k//2 ; even k: 0..k-1,
    all scaled by tensor_stride*dilation."""
    D = len(kernel_size)
    vol = int(np.prod(kernel_size))
    offs = np.zeros((vol, D), dtype=np.int64)
    for k in range(vol):
        r = k
        for a in range(D):
            i = r % kernel_size[a]
            r //= kernel_size[a]
            base = i - kernel_size[a] // 2 if kernel_size[a] % 2 == 1 else i
            offs[k, a] = base * tensor_stride[a] * dilation[a]
    return offs


# --------------------------------------------------------------------------------------------
# coordinate manager
# --------------------------------------------------------------------------------------------
class CoordinateManager:
    """Owns every coordinate map (one per tensor stride) and every cached kernel map derived from
    one input SparseTensor (Appendix A.2)."""

    def __init__(self, D: int = 3):
        self.D = D
        self._coords: Dict[CoordinateMapKey, np.ndarray] = {}  # int64 [N, 1+D]
        self._index: Dict[CoordinateMapKey, _KeyIndex] = {}
        self._kmaps: Dict[tuple, List[Tuple[torch.Tensor, torch.Tensor]]] = {}
        self._batch_rows: Dict[CoordinateMapKey, List[torch.Tensor]] = {}

    # -- maps ---------------------------------------------------------------------------
    def insert(self, coords: np.ndarray, tensor_stride=(1, 1, 1), string_id: str = ""):
        key = CoordinateMapKey(tensor_stride, string_id)
        keys = pack_keys(coords)
        first, inverse = _unique_first_occurrence(keys)
        self._coords[key] = coords[first].astype(np.int64)
        self._index[key] = _KeyIndex(keys[first])
        return key, first, inverse

    def size(self, key: CoordinateMapKey) -> int:
        return self._coords[key].shape[0]

    def get_coordinates(self, key: CoordinateMapKey) -> torch.Tensor:
        return torch.from_numpy(self._coords[key].astype(np.int32))

    def exists(self, key: CoordinateMapKey) -> bool:
        return key in self._coords

    def stride(self, in_key: CoordinateMapKey, stride: Sequence[int]) -> CoordinateMapKey:
        """Appendix A.3: c_out = floor(c_in / t_out) * t_out per axis, set-unique,
        canonical row order = first occurrence in input-row order."""
        if all(s == 1 for s in stride):
            return in_key
        t_out = tuple(a * b for a, b in zip(in_key.tensor_stride, stride))
        out_key = CoordinateMapKey(t_out, "")
        if out_key in self._coords:
            return out_key
        c = self._coords[in_key].copy()
        t = np.asarray(t_out, dtype=np.int64)
        c[:, 1:] = np.floor_divide(c[:, 1:], t) * t
        keys = pack_keys(c)
        first, _ = _unique_first_occurrence(keys)
        self._coords[out_key] = c[first]
        self._index[out_key] = _KeyIndex(keys[first])
        return out_key

    # -- kernel maps --------------------------------------------------------------------
    def kernel_map(self, in_key, out_key, kernel_size, dilation=(1, 1, 1)):
        """For each offset k: (in_rows, out_rows) with coords_in[in] == coords_out[out] + off_k.
        Offsets are scaled by the INPUT tensor stride (Appendix A.4)."""
        ck = (in_key, out_key, tuple(kernel_size), tuple(dilation))
        if ck in self._kmaps:
            return self._kmaps[ck]
        offs = kernel_offsets(kernel_size, in_key.tensor_stride, dilation)
        c_out = self._coords[out_key]
        idx_in = self._index[in_key]
        out_rows_all = np.arange(c_out.shape[0], dtype=np.int64)
        maps = []
        for k in range(offs.shape[0]):
            q = c_out.copy()
            q[:, 1:] += offs[k]
            hit = idx_in.lookup(pack_keys(q))
            sel = hit >= 0
            maps.append((torch.from_numpy(hit[sel]), torch.from_numpy(out_rows_all[sel])))
        self._kmaps[ck] = maps
        return maps

    def batch_rows(self, key: CoordinateMapKey) -> List[torch.Tensor]:
        if key not in self._batch_rows:
            b = self._coords[key][:, 0]
            nb = int(b.max()) + 1 if b.shape[0] else 0
            self._batch_rows[key] = [torch.from_numpy(np.nonzero(b == i)[0]) for i in range(nb)]
        return self._batch_rows[key]


# --------------------------------------------------------------------------------------------
# SparseTensor
# --------------------------------------------------------------------------------------------
class SparseTensor:
    """Appendix A.1, A.10, A.11, A.15."""

    def __init__(
        self,
        features: torch.Tensor = None,
        coordinates: torch.Tensor = None,
        tensor_stride=1,
        coordinate_map_key: CoordinateMapKey = None,
        coordinate_manager: CoordinateManager = None,
        quantization_mode=SparseTensorQuantizationMode.RANDOM_SUBSAMPLE,
        minkowski_algorithm=None,
        requires_grad=None,
        device=None,
    ):
        assert isinstance(features, torch.Tensor), "features must be a torch.Tensor"
        if features.ndim == 1:
            features = features[:, None]
        if device is not None:
            features = features.to(device)
        if coordinate_map_key is None:
            assert coordinates is not None, "coordinates or coordinate_map_key required"
            coords = coordinates.detach().cpu().numpy()
            assert coords.ndim == 2 and coords.shape[0] == features.shape[0]
            D = coords.shape[1] - 1
            if coordinate_manager is None:
                coordinate_manager = CoordinateManager(D)
            coordinate_map_key, first, _ = coordinate_manager.insert(np.floor(coords).astype(np.int64), _tuple3(tensor_stride, D))
            if first.shape[0] != features.shape[0]:  # RANDOM_SUBSAMPLE on duplicates
                features = features[torch.from_numpy(first).to(features.device)]
        else:
            assert coordinate_manager is not None
            assert coordinate_manager.size(coordinate_map_key) == features.shape[0], (
                f"feature rows {features.shape[0]} != coordinate map size {coordinate_manager.size(coordinate_map_key)}"
            )
        if requires_grad is not None:
            features.requires_grad_(requires_grad)
        self._F = features
        self.coordinate_map_key = coordinate_map_key
        self.coordinate_manager = coordinate_manager

    # -- accessors ----------------------------------------------------------------------
    @property
    def F(self):
        return self._F

    features = F

    @property
    def C(self):
        return self.coordinate_manager.get_coordinates(self.coordinate_map_key).to(self._F.device)

    coordinates = C

    @property
    def device(self):
        return self._F.device

    @property
    def dtype(self):
        return self._F.dtype

    @property
    def shape(self):
        return self._F.shape

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
        return self._F.size(*a)

    def __len__(self):
        return self._F.shape[0]

    def float(self):
        return self._like(self._F.float())

    def double(self):
        return self._like(self._F.double())

    def detach(self):
        return self._like(self._F.detach())

    def _like(self, feats):
        return SparseTensor(feats, coordinate_map_key=self.coordinate_map_key, coordinate_manager=self.coordinate_manager)

    @property
    def decomposed_features(self) -> List[torch.Tensor]:
        return [self._F[r.to(self._F.device)] for r in self.coordinate_manager.batch_rows(self.coordinate_map_key)]

    @property
    def decomposed_coordinates(self) -> List[torch.Tensor]:
        C = self.C
        return [C[r.to(C.device), 1:] for r in self.coordinate_manager.batch_rows(self.coordinate_map_key)]

    @property
    def decomposed_coordinates_and_features(self):
        return self.decomposed_coordinates, self.decomposed_features

    def dense(self, shape=None, min_coordinate=None, contract_stride=True):
        C = self.C.long()
        ts = torch.tensor(self.tensor_stride, dtype=torch.long)
        mn = C[:, 1:].min(0)[0] if min_coordinate is None else torch.as_tensor(min_coordinate).long().view(-1)
        idx = (C[:, 1:] - mn) // ts if contract_stride else (C[:, 1:] - mn)
        B = int(C[:, 0].max()) + 1
        sz = (idx.max(0)[0] + 1).tolist()
        out = self._F.new_zeros((B, self._F.shape[1], *sz))
        out[C[:, 0], :, idx[:, 0], idx[:, 1], idx[:, 2]] = self._F
        return out, mn[None].int(), ts.int()

    # -- arithmetic on identical keys (A.10) -----------------------------------------------
    def _check(self, other):
        assert isinstance(other, SparseTensor)
        assert self.coordinate_manager is other.coordinate_manager, "different coordinate managers"
        assert self.coordinate_map_key == other.coordinate_map_key, "different coordinate map keys"

    def __add__(self, other):
        if isinstance(other, SparseTensor):
            self._check(other)
            return self._like(self._F + other._F)
        return self._like(self._F + other)

    def __iadd__(self, other):
        if isinstance(other, SparseTensor):
            self._check(other)
            self._F = self._F + other._F
        else:
            self._F = self._F + other
        return self

    def __sub__(self, other):
        if isinstance(other, SparseTensor):
            self._check(other)
            return self._like(self._F - other._F)
        return self._like(self._F - other)

    def __mul__(self, other):
        if isinstance(other, SparseTensor):
            self._check(other)
            return self._like(self._F * other._F)
        return self._like(self._F * other)

    def __repr__(self):
        return f"SparseTensor(F={tuple(self._F.shape)}, key={self.coordinate_map_key})"


TensorField = SparseTensor  # only referenced, never exercised on the hot path


# --------------------------------------------------------------------------------------------
# functional kernels (the arithmetic that the CUDA path must reproduce)
# --------------------------------------------------------------------------------------------
def sparse_conv_forward(feats: torch.Tensor, kernel: torch.Tensor, kmap, n_out: int, transpose_map: bool = False):
    """Y[o] = sum_k X[i_k(o)] @ W[k] — gather, GEMM, scatter-add per kernel offset (Appendix A.5).
    `transpose_map=True` swaps the roles of the map's in/out rows (Appendix A.7)."""
    out = feats.new_zeros((n_out, kernel.shape[-1]))
    for k, (ii, oo) in enumerate(kmap):
        if transpose_map:
            ii, oo = oo, ii
        if ii.numel() == 0:
            continue
        out.index_add_(0, oo, feats.index_select(0, ii) @ kernel[k])
    return out


def sparse_pool_forward(feats: torch.Tensor, kmap, n_out: int, mode: str):
    ii = torch.cat([m[0] for m in kmap])
    oo = torch.cat([m[1] for m in kmap])
    if mode == "max":
        out = feats.new_full((n_out, feats.shape[1]), -float("inf"))
        out = out.index_reduce(0, oo, feats.index_select(0, ii), "amax", include_self=True)
        return torch.where(torch.isinf(out), torch.zeros_like(out), out)
    out = feats.new_zeros((n_out, feats.shape[1]))
    out.index_add_(0, oo, feats.index_select(0, ii))
    if mode == "avg":  # mean over PRESENT inputs (Appendix A.8)
        cnt = torch.zeros(n_out, dtype=feats.dtype).index_add_(0, oo, torch.ones(oo.shape[0], dtype=feats.dtype))
        out = out / cnt.clamp(min=1)[:, None]
    return out


# --------------------------------------------------------------------------------------------
# modules
# --------------------------------------------------------------------------------------------
class KernelGenerator:
    def __init__(self, kernel_size=-1, stride=1, dilation=1, is_transpose=False, region_type=RegionType.HYPER_CUBE,
                 region_offsets=None, expand_coordinates=False, axis_types=None, dimension=-1):
        assert dimension > 0
        assert region_type == RegionType.HYPER_CUBE and (axis_types is None or all(a == RegionType.HYPER_CUBE for a in axis_types)), \
            "only HYPER_CUBE is exercised (SURVEY §8(b))"
        self.dimension = dimension
        self.kernel_size = _tuple3(kernel_size, dimension)
        self.kernel_stride = _tuple3(stride, dimension)
        self.kernel_dilation = _tuple3(dilation, dimension)
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
        self.use_mm = self.kernel_volume == 1 and all(s == 1 for s in self.stride)  # Appendix A.6
        shape = (in_channels, out_channels) if self.use_mm else (self.kernel_volume, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):  # Appendix A.12
        n = (self.out_channels if self.IS_TRANSPOSE else self.in_channels) * self.kernel_volume
        stdv = 1.0 / math.sqrt(n)
        with torch.no_grad():
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def forward(self, x: SparseTensor) -> SparseTensor:
        cm = x.coordinate_manager
        in_key = x.coordinate_map_key
        if self.use_mm:
            out_f, out_key = x.F @ self.kernel, in_key
        elif not self.IS_TRANSPOSE:
            out_key = cm.stride(in_key, self.stride)
            kmap = cm.kernel_map(in_key, out_key, self.kernel_size, self.dilation)
            out_f = sparse_conv_forward(x.F, self.kernel, kmap, cm.size(out_key))
        else:
            ts = in_key.tensor_stride
            assert all(t % s == 0 for t, s in zip(ts, self.stride)), "transposed conv below tensor stride 1"
            out_key = CoordinateMapKey(tuple(t // s for t, s in zip(ts, self.stride)), "")
            if not cm.exists(out_key):
                raise NotImplementedError("transposed convolution that generates new coordinates is not on the hot path")
            # transpose of the forward (fine -> coarse) map (Appendix A.7)
            kmap = cm.kernel_map(out_key, in_key, self.kernel_size, self.dilation)
            out_f = sparse_conv_forward(x.F, self.kernel, kmap, cm.size(out_key), transpose_map=True)
        if self.bias is not None:
            out_f = out_f + self.bias
        return SparseTensor(out_f, coordinate_map_key=out_key, coordinate_manager=cm)


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

    def forward(self, x: SparseTensor) -> SparseTensor:
        cm = x.coordinate_manager
        out_key = cm.stride(x.coordinate_map_key, self.stride)
        kmap = cm.kernel_map(x.coordinate_map_key, out_key, self.kernel_size, self.dilation)
        out_f = sparse_pool_forward(x.F, kmap, cm.size(out_key), self.MODE)
        return SparseTensor(out_f, coordinate_map_key=out_key, coordinate_manager=cm)


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
    """Appendix A.9: nn.BatchNorm1d over all rows of .F."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine, track_running_stats=track_running_stats)

    def forward(self, x: SparseTensor) -> SparseTensor:
        return x._like(self.bn(x.F))


class MinkowskiInstanceNorm(nn.Module):
    def __init__(self, num_features):
        super().__init__()
        self.num_features = num_features
        self.eps = 1e-6
        self.weight = nn.Parameter(torch.ones(1, num_features))
        self.bias = nn.Parameter(torch.zeros(1, num_features))

    def forward(self, x: SparseTensor) -> SparseTensor:
        out = torch.empty_like(x.F)
        for rows in x.coordinate_manager.batch_rows(x.coordinate_map_key):
            f = x.F[rows]
            mean = f.mean(0, keepdim=True)
            var = f.var(0, unbiased=False, keepdim=True)
            out[rows] = (f - mean) / torch.sqrt(var + self.eps)
        return x._like(out * self.weight + self.bias)


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()
        self.inplace = inplace

    def forward(self, x: SparseTensor) -> SparseTensor:
        return x._like(torch.relu(x.F))


def cat(*tensors) -> SparseTensor:
    if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
        tensors = tensors[0]
    for t in tensors[1:]:
        tensors[0]._check(t)
    return tensors[0]._like(torch.cat([t.F for t in tensors], dim=1))


# --------------------------------------------------------------------------------------------
# utils (Appendix A.13, A.14)
# --------------------------------------------------------------------------------------------
def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                    return_inverse=False, return_maps_only=False, quantization_size=None, device="cpu"):
    is_torch = isinstance(coordinates, torch.Tensor)
    c = coordinates.detach().cpu().numpy() if is_torch else np.asarray(coordinates)
    assert c.ndim == 2
    if quantization_size is not None:
        c = c / np.asarray(quantization_size)
    disc = np.floor(c).astype(np.int64)
    keys = pack_keys(np.concatenate([np.zeros((disc.shape[0], 1), np.int64), disc], 1))
    first, inverse = _unique_first_occurrence(keys)
    out_labels = None
    if labels is not None:
        lab = labels.detach().cpu().numpy() if isinstance(labels, torch.Tensor) else np.asarray(labels)
        out_labels = lab[first].copy()
        # voxels hit by points with different labels get ignore_label
        mism = lab != out_labels[inverse]
        out_labels[np.unique(inverse[mism])] = ignore_label
    conv = (lambda a: torch.from_numpy(a)) if is_torch else (lambda a: a)
    umap, imap = conv(first.astype(np.int64)), conv(inverse.astype(np.int64))
    if return_maps_only:
        return (umap, imap) if return_inverse else umap
    ret = [conv(disc[first].astype(np.int32))]
    if features is not None:
        ret.append(features[umap] if isinstance(features, torch.Tensor) else np.asarray(features)[first])
    if labels is not None:
        ret.append(conv(out_labels) if is_torch else out_labels)
    if return_index:
        ret.append(umap)
    if return_inverse:
        ret.append(imap)
    return ret[0] if len(ret) == 1 else tuple(ret)


def batched_coordinates(coords, dtype=torch.int32, device=None):
    out = []
    for b, c in enumerate(coords):
        c = torch.as_tensor(c)
        out.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=dtype), torch.floor(c.double()).to(dtype)], 1))
    res = torch.cat(out, 0) if out else torch.zeros((0, 4), dtype=dtype)
    return res.to(device) if device is not None else res


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    bcoords = batched_coordinates(coords, dtype=dtype, device=device)
    f = torch.cat([torch.as_tensor(x) for x in feats], 0)
    if labels is None:
        return bcoords, f
    return bcoords, f, torch.cat([torch.as_tensor(x) for x in labels], 0)


# --------------------------------------------------------------------------------------------
# registration under the reference's import names (used by tests / fixture scripts only)
# --------------------------------------------------------------------------------------------
def as_module_tree():
    """Builds module objects named `MinkowskiEngine`, `.MinkowskiOps`, `.MinkowskiPooling`, `.utils`
    exporting this oracle, so unmodified reference files can `import MinkowskiEngine as ME`."""
    import types

    g = globals()
    names = [
        "SparseTensor", "TensorField", "CoordinateManager", "CoordinateMapKey", "KernelGenerator", "RegionType",
        "MinkowskiAlgorithm", "SparseTensorQuantizationMode", "MinkowskiNetwork", "MinkowskiConvolution",
        "MinkowskiConvolutionTranspose", "MinkowskiAvgPooling", "MinkowskiSumPooling", "MinkowskiMaxPooling",
        "MinkowskiAvgUnpooling", "MinkowskiBatchNorm", "MinkowskiInstanceNorm", "MinkowskiReLU", "cat",
    ]
    root = types.ModuleType("MinkowskiEngine")
    ops = types.ModuleType("MinkowskiEngine.MinkowskiOps")
    pool = types.ModuleType("MinkowskiEngine.MinkowskiPooling")
    utils = types.ModuleType("MinkowskiEngine.utils")
    for n in names:
        setattr(root, n, g[n])
        setattr(ops, n, g[n])
    for n in ["MinkowskiAvgPooling", "MinkowskiSumPooling", "MinkowskiMaxPooling", "MinkowskiAvgUnpooling"]:
        setattr(pool, n, g[n])
    for n in ["sparse_quantize", "sparse_collate", "batched_coordinates"]:
        setattr(utils, n, g[n])
    root.MinkowskiOps, root.MinkowskiPooling, root.utils = ops, pool, utils
    root.__version__ = "0.5.4-oracle"
    root.__path__ = []
    return {"MinkowskiEngine": root, "MinkowskiEngine.MinkowskiOps": ops,
            "MinkowskiEngine.MinkowskiPooling": pool, "MinkowskiEngine.utils": utils}
