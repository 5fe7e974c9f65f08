"""Device-side coordinate manager: coordinate maps (one per tensor stride) and neighbour tables.

Host-side mirror of MinkowskiEngine's CoordinateManager as the reference uses it (SURVEY.md
Appendix A.1–A.4, A.7): one manager per input SparseTensor, maps are created on first use and shared
by every layer / pooling op / backward pass of that level (trainer/trainer.py:115-117,
models/mask3d.py:206-215, 425-436).  All integer work runs in libus3d (csrc/coords.cu).

Layout in HBM
  coordinate map   int32 [N, 4] rows (b, x, y, z); hash table uint64 keys[cap] + int32 vals[cap]
  neighbour table  int32 [K, N_out]  (nbr[k, o] = input row at coords_out[o] + off_k, or -1)
                   + uint32 tile mask [ceil(N_out/128)] (bit k: any neighbour at offset k in the tile)
"""
from __future__ import annotations

import contextlib
import ctypes
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .._lib import check, lib

TILE_ROWS = 128
# strided coordinate maps (x2 per level) derived when a coordinate map is inserted; Res16UNet uses strides 2..16
EAGER_STRIDE_LEVELS = int(os.environ.get("US3D_EAGER_STRIDES", "4"))


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream (the raw getter is ~50x cheaper than torch.cuda.current_stream())."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _tuple(v, D=3) -> Tuple[int, ...]:
    if isinstance(v, (list, tuple)):
        assert len(v) == D, f"expected {D} entries, got {v}"
        return tuple(int(a) for a in v)
    if isinstance(v, (torch.Tensor, np.ndarray)):
        return tuple(int(a) for a in v)
    return (int(v),) * D


class CoordinateMapKey:
    __slots__ = ("tensor_stride", "string_id")

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
        return (isinstance(other, CoordinateMapKey) and self.tensor_stride == other.tensor_stride
                and self.string_id == other.string_id)

    def __repr__(self):
        return f"CoordinateMapKey(stride={list(self.tensor_stride)}, id='{self.string_id}')"


def kernel_offsets(kernel_size: Sequence[int], tensor_stride: Sequence[int], dilation: Sequence[int]) -> np.ndarray:
    """HYPER_CUBE offsets, x fastest; odd k centred, even k starts at 0 (Appendix A.4)."""
    ks = np.asarray(kernel_size)
    vol = int(ks.prod())
    idx = np.arange(vol)
    out = np.zeros((vol, len(ks)), dtype=np.int32)
    for a, k in enumerate(ks):
        i = idx % k
        idx = idx // k
        out[:, a] = (i - k // 2 if k % 2 else i) * tensor_stride[a] * dilation[a]
    return out


class CoordinateMap:
    """Unique coordinates of one tensor stride + the hash table that indexes them."""

    def __init__(self, coords: torch.Tensor, keys: torch.Tensor, vals: torch.Tensor):
        self.coords, self.keys, self.vals = coords, keys, vals
        self.n = coords.shape[0]
        self.cap = keys.shape[0]
        self._batch_slices: Optional[List] = None
        self._side = None  # the coordinate stream the map was built on (the host has synchronised with it), if any


# Tables with at least this many rows are also kept in neighbour-pattern order for the tcgen05 kernels (0 disables)
_row_order = {"min_rows": int(os.environ.get("US3D_ROW_ORDER_MIN", "32768"))}


def set_row_ordering(min_rows: int):
    """Tables with >= min_rows rows (and more than one kernel offset) are handed to the tensor-core kernels in
    neighbour-pattern order (see NeighbourTable.ordered); 0 switches the re-ordering off."""
    _row_order["min_rows"] = int(min_rows)


class NeighbourTable:
    __slots__ = ("nbr", "mask", "n_rows", "kvol", "_pairs", "_ordered", "_part")

    def __init__(self, nbr: torch.Tensor, mask: torch.Tensor, n_rows: int, kvol: int):
        self.nbr, self.mask, self.n_rows, self.kvol, self._pairs, self._ordered, self._part = nbr, mask, n_rows, kvol, None, None, None

    def partition(self):
        """Cost-weighted split of the (pattern-ordered) table's 128-row tiles over the SMs: int32 [SMs + 1] tile boundaries such
        that every CTA's contiguous range carries the same number of active (tile, offset) products.  Tiles of a pattern-ordered
        table differ by up to 27x in cost (the rarest patterns sort first), so equal COUNTS leave the CTAs unbalanced.  One tiny
        kernel, built with the order and cached."""
        if self._part is None:
            o = self.ordered()
            if not o:
                self._part = False
            else:
                part = torch.empty(int(lib.us3d_spconv_partition_size()), dtype=torch.int32, device=self.nbr.device)
                n_tiles = (self.n_rows + TILE_ROWS - 1) // TILE_ROWS
                check(lib.us3d_spconv_partition(o[1].data_ptr(), n_tiles, self.kvol, part.data_ptr(), _stream()))
                self._part = part
        return None if self._part is False else self._part

    def ordered(self):
        """(nbr, tile mask, order) with the table's rows grouped by neighbour pattern, or None for small tables.

        The tensor-core kernels skip a kernel offset for a 128-row tile only if no row of the tile has a neighbour there.
        Spatially consecutive rows of a voxelised surface keep every offset active (200k-voxel scene: 100 % of the
        (tile, offset) pairs at a pair density of 49 %); rows sorted by presence pattern, rarest offset most significant,
        leave 67 % (k3) and 1/8 (the fine side of a k2s2 map: every row has exactly one parent).  Column j of the
        re-ordered table belongs to row order[j]; results are written to / read from their natural rows, so nothing
        outside the convolution kernels sees the order.  Built on first use: two integer kernels + one stable sort."""
        if self._ordered is None:
            m = _row_order["min_rows"]
            if m <= 0 or self.n_rows < m or self.kvol == 1:
                self._ordered = False
            else:
                dev = self.nbr.device
                n, kvol, st = self.n_rows, self.kvol, _stream()
                scratch = torch.empty(n + 32, dtype=torch.int32, device=dev)
                keys = torch.empty(n, dtype=torch.int32, device=dev)
                check(lib.us3d_neighbour_pattern_keys(self.nbr.data_ptr(), n, kvol, scratch.data_ptr(), keys.data_ptr(), st))
                order = torch.sort(keys, stable=True)[1].to(torch.int32)
                nbr = torch.empty_like(self.nbr)
                mask = torch.empty_like(self.mask)
                check(lib.us3d_kernel_map_reorder(self.nbr.data_ptr(), n, kvol, order.data_ptr(), nbr.data_ptr(), mask.data_ptr(),
                                                  TILE_ROWS, st))
                self._ordered = (nbr, mask, order)
        return self._ordered or None

    def pairs(self) -> int:
        """Present (input row, output row) pairs — the P of the algorithmic FLOP count 2 P Cin Cout (host read, cached;
        used by the benchmark's roofline leg only)."""
        if self._pairs is None:
            self._pairs = int((self.nbr >= 0).sum())
        return self._pairs


_coordinate_streams: Dict[int, "torch.cuda.Stream"] = {}
_SLICES_ON_SIDE = os.environ.get("US3D_SLICES_ON_SIDE", "1") == "1"  # batch_slices' host reads on the coordinate stream


def set_coordinate_stream(stream: Optional["torch.cuda.Stream"], device=None):
    """Build coordinate maps on `stream` (None = on the current stream, the default).

    De-duplicating a coordinate map hands its size back to the host — a stream synchronisation.  On the compute stream
    that synchronisation waits for everything queued there, i.e. for the previous step's backward pass, and the host
    cannot start queueing the next step until the device has drained.  With a dedicated (high-priority) coordinate stream
    the host waits only for the few integer kernels of the map itself, so steps pipeline like they do behind a data
    loader's prefetch stream.  Contract: coordinates handed to SparseTensor / CoordinateManager.insert must be complete
    with respect to `stream` (resident tensors, or copies issued on it); anything that first needs a conversion kernel
    falls back to the compute stream."""
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    if stream is None:
        _coordinate_streams.pop(dev, None)
    else:
        _coordinate_streams[dev] = stream


def get_coordinate_stream(device=None):
    dev = torch.cuda.current_device() if device is None else torch.device(device).index
    return _coordinate_streams.get(dev)


def unique_coords(coords: torch.Tensor, tensor_stride=(1, 1, 1), side_ok: bool = False):
    """libus3d us3d_coords_unique on a CUDA int32 [n,4] tensor.
    Returns (CoordinateMap, first_rows int32 [m], inverse int32 [n])."""
    assert coords.is_cuda and coords.dtype == torch.int32 and coords.ndim == 2 and coords.shape[1] == 4
    side = _coordinate_streams.get(coords.device.index) if side_ok and coords.is_contiguous() else None
    if side is None:
        return _unique_coords_on_current(coords.contiguous(), tensor_stride)
    main = torch.cuda.current_stream(coords.device)
    with torch.cuda.stream(side):
        cmap, first, inverse = _unique_coords_on_current(coords, tensor_stride)
    cmap._side = side
    # the maps are consumed by kernels of the compute stream: their memory must not be recycled by the coordinate
    # stream's allocator pool while those kernels are pending
    for t in (cmap.coords, cmap.keys, cmap.vals, first, inverse):
        t.record_stream(main)
    return cmap, first, inverse


def _unique_coords_on_current(coords: torch.Tensor, tensor_stride):
    n = coords.shape[0]
    dev = coords.device
    cap = lib.us3d_hash_capacity(n)
    keys = torch.empty(cap, dtype=torch.int64, device=dev)
    vals = torch.empty(cap, dtype=torch.int32, device=dev)
    out = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dev)
    first = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    inverse = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    scratch = torch.empty(2 * n + 4200, dtype=torch.int32, device=dev)
    count = ctypes.c_int(0)
    check(lib.us3d_coords_unique(coords.data_ptr(), n, int(tensor_stride[0]), int(tensor_stride[1]), int(tensor_stride[2]),
                                 keys.data_ptr(), vals.data_ptr(), cap, out.data_ptr(), first.data_ptr(),
                                 inverse.data_ptr(), scratch.data_ptr(), ctypes.byref(count), _stream()))
    m = count.value
    return CoordinateMap(out[:m], keys, vals), first[:m], inverse[:n]


class CoordinateManager:
    def __init__(self, D: int = 3, device=None):
        assert D == 3, "the hot path is 3-D (SURVEY §8(b))"
        self.D = D
        self.device = device
        self._maps: Dict[CoordinateMapKey, CoordinateMap] = {}
        self._parents: Dict[Tuple[CoordinateMapKey, CoordinateMapKey], torch.Tensor] = {}
        self._tables: Dict[tuple, NeighbourTable] = {}
        self._identity: Dict[CoordinateMapKey, NeighbourTable] = {}

    # ---- coordinate maps --------------------------------------------------------------------
    def insert(self, coords: torch.Tensor, tensor_stride=(1, 1, 1), string_id: str = "", ready: bool = True):
        """`ready`: the coordinate rows are complete with respect to the coordinate stream (see set_coordinate_stream)."""
        key = CoordinateMapKey(tensor_stride, string_id)
        cmap, first, inverse = unique_coords(coords, (1, 1, 1), side_ok=ready)
        self._maps[key] = cmap
        self.device = coords.device
        # De-duplicating a map returns its size to the host (one stream synchronisation).  The strided maps of the
        # U-Net pyramid are therefore derived right away, while the stream is still empty, instead of on first use in
        # the middle of the forward pass — where every synchronisation would drain the launches the host has queued
        # ahead of the device and serialise the host-bound coarse levels with the device-bound fine ones.
        k = key
        for _ in range(EAGER_STRIDE_LEVELS if string_id == "" else 0):
            if self._maps[k].n <= 1:
                break
            k = self.stride(k, (2, 2, 2))
        return key, first, inverse

    def exists(self, key: CoordinateMapKey) -> bool:
        return key in self._maps

    def size(self, key: CoordinateMapKey) -> int:
        return self._maps[key].n

    def get_coordinates(self, key: CoordinateMapKey) -> torch.Tensor:
        return self._maps[key].coords

    def stride(self, in_key: CoordinateMapKey, stride: Sequence[int]) -> CoordinateMapKey:
        """c_out = floor(c_in / t_out) * t_out, unique in first-occurrence order (Appendix A.3)."""
        if all(s == 1 for s in stride):
            return in_key
        t_out = tuple(a * b for a, b in zip(in_key.tensor_stride, stride))
        out_key = CoordinateMapKey(t_out, "")
        if out_key not in self._maps:
            # a map of this manager is complete on whichever stream built it (the host has synchronised with it)
            cmap, _, inverse = unique_coords(self._maps[in_key].coords, t_out, side_ok=True)
            self._maps[out_key] = cmap
            self._parents[(in_key, out_key)] = inverse
        return out_key

    def batch_slices(self, key: CoordinateMapKey):
        """Per-batch row selectors: slices when rows are batch-sorted (always true for collated input
        and every map derived from it), index tensors otherwise."""
        cmap = self._maps[key]
        if cmap._batch_slices is None:
            b = cmap.coords[:, 0]
            if cmap.n == 0:
                cmap._batch_slices = []
            else:
                # Three host reads.  On the compute stream each of them waits for everything queued there — in Mask3D that is the
                # whole backbone forward, after which the device idles while the host queues the decoder.  A map built on the
                # coordinate stream is complete there (the host has synchronised with it), so the reads go to that stream.
                side = cmap._side if _SLICES_ON_SIDE else None
                main = torch.cuda.current_stream(b.device)
                with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                    counts = torch.bincount(b.long())
                    sorted_ok = bool((b[1:] >= b[:-1]).all()) if cmap.n > 1 else True
                    if sorted_ok:
                        ends = torch.cumsum(counts, 0).tolist()
                        starts = [0] + ends[:-1]
                        cmap._batch_slices = [slice(s, e) for s, e in zip(starts, ends)]
                    else:
                        sel = [torch.nonzero(b == i).flatten() for i in range(counts.shape[0])]
                        if side is not None:
                            side.synchronize()  # index tensors consumed on the compute stream
                            for t in sel:
                                t.record_stream(main)
                        cmap._batch_slices = sel
        return cmap._batch_slices

    # ---- neighbour tables ------------------------------------------------------------------
    def _build_table(self, query_key, table_key, offsets: np.ndarray) -> NeighbourTable:
        q, t = self._maps[query_key], self._maps[table_key]
        kvol = offsets.shape[0]
        nbr = torch.empty((kvol, q.n), dtype=torch.int32, device=q.coords.device)
        mask = torch.empty(max((q.n + TILE_ROWS - 1) // TILE_ROWS, 1), dtype=torch.int32, device=q.coords.device)
        offs = np.ascontiguousarray(offsets, dtype=np.int32)
        check(lib.us3d_kernel_map(q.coords.data_ptr(), q.n, offs.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), kvol,
                                  t.keys.data_ptr(), t.vals.data_ptr(), t.cap, nbr.data_ptr(), mask.data_ptr(),
                                  TILE_ROWS, _stream()))
        return NeighbourTable(nbr, mask, q.n, kvol)

    def forward_table(self, in_key, out_key, kernel_size, dilation=(1, 1, 1)) -> NeighbourTable:
        """nbr[k, o] = in row at coords_out[o] + off_k  (offsets scaled by the INPUT stride)."""
        ck = ("f", in_key, out_key, tuple(kernel_size), tuple(dilation))
        if ck not in self._tables:
            offs = kernel_offsets(kernel_size, in_key.tensor_stride, dilation)
            self._tables[ck] = self._build_table(out_key, in_key, offs)
        return self._tables[ck]

    def backward_table(self, in_key, out_key, kernel_size, dilation=(1, 1, 1)) -> Tuple[NeighbourTable, bool]:
        """Transposed map: nbrT[k, i] = out row at coords_in[i] - off_k.  Returns (table, flip_k): for an
        odd kernel on one map the forward table read with k -> K-1-k IS the transposed map."""
        same = in_key == out_key and all(k % 2 == 1 for k in kernel_size)
        if same:
            return self.forward_table(in_key, out_key, kernel_size, dilation), True
        ck = ("b", in_key, out_key, tuple(kernel_size), tuple(dilation))
        if ck not in self._tables:
            offs = -kernel_offsets(kernel_size, in_key.tensor_stride, dilation)
            self._tables[ck] = self._build_table(in_key, out_key, offs)
        return self._tables[ck], False

    def identity_table(self, key) -> NeighbourTable:
        if key not in self._identity:
            n = self._maps[key].n
            nbr = torch.arange(n, dtype=torch.int32, device=self._maps[key].coords.device)[None]
            mask = torch.ones(max((n + TILE_ROWS - 1) // TILE_ROWS, 1), dtype=torch.int32, device=nbr.device)
            self._identity[key] = NeighbourTable(nbr, mask, n, 1)
        return self._identity[key]
