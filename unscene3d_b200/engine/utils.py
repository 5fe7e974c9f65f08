"""ME.utils.sparse_quantize / sparse_collate / batched_coordinates (SURVEY.md Appendix A.13, A.14;
call sites datasets/utils.py:266-287, 403-432).

The reference runs these on the CPU inside forked DataLoader workers (conf/data/indoor.yaml:24), where a CUDA context
must not be created: host inputs (numpy arrays, CPU tensors) are de-duplicated by a HOST function of libus3d
(us3d_coords_unique_h, sequential open-addressing hash — the algorithm of ME's CPU path), CUDA tensors by the hash
kernels that build coordinate maps (csrc/coords.cu).  Same results either way: unique rows in first-occurrence order.
Inputs may be numpy arrays or CPU/CUDA torch tensors; results come back in the input's flavour.
"""
import numpy as np
import torch

from .._lib import check, lib
from .coords import unique_coords


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                    return_inverse=False, return_maps_only=False, quantization_size=None, device="cpu"):
    is_np = isinstance(coordinates, np.ndarray)
    c = torch.from_numpy(np.ascontiguousarray(coordinates)) if is_np else coordinates
    assert c.ndim == 2, "coordinates must be [N, D]"
    home = c.device
    dev = home
    if quantization_size is not None:
        c = c.double() / torch.as_tensor(quantization_size, dtype=torch.float64, device=dev)
    if c.dtype.is_floating_point:
        c = torch.floor(c)
    disc = c.to(torch.int32).contiguous()
    if home.type == "cuda":
        rows = torch.cat([torch.zeros((disc.shape[0], 1), dtype=torch.int32, device=dev), disc], 1)
        cmap, first, inverse = unique_coords(rows, (1, 1, 1))
        unique_rows = cmap.coords[:, 1:].contiguous()
    else:  # host path: no CUDA context (fork-safe)
        n, d = disc.shape
        first = torch.empty(max(n, 1), dtype=torch.int32)
        inverse = torch.empty(max(n, 1), dtype=torch.int32)
        m = lib.us3d_coords_unique_h(disc.data_ptr(), n, d, first.data_ptr(), inverse.data_ptr())
        if m < 0:
            check(m)
        first, inverse = first[:m], inverse[:n]
        unique_rows = disc[first.long()]
    first, inverse = first.long(), inverse.long()

    def back(t):
        t = t.to(home)
        return t.numpy() if is_np else t

    out_labels = None
    if labels is not None:
        lab = torch.as_tensor(labels).to(dev)
        out_labels = lab[first].clone()
        mism = lab != out_labels[inverse]
        bad = torch.zeros(first.shape[0], dtype=torch.bool, device=dev)
        bad[inverse[mism]] = True
        out_labels[bad] = ignore_label
    if return_maps_only:
        return (back(first), back(inverse)) if return_inverse else back(first)
    ret = [back(unique_rows)]
    if features is not None:
        if isinstance(features, torch.Tensor):
            ret.append(features[first.to(features.device)])
        else:
            ret.append(np.asarray(features)[first.cpu().numpy()])
    if labels is not None:
        ret.append(back(out_labels))
    if return_index:
        ret.append(back(first))
    if return_inverse:
        ret.append(back(inverse))
    return ret[0] if len(ret) == 1 else tuple(ret)


def batched_coordinates(coords, dtype=torch.int32, device=None):
    out = []
    for b, c in enumerate(coords):
        c = torch.as_tensor(c)
        if c.dtype.is_floating_point:
            c = torch.floor(c)
        out.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=dtype, device=c.device), c.to(dtype)], 1))
    res = torch.cat(out, 0) if out else torch.zeros((0, 4), dtype=dtype)
    return res.to(device) if device is not None else res


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    bcoords = batched_coordinates(coords, dtype=dtype, device=device)
    f = torch.cat([torch.as_tensor(x) for x in feats], 0)
    if labels is None:
        return bcoords, f
    return bcoords, f, torch.cat([torch.as_tensor(x) for x in labels], 0)
