"""Drop-in `torch_scatter` for the three functions the reference imports (models/mask3d.py:12, 65-67, 223;
trainer/trainer.py:9, 449).  scatter_mean over rows runs on the libus3d segment kernels."""
import torch

from unscene3d_b200.engine import functional as _Fn


def _dim0_2d(src, index, dim):
    if dim not in (0, -src.ndim) or src.ndim != 2 or index.ndim != 1:
        raise NotImplementedError("unscene3d_b200.torch_scatter: only row scatter (dim=0) of [N, C] by [N] is on the hot path")


def _n_segments(index, dim_size):
    if dim_size is None:
        return int(index.max()) + 1 if index.numel() else 0
    n_seg = int(dim_size)
    if index.numel() and (int(index.min()) < 0 or int(index.max()) >= n_seg):  # torch_scatter raises here as well
        raise RuntimeError(f"scatter: index out of range for dim_size {n_seg}")
    return n_seg


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    _dim0_2d(src, index, dim)
    if out is not None:
        raise NotImplementedError("scatter_mean(out=...) is not used by the reference")
    n_seg = _n_segments(index, dim_size)
    if not src.is_cuda:
        # torch_scatter's CPU path: the reference's evaluation post-processing scatters CPU tensors
        # (trainer/trainer.py:449 `scatter_mean(mask.detach().cpu(), point2segment_full, dim=0)`); not on the training path
        idx = index.to(src.device).long()
        out_ = torch.zeros((n_seg, src.shape[1]), dtype=src.dtype, device=src.device).index_add_(0, idx, src)
        cnt = torch.bincount(idx, minlength=n_seg).clamp_(min=1).to(out_.dtype if out_.dtype.is_floating_point else torch.float32)
        return out_ / cnt[:, None] if out_.dtype.is_floating_point else torch.div(out_, cnt[:, None].to(out_.dtype), rounding_mode="floor")
    return _Fn.SegmentMeanFunction.apply(src, index, n_seg)


def _scatter_minmax(src, index, dim, dim_size, reduce):
    _dim0_2d(src, index, dim)
    n_seg = _n_segments(index, dim_size)
    index = index.to(src.device)
    idx = index.long()[:, None].expand_as(src)
    out = torch.zeros((n_seg, src.shape[1]), dtype=src.dtype, device=src.device)
    out = out.scatter_reduce(0, idx, src, reduce=reduce, include_self=False)
    hit = src == out[index.long()]
    rows = torch.arange(src.shape[0], device=src.device)[:, None].expand_as(src)
    arg = torch.full((n_seg, src.shape[1]), src.shape[0], dtype=torch.long, device=src.device)
    arg = arg.scatter_reduce(0, idx, torch.where(hit, rows, torch.full_like(rows, src.shape[0])), reduce="amin", include_self=True)
    return out, arg


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    return _scatter_minmax(src, index, dim, dim_size, "amax")


def scatter_min(src, index, dim=-1, out=None, dim_size=None):
    return _scatter_minmax(src, index, dim, dim_size, "amin")
