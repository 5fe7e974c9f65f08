"""Voxelise + collate + target building on the device — the training branch of the reference's
datasets/utils.py::freemask_voxelize (:370-477) and get_instance_freemasks (:480-527).

The reference runs this in 4 DataLoader worker processes on the CPU (numpy floor, ME.utils.sparse_quantize's hash insert over
10^5..10^6 points per scene, per-instance Python loops with one N-long boolean mask and one `unique` each) and ships un-pinned
tensors to the GPU afterwards (conf/data/indoor.yaml:23-25).  Here the raw per-scene arrays are staged in pinned host memory and
copied asynchronously on a side stream; everything else — quantisation (the coordinate manager's hash kernels,
us3d_coords_unique), gathers, padding, the batch column, segment compaction and the [T, N] / [T, S] target masks — happens on the
device without a host round trip except the two sizes torch needs to allocate results (unique voxels, unique segments).

Results are element-for-element those of the reference function (tests/test_collate.py runs the UNMODIFIED reference function over
the shim's host path next to this one), including its quirk in the segment masks: `segment_mask[rows.unique()] = True` (:505) takes
the unique values over ALL columns of the selected rows — label column and the 0/1 mask columns included — so segments 0 and 1
(and the label values) are switched on for every instance.  `keep_reference_quirks=False` uses the segment column only.
"""
from typing import List, Optional, Sequence

import numpy as np
import torch

import MinkowskiEngine as ME  # the shim (unscene3d_b200/shims)


def _stage(array, device, stream):
    """Host array -> pinned buffer -> device, asynchronously on `stream`."""
    t = torch.from_numpy(np.ascontiguousarray(array))
    if device.type != "cuda":
        return t
    t = t.pin_memory()
    with torch.cuda.stream(stream):
        out = t.to(device, non_blocking=True)
    out.record_stream(torch.cuda.current_stream(device))
    return out


def instance_freemasks(labels: torch.Tensor, segments: Optional[torch.Tensor] = None, keep_reference_quirks: bool = True):
    """get_instance_freemasks (:480-527) for ONE scene: labels [N, 2 + M] (label, M mask columns, segment index) ->
    dict(labels [T], masks [T, N], segment_mask [T, S]) over the instances with at least one point, or None if there is none
    (the reference then returns an empty target list for the whole batch)."""
    n_inst = labels.shape[1] - 2
    cols = labels[:, 1:1 + n_inst].bool()
    keep = cols.any(0)
    if not bool(keep.any()):
        return None
    masks = cols[:, keep].T.contiguous()
    out = {"labels": torch.ones(masks.shape[0], dtype=torch.int64, device=labels.device), "masks": masks}
    if segments is not None:
        S = segments.shape[0]
        t_idx, r_idx = torch.nonzero(masks, as_tuple=True)
        vals = labels[r_idx] if keep_reference_quirks else labels[r_idx][:, -1:]
        if vals.numel() and (int(vals.max()) >= S or int(vals.min()) < -S):
            raise IndexError(f"index {int(vals.max())} is out of bounds for dimension 0 with size {S}")  # as the reference's indexing does
        seg_mask = torch.zeros((masks.shape[0], S), dtype=torch.bool, device=labels.device)
        seg_mask[t_idx[:, None].expand_as(vals), vals] = True
        out["segment_mask"] = seg_mask
    return out


def freemask_voxelize_device(batch: Sequence, voxel_size: float, device, stream: Optional["torch.cuda.Stream"] = None,
                             keep_reference_quirks: bool = True):
    """batch: list of samples (coordinates float [P, 3], features float [P, C], freemasks int [P, 2 + M] = (label, masks..., segment
    id), ...) as the dataset yields them (datasets/freemask_semseg.py).  Returns dict(coordinates int32 [sum N, 4], features float32
    [sum N, C], inverse_maps list of int64 [P], unique_maps, target list, target_full list) on `device`."""
    device = torch.device(device)
    if device.type == "cuda" and stream is None:
        stream = torch.cuda.Stream(device=device)
    coords_l, feats_l, masks_l, inverse_l, unique_l, full_l = [], [], [], [], [], []
    for sample in batch:
        xyz = _stage(sample[0], device, stream)
        feats = _stage(sample[1], device, stream)
        fm = _stage(sample[2], device, stream)
        if device.type == "cuda":
            torch.cuda.current_stream(device).wait_stream(stream)
        c = torch.floor(xyz.double() / voxel_size)  # np.floor(sample[0] / voxel_size), the same IEEE operations
        umap, imap = ME.utils.sparse_quantize(coordinates=c, return_index=True, return_inverse=True, return_maps_only=True)
        unique_l.append(umap)
        inverse_l.append(imap)
        coords_l.append(c[umap].int())
        feats_l.append(feats[umap].float())
        masks_l.append(fm[umap].long())
        full_l.append(fm.long())
    width = max(f.shape[1] for f in masks_l)
    padded = []
    for f in masks_l:  # pad the mask columns to the widest scene, segment ids stay last (:418-421)
        pad = torch.zeros((f.shape[0], width - f.shape[1]), dtype=f.dtype, device=f.device)
        padded.append(torch.cat([f[:, :-1], pad, f[:, -1:]], 1))
    coordinates, features = ME.utils.sparse_collate(coords_l, feats_l)
    target, target_full = [], []
    for lab in padded:
        uniq, inv = torch.unique(lab[:, -1], return_inverse=True)  # np.unique over the segment ids (:440)
        first = torch.full((uniq.shape[0],), lab.shape[0], dtype=torch.int64, device=lab.device)
        first.scatter_reduce_(0, inv, torch.arange(lab.shape[0], device=lab.device), "amin")
        lab = lab.clone()
        lab[:, -1] = inv
        segment2label = lab[first][:, :-1]
        t = instance_freemasks(lab, segment2label, keep_reference_quirks)
        if t is None:
            target = []
            break
        t["point2segment"] = lab[:, -1]
        target.append(t)
    if target:
        for fm in full_l:
            t = instance_freemasks(fm, None)
            if t is None:
                target_full = []
                break
            t["point2segment"] = fm[:, -1]
            target_full.append(t)
    return {"coordinates": coordinates, "features": features, "inverse_maps": inverse_l, "unique_maps": unique_l, "target": target,
            "target_full": target_full}
