"""CPU oracle of the FreeMask-style pseudo-mask variant — TEST INFRASTRUCTURE ONLY
(imported by tests/ and the golden generator, never by unscene3d_b200).

Restates, statement by statement, the segment branch of the scene loop in the reference's
pseudo_masks/freemask_main.py (the code is the body of `main()`, not a function):

    :203-221  per-segment mean of the valid (non-zero) point features, all-zero segments dropped
    :226-232  directed connectivity dictionary
    :242      soft masks = cosine_sim(keys, queries)                      (utils/freemask_utils.py:8-18)
    :266-279  zero-feature columns, hard threshold, candidates with > 2 segments
    :282-349  separation of non-connected blobs activated by the same query (incremental merging, including the
              reference's index skip after `pop`), candidates with > 3 segments
    :353-356  maskness, descending sort
    :359-372  segment masks mapped onto the low-resolution points
    :375-396  XY-extent filter
    :398-417  matrix_nms(kernel='mask') (utils/pc_utils.py:724-757), top max_instance_num, maskness threshold

Pinned by tests/golden/freemask_scene.npz, which tests/golden/make_freemask_golden.py produces by executing the reference's
own source lines (taken from the file untouched) on the same inputs.
"""
from types import SimpleNamespace

import numpy as np
import torch

DEFAULTS = SimpleNamespace(hard_mask_threshold=0.35, nms_maskness_threshold=0.6, instance_to_scene_max_ratio=0.8,
                           max_instance_num=50)  # pseudo_masks/config/default.yaml:57-63


def cosine_sim(feats_k, feats_q):
    """utils/freemask_utils.py:8-18."""
    eps = 10e-10
    key_feats = feats_k / (feats_k.norm(dim=1, keepdim=True) + eps)
    queries = feats_q / (feats_q.norm(dim=1, keepdim=True) + eps)
    attn = queries @ key_feats.T
    attn -= attn.min(-1, keepdim=True)[0]
    attn /= attn.max(-1, keepdim=True)[0] + eps
    return attn


def matrix_nms_mask(cate_labels, seg_masks, sum_masks, cate_scores, nms_thr=0.5):
    """utils/pc_utils.py:724-757, kernel == 'mask'."""
    n_samples = len(cate_scores)
    if n_samples == 0:
        return []
    keep = seg_masks.new_ones(cate_scores.shape)
    seg_masks = seg_masks.float()
    for i in range(n_samples - 1):
        if not keep[i]:
            continue
        for j in range(i + 1, n_samples):
            if not keep[j]:
                continue
            if cate_labels[i] != cate_labels[j]:
                continue
            inter = (seg_masks[i] * seg_masks[j]).sum()
            union = sum_masks[i] + sum_masks[j] - inter
            if union > 0:
                if inter / union > nms_thr:
                    keep[j] = False
            else:
                keep[j] = False
    cate_scores[~keep] = 0.0
    return cate_scores


def segment_features(keys_F, matching_segment_ids):
    """:203-222 -> (segment_feats of the valid segments, their ids)."""
    unique_segments = matching_segment_ids.unique()
    segment_feats = torch.zeros((len(unique_segments), keys_F.shape[1]))
    valid_mask = torch.any(keys_F != 0, dim=-1)
    for i, s_id in enumerate(unique_segments):
        segment_mask = valid_mask * (matching_segment_ids == s_id)
        if segment_mask.sum() > 0:
            segment_feats[i, :] = keys_F[segment_mask].mean(0)
    valid_segments = torch.any(segment_feats != 0, dim=-1)
    return segment_feats[valid_segments], unique_segments[valid_segments]


def separate_blobs(masks, unique_segments, connectivity_dict):
    """:289-326 — per query, the list of blobs (sets of segment ids) in the order the reference builds them."""
    all_fused_instances = []
    for m in masks:
        curr_instances = []
        for c in unique_segments[m].cpu().numpy():
            neighbour_segments = connectivity_dict[c.item()]
            last_fused_match = -1
            merged = False
            fused_id = 0
            while fused_id < len(curr_instances):
                fused_segments = curr_instances[fused_id]
                if len(neighbour_segments.intersection(fused_segments)) != 0:
                    merged = True
                    fused_segments.add(c)
                    if last_fused_match != -1:
                        curr_instances[last_fused_match] = curr_instances[last_fused_match].union(fused_segments)
                        curr_instances.pop(fused_id)
                    else:
                        last_fused_match = fused_id
                fused_id += 1  # also after a pop: the blob that slid into this slot is skipped (reference behaviour)
            if not merged:
                curr_instances += [set([c])]
        all_fused_instances += [curr_instances]
    return all_fused_instances


def freemask(keys_F, matching_segment_ids, seg_connectivity, lr_coords, coords, cfg=DEFAULTS, trace=None):
    """keys_F [N, C] float32 low-resolution point features, matching_segment_ids [N] int64, seg_connectivity [E, 2] int64
    (directed), lr_coords [N, 3] numpy, coords [N0, 4] (batch, xyz) of the full-resolution cloud.
    Returns (soft_masks [M, N] float32 on the low-resolution points, maskness [M]) or None when the reference `continue`s."""
    thr = cfg.hard_mask_threshold
    segment_feats, unique_segments = segment_features(keys_F, matching_segment_ids)
    key_feats = segment_feats.clone().detach()
    queries = segment_feats.clone().detach()
    connectivity_dict = {}
    for s_id in unique_segments:
        connectivity_dict[s_id.item()] = set(seg_connectivity[seg_connectivity[:, 0] == s_id, 1].cpu().numpy())

    soft_masks = cosine_sim(key_feats, queries)
    soft_masks[:, torch.all(key_feats == 0, dim=-1)] = 0.0
    masks = soft_masks >= thr
    sum_masks = masks.sum(1)
    keep = sum_masks > 2
    if keep.sum() == 0:
        return None
    masks, soft_masks, sum_masks = masks[keep], soft_masks[keep], sum_masks[keep]
    if trace is not None:
        trace["soft_segments"] = soft_masks.clone()

    masks = (soft_masks >= thr).bool()
    all_fused_instances = separate_blobs(masks, unique_segments, connectivity_dict)
    all_fused_instance_num = sum(len(q) for q in all_fused_instances)
    separated_id = 0
    separated_soft_masks = torch.zeros((all_fused_instance_num, soft_masks.shape[1]))
    for query_id, separated_query in enumerate(all_fused_instances):
        for separated_segment in separated_query:
            for segment_id_in_inst in separated_segment:
                segment_location = torch.nonzero(unique_segments == segment_id_in_inst)[0][0]
                separated_soft_masks[separated_id, segment_location] = soft_masks[query_id, segment_location]
            separated_id += 1
    soft_masks = separated_soft_masks
    masks = (soft_masks >= thr).bool()
    sum_masks = masks.sum(1)
    keep = sum_masks > 3
    masks, soft_masks, sum_masks = masks[keep], soft_masks[keep], sum_masks[keep]
    if trace is not None:
        trace["separated_segments"] = soft_masks.clone()

    maskness = (soft_masks * masks.float()).sum(1) / sum_masks
    sort_inds = torch.argsort(maskness, descending=True)
    maskness = maskness[sort_inds]
    soft_masks = soft_masks[sort_inds]

    mapped_soft_masks = torch.zeros((soft_masks.shape[0], lr_coords.shape[0]))
    for i, s_id in enumerate(unique_segments):
        segment_mask = matching_segment_ids == s_id
        mapped_soft_masks[:, segment_mask] = soft_masks[:, i].view(-1, 1)
    soft_masks = mapped_soft_masks.clone()
    masks = (soft_masks >= thr).bool()

    filtered_masks = []
    scene_extents = (coords[:, 1:].max(0)[0] - coords[:, 1:].min(0)[0]).cpu().numpy()
    for mask_id in range(len(masks)):
        if torch.all(~masks[mask_id]):
            continue
        inst_coords = lr_coords[masks[mask_id].cpu().numpy()]
        inst_extent = inst_coords.max(0) - inst_coords.min(0)
        extent_ratios = inst_extent / scene_extents
        if np.any(extent_ratios[:2] > cfg.instance_to_scene_max_ratio):
            continue
        filtered_masks += [mask_id]
    if filtered_masks != []:
        masks = masks[filtered_masks]
        soft_masks = soft_masks[filtered_masks]
        maskness = maskness[filtered_masks]
        sum_masks = masks.sum(1)

    maskness = matrix_nms_mask(maskness * 0, masks, sum_masks, maskness)
    sort_inds = torch.argsort(maskness, descending=True)
    if len(sort_inds) > cfg.max_instance_num:
        sort_inds = sort_inds[:cfg.max_instance_num]
    maskness = maskness[sort_inds].cpu()
    soft_masks = soft_masks[sort_inds].cpu()
    keep = maskness > cfg.nms_maskness_threshold
    if keep.sum() == 0:
        return None
    return soft_masks[keep], maskness[keep]
