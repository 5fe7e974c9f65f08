"""CPU restatement of the pseudo-mask NCut path — TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows /root/reference/pseudo_masks/unscene3d_pseudo_main.py:
    normalize_mat :82-86, get_affinity_matrix :89-119 (two-modality branch), get_masked_affinity_matrix :122-135,
    second_smallest_eigenvector :138-146, get_salient_areas :149-153, separate_segments :181-250 (mode 'max'),
    segment_ids_to_mask :254-260, aggregate_features :350-402, unscene3d :405-502.
Pinned against the reference functions themselves (imported unmodified behind import stubs) in
tests/test_ncut.py::test_oracle_matches_reference_functions, which only runs where /root/reference exists.

`sign_hook(v) -> +1/-1` lets a caller impose the eigenvector sign: LAPACK's sign is arbitrary, yet the reference's
foreground choice (v > mean(v), :437) depends on it unless more than 80 % of the segments land in the foreground.
"""
import numpy as np
import torch
import torch.nn.functional as F
from scipy.linalg import eigh


def normalize_mat(A, eps=1e-5):
    A = A.copy()
    A -= np.min(A[np.nonzero(A)]) if np.any(A > 0) else 0
    A[A < 0] = 0.0
    A /= A.max() + eps
    return A


def affinity(feats_a: torch.Tensor, feats_b: torch.Tensor, tau: float, eps: float = 1e-5):
    """Thresholded two-modality affinity: returns (A float64 {eps, 1}, D diag float64)."""
    fa, fb = F.normalize(feats_a, p=2, dim=-1), F.normalize(feats_b, p=2, dim=-1)
    A_a, A_b = (fa @ fa.T).cpu().numpy(), (fb @ fb.T).cpu().numpy()
    A = (normalize_mat(A_a) + normalize_mat(A_b)) / 2
    A = A > tau
    A = np.where(A.astype(float) == 0, eps, A)
    return A, np.diag(np.sum(A, axis=0))


def cosine_sim(feats_k: torch.Tensor, feats_q: torch.Tensor):
    """utils/freemask_utils.py:8-18 — NOT symmetric: every query row has its own minimum subtracted and its own maximum divided out."""
    eps = 10e-10
    key_feats = feats_k / (feats_k.norm(dim=1, keepdim=True) + eps)
    queries = feats_q / (feats_q.norm(dim=1, keepdim=True) + eps)
    attn = queries @ key_feats.T
    attn = attn - attn.min(-1, keepdim=True)[0]
    attn = attn / (attn.max(-1, keepdim=True)[0] + eps)
    return attn


def affinity_single(feats: torch.Tensor, tau: float, eps: float = 1e-5):
    """Single-modality branch of get_affinity_matrix (:92-98): row-normalised cosine_sim of the L2-normalised features, then
    normalize_mat, threshold; D = COLUMN sums of the (asymmetric) thresholded matrix.  scipy's eigh(D - A, D) reads the lower
    triangle of A only."""
    f = F.normalize(feats, p=2, dim=-1)
    A = normalize_mat(cosine_sim(f, f).cpu().numpy())
    A = A > tau
    A = np.where(A.astype(float) == 0, eps, A)
    return A, np.diag(np.sum(A, axis=0))


def fiedler(A, D):
    _, vecs = eigh(D - A, D, subset_by_index=[1, 2])
    return vecs[:, 0].copy()


def connected_component_of(seed_id, fg_ids, connectivity):
    """Union of fg segments connected (through fg segments only) to seed_id; connectivity: dict id -> set(ids)."""
    fg = set(int(s) for s in fg_ids)
    comp, frontier = {int(seed_id)}, [int(seed_id)]
    while frontier:
        cur = frontier.pop()
        for nb in connectivity.get(cur, ()):  # directed rows [a, b] as stored in seg_connectivity
            if nb in fg and nb not in comp:
                comp.add(nb)
                frontier.append(nb)
    return comp


def separate_segments_max(bipartition, vec, unique_segments, seg_connectivity):
    return separate_segments(bipartition, vec, unique_segments, seg_connectivity, "max")


def separate_segments(bipartition, vec, unique_segments, seg_connectivity, mode="max"):
    """:181-250, restated literally: foreground segments are visited in id order; a segment joins every
    existing blob that contains one of ITS listed neighbours (rows [c, *] of seg_connectivity, as stored — directed),
    blobs bridged by it are merged (the scan index still advances after a merge, as in the reference), otherwise
    it opens a new blob.  'max': the blob containing the segment of argmax(vec); 'avg': the blob with the highest mean of vec;
    'largest': the blob with most segments (first one on ties, np.argmax); 'all': every foreground segment."""
    ids = unique_segments.cpu().numpy()
    conn = seg_connectivity.cpu().numpy()
    nbrs = {int(s): set(conn[conn[:, 0] == s, 1].tolist()) for s in ids}
    blobs = []
    for c in ids[bipartition]:
        c = int(c)
        first, merged, pos = -1, False, 0
        while pos < len(blobs):
            blob = blobs[pos]
            if nbrs[c] & blob:
                merged = True
                blob.add(c)
                if first != -1:
                    blobs[first] = blobs[first] | blob
                    blobs.pop(pos)
                else:
                    first = pos
            pos += 1
        if not merged:
            blobs.append({c})
    if mode == "max":
        seed_id = int(ids[int(np.argmax(vec))])
        return next(b for b in blobs if seed_id in b)
    if mode == "avg":
        means = [np.mean(vec[np.isin(ids, list(b))]) for b in blobs]
        return blobs[int(np.argmax(means))]
    if mode == "largest":
        return blobs[int(np.argmax(np.array([len(b) for b in blobs])))]
    if mode == "all":
        return set(int(c) for c in ids[bipartition])
    raise NotImplementedError(mode)


def aggregate_features(encoded, segment_ids, seg_connectivity, mode="mean"):
    unique_segments = segment_ids.unique()
    seg = torch.zeros((len(unique_segments), encoded.shape[1]))
    valid = torch.any(encoded != 0, dim=-1)
    for i, s_id in enumerate(unique_segments):
        m = valid * (segment_ids == s_id)
        if m.sum() > 0:
            rows = encoded[m]
            seg[i] = rows.max(0)[0] if mode == "max" else rows.mean(0)
    agg = seg.clone()
    zero_segments = unique_segments[torch.all(agg == 0, dim=-1)]
    for z in zero_segments:
        idx = (unique_segments == z).nonzero(as_tuple=True)[0]
        # reference quirk (:387): the neighbours of the FIRST zero segment are used for every zero segment
        nb = seg_connectivity[seg_connectivity[:, 0] == zero_segments[0]][:, 1]
        nb_idx = torch.LongTensor([int((unique_segments == s).nonzero(as_tuple=True)[0]) for s in nb])
        nb_feats = agg[nb_idx]
        nb_feats = nb_feats[torch.any(nb_feats != 0.0, dim=-1)]
        agg[idx] = nb_feats.mean(0) if len(nb_feats) else agg.mean(0)
    return agg, unique_segments


def unscene3d(feats_a, feats_b, unique_segments, seg_connectivity, affinity_tau=0.65, max_number_of_instances=20,
              max_extent_ratio=0.8, eps=1e-5, min_segment_size=4, sign_hook=None, trace=None, separation_mode="max"):
    """Greedy NCut mask extraction over segments; returns bool [n_masks, S].  feats_b = None: single-modality affinity."""
    S = len(unique_segments)
    if S < 3:
        return np.ones((1, S), dtype=bool)
    ids = unique_segments.cpu().numpy()
    masks, foreground = [], set()
    painting = torch.zeros(S)
    current = None
    fa, fb = feats_a.clone(), (None if feats_b is None else feats_b.clone())
    for it in range(max_number_of_instances):
        if it > 0:
            painting = ((painting.view(S, 1) + current.view(S, 1).float()) > 0).float()
            fa, fb = (1 - painting) * fa, (None if fb is None else (1 - painting) * fb)
            painting = painting.squeeze()
        A, D = affinity(fa, fb, affinity_tau, eps) if fb is not None else affinity_single(fa, affinity_tau, eps)
        pm = painting.bool().numpy()
        A[pm] = eps
        A[:, pm] = eps
        vec = fiedler(A, D)
        if sign_hook is not None:
            vec = vec * sign_hook(vec)
        if trace is not None:
            trace.append(vec.copy())
        bip = vec > vec.sum() / len(vec)
        if bip.sum() / len(bip) > max_extent_ratio:
            bip, vec = np.logical_not(bip), -vec
        part = (separate_segments_max(bip, vec, unique_segments, seg_connectivity) if separation_mode == "max"
                else separate_segments(bip, vec, unique_segments, seg_connectivity, separation_mode))
        current = torch.from_numpy(np.isin(ids, list(part)))
        iou = len(part & foreground) / len(part)
        if iou > 0.5 or len(part) < min_segment_size:
            continue
        masks.append(np.isin(ids, list(part - foreground)))
        foreground |= part
    return np.stack(masks) if masks else np.zeros((0, S), dtype=bool)
