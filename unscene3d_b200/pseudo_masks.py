"""Pseudo-mask generation by iterative Normalized Cut over segments — host-side mirror of the reference's
pseudo_masks/unscene3d_pseudo_main.py (aggregate_features :350-402, get_affinity_matrix :89-119,
second_smallest_eigenvector :138-146, separate_segments :181-250, unscene3d :405-502).

What moves to the device (libus3d, csrc/ncut.cu + decoder_ops.cu):
  * per-segment feature means                       segment-mean kernels (the reference loops over S segments with an
                                                    N-long boolean mask each)
  * affinity: row-normalise, Gram matrix per modality with the global statistics of normalize_mat fused in,
    threshold the averaged matrices into a BIT matrix + degrees (the reference copies two [S, S] fp32 matrices to
    the host and finishes in float64 numpy)
  * spectral step: Lanczos with full re-orthogonalisation in fp64 over M = D^-1/2 W D^-1/2, W x computed from the
    bit matrix (the reference: dense LAPACK `eigh` on the host, O(S^3), x <= 20 per scene)
Left on the host, as in the reference: the set logic on <= a few thousand segment ids (blob growing, IoU against
previous foregrounds) and the m x m tridiagonal eigen-solve of the Lanczos recurrence (m <= 600).

The eigenvector's sign is arbitrary in LAPACK, yet the reference's foreground (v > mean(v)) depends on it unless one
side holds > 80 % of the segments.  `sign_rule="minority"` (default) orients v so that the foreground is the
smaller side — identical to the reference whenever its own flip rule fires or the small side is below 20 %;
pass a callable to impose another convention (the parity tests pass the oracle's sign).
"""
from __future__ import annotations

import ctypes
from typing import Callable, Optional, Union

import numpy as np
import torch

from ._lib import check, lib
from .engine import functional as Fn
from .engine.coords import _stream


class NCutGraph:
    """Thresholded affinity W = eps 11^T + (1 - eps) B as a bit matrix, plus the degrees of the unpainted graph."""

    def __init__(self, bits: torch.Tensor, degree: torch.Tensor, n: int, eps: float, trivial_known: bool = True):
        self.bits, self.degree, self.n, self.eps = bits, degree, n, eps
        # D = row sums of W: D^1/2 1 is the leading eigenvector of D^-1/2 W D^-1/2 and can be deflated.  False for the
        # single-modality graph (degrees of the asymmetric thresholded matrix): the solver then takes the second Ritz pair.
        self.trivial_known = trivial_known

    def dense(self) -> torch.Tensor:
        """float64 [S, S] (tests / debugging)."""
        words = self.bits.shape[1]
        shifts = torch.arange(32, device=self.bits.device, dtype=torch.int64)
        b = ((self.bits.to(torch.int64)[:, :, None] & 0xFFFFFFFF) >> shifts) & 1
        b = b.reshape(self.n, words * 32)[:, : self.n].double()
        return self.eps + (1.0 - self.eps) * b


def aggregate_features(encoded_features: torch.Tensor, segment_ids: torch.Tensor, seg_connectivity: torch.Tensor, aggregation_mode: str = "mean"):
    """Per-segment mean of the rows that are not all-zero; segments without a valid row take the mean of the
    non-zero neighbours listed for the FIRST such segment (reference quirk, :387), else the global mean."""
    if aggregation_mode not in ("mean", "max"):
        raise ValueError(f"aggregation_mode {aggregation_mode!r}: the reference knows 'mean' and 'max' (:366)")
    unique_segments, index = torch.unique(segment_ids, return_inverse=True)
    valid = torch.any(encoded_features != 0, dim=-1)
    S = unique_segments.shape[0]
    src, idx = encoded_features[valid].float().contiguous(), index[valid].contiguous().long()
    if aggregation_mode == "max":
        # per-segment maximum over the valid rows (order-independent, exact); segments without a valid row stay zero
        agg = torch.full((S, src.shape[1]), float("-inf"), dtype=torch.float32, device=src.device)
        agg.scatter_reduce_(0, idx[:, None].expand(-1, src.shape[1]), src, "amax", include_self=True)
        agg[torch.isinf(agg)] = 0.0
    else:
        agg = torch.empty((S, src.shape[1]), dtype=torch.float32, device=src.device)
        acc = torch.zeros((S, src.shape[1]), dtype=torch.float64, device=src.device)
        count = torch.zeros(S, dtype=torch.float32, device=src.device)
        # fp64 sums, rounded once: the affinity threshold downstream must not see the order of the atomics
        check(lib.us3d_segment_mean_f64(src.data_ptr(), idx.data_ptr(), src.shape[0], src.shape[1], S, acc.data_ptr(), agg.data_ptr(),
                                        count.data_ptr(), _stream()))
    zero = torch.all(agg == 0, dim=-1)
    if bool(zero.any()):
        first_zero = unique_segments[zero][0]
        nb = seg_connectivity[seg_connectivity[:, 0] == first_zero][:, 1]
        nb_idx = torch.searchsorted(unique_segments, nb)
        # the reference fills the zero segments one after another from the running matrix (earlier fills are visible)
        for i in torch.nonzero(zero).flatten().tolist():
            nb_feats = agg[nb_idx]
            nb_feats = nb_feats[torch.any(nb_feats != 0.0, dim=-1)]
            agg[i] = nb_feats.mean(0) if nb_feats.shape[0] else agg.mean(0)
    return agg, unique_segments


def get_affinity_matrix(feats, tau: float = 0.15, eps: float = 1e-5, painted: Optional[torch.Tensor] = None) -> NCutGraph:
    """Affinity -> thresholded bit graph.  feats = (feats_a [S, Da], feats_b [S, Db]) fp32 CUDA: the configured two-modality
    path (fused kernels); a single tensor: the reference's single-modality branch."""
    if not isinstance(feats, tuple):
        return _affinity_single(feats, tau, eps, painted)
    fa, fb = [f.float().contiguous() for f in feats]
    S = fa.shape[0]
    dev = fa.device
    st = _stream()
    mats, stats = [], []
    for f in (fa, fb):
        A = torch.empty((S, S), dtype=torch.float32, device=dev)
        inv = torch.empty(S, dtype=torch.float32, device=dev)
        s3 = torch.empty(4, dtype=torch.int32, device=dev)
        check(lib.us3d_ncut_gram(f.data_ptr(), S, f.shape[1], inv.data_ptr(), A.data_ptr(), s3.data_ptr(), st))
        mats.append(A)
        stats.append(s3)
    words = (S + 31) // 32
    bits = torch.empty((S, words), dtype=torch.int32, device=dev)
    degree = torch.empty(S, dtype=torch.float64, device=dev)
    pm = None if painted is None else painted.to(torch.uint8).contiguous()
    check(lib.us3d_ncut_threshold(mats[0].data_ptr(), mats[1].data_ptr(), S, stats[0].data_ptr(), stats[1].data_ptr(), float(tau),
                                  float(eps), 0 if pm is None else pm.data_ptr(), bits.data_ptr(), degree.data_ptr(), st))
    return NCutGraph(bits, degree, S, eps)


def _affinity_single(feats: torch.Tensor, tau: float, eps: float, painted: Optional[torch.Tensor]) -> NCutGraph:
    """Single-modality branch of get_affinity_matrix (:92-98; off the configured path, device tensor ops): the row-normalised
    cosine_sim (utils/freemask_utils.py:8-18) of the L2-normalised features is NOT symmetric; the reference thresholds it as
    it is, takes D from the COLUMN sums, and scipy's eigh(D - A, D) then reads the lower triangle of A only — so the graph is
    the lower triangle mirrored, with the degrees of the full asymmetric matrix."""
    f = torch.nn.functional.normalize(feats.float(), p=2, dim=-1)
    k = f / (f.norm(dim=1, keepdim=True) + 10e-10)
    A = k @ k.T
    A = A - A.min(-1, keepdim=True)[0]
    A = A / (A.max(-1, keepdim=True)[0] + 10e-10)
    if bool((A > 0).any()):
        A = A - A[A != 0].min()
    A = A.clamp_min(0.0)
    A = A / (A.max() + 1e-5)
    on = A > tau
    S = on.shape[0]
    degree = torch.where(on, 1.0, float(eps)).double().sum(0)
    low = torch.tril(on)
    sym = low | low.T
    if painted is not None:
        pb = painted.bool()
        sym = sym & ~pb[:, None] & ~pb[None, :]
    words = (S + 31) // 32
    pad = torch.zeros((S, words * 32), dtype=torch.int64, device=on.device)
    pad[:, :S] = sym
    bits = (pad.view(S, words, 32) << torch.arange(32, device=on.device)).sum(-1)
    bits = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32).contiguous()
    return NCutGraph(bits, degree.contiguous(), S, eps, trivial_known=False)


def _lanczos_top_deflated(matvec, u1: torch.Tensor, max_steps: int, tol: float, seed: int, check_every: int = 20,
                          min_steps: int = 512, breakdown: float = 1e-10, info: Optional[dict] = None, pick: int = -1) -> torch.Tensor:
    """Unit eigenvector of the LARGEST eigenvalue of the symmetric operator `matvec` restricted to the complement of the
    known unit eigenvector `u1` (fp64, full re-orthogonalisation twice per step, u1 included in the basis).

    Why this shape.  The leading eigenvalues of M = D^-1/2 W D^-1/2 of a clustered graph sit within 1e-5 of each other
    and single-vector Lanczos "misconverges" there (a Ritz pair stalls on the third eigenvalue with a tiny residual
    before moving on), so no early exit is taken before `min_steps` — with full re-orthogonalisation the recurrence run
    to the dimension of the Krylov space is an exact tridiagonalisation, as reliable as the reference's dense LAPACK
    solve, and S is 1-3 k segments.  Once most segments are painted the Krylov space is tiny (a handful of distinct
    eigenvalues): the recurrence breaks down (beta -> 0) after a few steps and whatever is normalised after that is
    rounding noise that re-discovers the trivial eigenvalue 1 as a "ghost"; deflating the known trivial vector
    u1 = D^1/2 1 / |.| and cutting the recurrence at the first beta < `breakdown` removes both failure modes."""
    S, dev = u1.shape[0], u1.device
    m = min(max_steps, S - 1)
    Q = torch.zeros((m + 2, S), dtype=torch.float64, device=dev)
    Q[0] = u1  # row 0 is the deflated vector: part of every re-orthogonalisation, not of the recurrence
    g = torch.Generator(device="cpu").manual_seed(seed)
    q0 = torch.randn(S, generator=g, dtype=torch.float64).to(dev)
    q0 = q0 - u1 * (u1 @ q0)
    q0 = q0 - u1 * (u1 @ q0)
    Q[1] = q0 / q0.norm()
    alpha = torch.zeros(m, dtype=torch.float64, device=dev)
    beta = torch.zeros(m, dtype=torch.float64, device=dev)
    steps, window = 0, 8
    for j in range(m):
        w = matvec(Q[j + 1])
        alpha[j] = w @ Q[j + 1]
        basis = Q[: j + 2]
        w = w - basis.T @ (basis @ w)
        w = w - basis.T @ (basis @ w)
        beta[j] = w.norm()
        steps = j + 1
        if steps % window == 0 or steps == m:  # one host sync per window: has the Krylov space been exhausted?
            lo = max(steps - window, 0)
            small = torch.nonzero(beta[lo:steps] < breakdown)
            if small.numel():
                steps = lo + int(small[0]) + 1
                break
        if steps >= min_steps and steps % check_every == 0 and steps < m:
            a, b = alpha[:steps].cpu(), beta[:steps].cpu()
            T = torch.diag(a) + torch.diag(b[: steps - 1], 1) + torch.diag(b[: steps - 1], -1)
            evecs = torch.linalg.eigh(T)[1]
            if float(b[steps - 1]) * float(evecs[-1, -3:].abs().max()) < tol:
                break
        if j + 1 < m:
            Q[j + 2] = w / beta[j].clamp_min(1e-300)
    a, b = alpha[:steps].cpu(), beta[:steps].cpu()
    T = torch.diag(a) + torch.diag(b[: steps - 1], 1) + torch.diag(b[: steps - 1], -1)
    evals, evecs = torch.linalg.eigh(T)
    ritz = evecs[:, pick]
    if info is not None:
        info.update(steps=steps, beta=b.tolist(), ritz_values=evals[-4:].tolist())
    u = Q[1: steps + 1].T @ ritz.to(dev)
    return u / u.norm()


def _lanczos_device(graph: NCutGraph, dinv: torch.Tensor, u1: torch.Tensor, max_steps: int, tol: float, seed: int,
                    min_steps: int = 512, segment: int = 32, breakdown: float = 1e-10, info: Optional[dict] = None, pick: int = -1) -> torch.Tensor:
    """The recurrence of `_lanczos_top_deflated` with every step on the device (us3d_ncut_lanczos, one cooperative launch per
    segment): the first launch runs `min_steps` steps (or to breakdown), later ones `segment` steps each; the host is touched
    once per launch — the completed-step count — plus the small tridiagonal eigen-solves of the convergence test."""
    S, dev = graph.n, graph.bits.device
    st = _stream()
    m = min(max_steps, S - 1)
    Q = torch.zeros((m + 2, S), dtype=torch.float64, device=dev)
    Q[0] = u1
    g = torch.Generator(device="cpu").manual_seed(seed)
    q0 = torch.randn(S, generator=g, dtype=torch.float64).to(dev)
    q0 = q0 - u1 * (u1 @ q0)
    q0 = q0 - u1 * (u1 @ q0)
    Q[1] = q0 / q0.norm()
    alpha = torch.zeros(m, dtype=torch.float64, device=dev)
    beta = torch.zeros(m, dtype=torch.float64, device=dev)
    ws_bytes = int(lib.us3d_ncut_lanczos_workspace_bytes(m))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    done = torch.zeros(1, dtype=torch.int32, device=dev)
    dinv = dinv.contiguous()
    steps, launches = 0, 0
    while steps < m:
        j1 = min(m, min_steps if steps == 0 else steps + segment)
        check(lib.us3d_ncut_lanczos(graph.bits.data_ptr(), S, float(graph.eps), dinv.data_ptr(), Q.data_ptr(), alpha.data_ptr(),
                                    beta.data_ptr(), steps, j1, m, float(breakdown), ws.data_ptr(), ws_bytes, done.data_ptr(), st))
        launches += 1
        steps = int(done.item())
        if steps < j1:  # the Krylov space is exhausted
            break
        if steps >= min_steps and steps < m:
            a, b = alpha[:steps].cpu(), beta[:steps].cpu()
            T = torch.diag(a) + torch.diag(b[: steps - 1], 1) + torch.diag(b[: steps - 1], -1)
            evecs = torch.linalg.eigh(T)[1]
            if float(b[steps - 1]) * float(evecs[-1, -3:].abs().max()) < tol:
                break
    a, b = alpha[:steps].cpu(), beta[:steps].cpu()
    T = torch.diag(a) + torch.diag(b[: steps - 1], 1) + torch.diag(b[: steps - 1], -1)
    evals, evecs = torch.linalg.eigh(T)
    ritz = evecs[:, pick]
    if info is not None:
        info.update(steps=steps, beta=b.tolist(), ritz_values=evals[-4:].tolist(), launches=launches)
    u = Q[1: steps + 1].T @ ritz.to(dev)
    return u / u.norm()


_fused_lanczos = {"on": True}


def set_fused_lanczos(on: bool):
    """Debugging switch: False drives the same recurrence from the host, one matvec launch per step."""
    _fused_lanczos["on"] = bool(on)


def second_smallest_eigenvector(graph: NCutGraph, max_steps: int = 4096, tol: float = 1e-10, seed: int = 0, info: Optional[dict] = None) -> torch.Tensor:
    """Eigenvector of the second smallest eigenvalue of (D - W) v = lambda D v, normalised v^T D v = 1 like
    scipy.linalg.eigh(D - A, D): Lanczos for the largest eigenpair of M = D^-1/2 W D^-1/2 (W x from the bit matrix on
    the device) in the complement of its known leading eigenvector D^1/2 1, then v = D^-1/2 u."""
    S, dev = graph.n, graph.bits.device
    st = _stream()
    dinv = graph.degree.rsqrt()
    # with the leading eigenvector known it is deflated and the largest remaining Ritz pair is the answer; otherwise nothing is
    # deflated (a zero row takes its place in the basis) and the second largest pair is taken
    pick = -1 if graph.trivial_known else -2
    if _fused_lanczos["on"]:
        u1 = graph.degree.sqrt()
        u1 = u1 / u1.norm() if graph.trivial_known else torch.zeros_like(u1)
        return dinv * _lanczos_device(graph, dinv, u1, max_steps, tol, seed, info=info, pick=pick)
    y = torch.empty(S, dtype=torch.float64, device=dev)

    def matvec(u):
        x = (dinv * u).contiguous()
        xs = x.sum().reshape(1)
        check(lib.us3d_ncut_matvec(graph.bits.data_ptr(), S, float(graph.eps), x.data_ptr(), xs.data_ptr(), y.data_ptr(), st))
        return dinv * y

    u1 = graph.degree.sqrt()
    u1 = u1 / u1.norm() if graph.trivial_known else torch.zeros_like(u1)
    return dinv * _lanczos_top_deflated(matvec, u1, max_steps, tol, seed, info=info, pick=pick)


class SegmentGraph:
    """Directed segment adjacency as CSR over POSITIONS in `unique_segments` (rows [a, b] of seg_connectivity: b is listed for a;
    ids that are not segments of the scene are dropped — they can never be part of a blob).  Built once per scene."""

    def __init__(self, unique_segments: torch.Tensor, seg_connectivity: torch.Tensor):
        ids = unique_segments.cpu().numpy()
        conn = seg_connectivity.cpu().numpy().reshape(-1, 2)
        order = np.argsort(ids, kind="stable")
        sorted_ids = ids[order]
        pa = np.searchsorted(sorted_ids, conn[:, 0])
        pb = np.searchsorted(sorted_ids, conn[:, 1])
        ok = (pa < len(ids)) & (pb < len(ids))
        ok[ok] &= (sorted_ids[pa[ok]] == conn[ok, 0]) & (sorted_ids[pb[ok]] == conn[ok, 1])
        a, b = order[pa[ok]], order[pb[ok]]
        by_a = np.argsort(a, kind="stable")
        self.ids = ids
        self.adj = np.ascontiguousarray(b[by_a], dtype=np.int32)
        self.adj_ptr = np.zeros(len(ids) + 1, dtype=np.int32)
        np.cumsum(np.bincount(a, minlength=len(ids)), out=self.adj_ptr[1:])

    def blobs(self, bipartition: np.ndarray):
        """The reference's incremental blob growing over the foreground segments (:190-226) — libus3d host function, literal
        list semantics (visit in position order, join every blob that holds a listed neighbour, merge bridged blobs, the scan
        index still advances after a merge).  Returns the blobs as arrays of positions, in the reference's list order."""
        s = len(self.ids)
        mask = np.ascontiguousarray(bipartition, dtype=np.uint8).reshape(1, s)
        k = int(mask.sum())
        if k == 0:
            return []
        query = np.empty(k, dtype=np.int32)
        ptr = np.empty(k + 1, dtype=np.int32)
        members = np.empty(k, dtype=np.int32)
        nb = lib.us3d_freemask_separate_h(mask.ctypes.data, 1, s, self.adj_ptr.ctypes.data, self.adj.ctypes.data, query.ctypes.data,
                                          ptr.ctypes.data, members.ctypes.data, k, k)
        if nb < 0:
            check(nb)
        return [members[ptr[i]:ptr[i + 1]] for i in range(nb)]


def separate_segments(bipartition: np.ndarray, vec: np.ndarray, unique_segments: torch.Tensor, seg_connectivity: torch.Tensor, mode: str = "max",
                      graph: Optional[SegmentGraph] = None):
    """separate_segments (:181-250): blobs of connected foreground segments, then 'max' (configured,
    pseudo_masks/config/default.yaml:64-74): the blob containing argmax(vec); 'avg': the blob with the highest mean of vec;
    'largest': the blob with most segments; 'all': every foreground segment.  Returns a set of segment ids."""
    if mode not in ("max", "avg", "largest", "all"):
        raise NotImplementedError(mode)
    graph = graph or SegmentGraph(unique_segments, seg_connectivity)
    ids = graph.ids
    if mode == "all":
        return set(int(c) for c in ids[bipartition])
    blobs = graph.blobs(bipartition)
    if mode == "avg":
        pick = blobs[int(np.argmax([np.mean(vec[b]) for b in blobs]))]
    elif mode == "largest":
        pick = blobs[int(np.argmax(np.array([len(b) for b in blobs])))]
    else:
        seed = int(np.argmax(vec))
        pick = next(b for b in blobs if seed in b)
    return set(ids[pick].tolist())


def unscene3d(aggregated_features, unique_segments, seg_connectivity, affinity_tau=0.65, max_number_of_instances=20,
              max_extent_ratio=0.8, eps=1e-5, min_segment_size=4, separation_mode="max",
              sign_rule: Union[str, Callable[[np.ndarray], float]] = "minority", trace=None) -> np.ndarray:
    """Greedy NCut extraction; aggregated_features = (feats_a, feats_b) per segment (CUDA), or one tensor for the
    single-modality affinity.  Returns bool [M, S]."""
    single = not isinstance(aggregated_features, tuple)
    fa, fb = (aggregated_features, aggregated_features) if single else aggregated_features
    S = len(unique_segments)
    if S < 3:
        return np.ones((1, S), dtype=bool)
    dev = fa.device
    ids = unique_segments.cpu().numpy()
    masks, foreground = [], set()
    seg_graph = SegmentGraph(unique_segments, seg_connectivity)
    painting = torch.zeros(S, dtype=torch.bool, device=dev)
    current = None
    fa, fb = fa.clone(), fb.clone()
    for it in range(max_number_of_instances):
        if it > 0:
            painting = painting | current
            keep = (~painting).float()[:, None]
            fa, fb = keep * fa, keep * fb
        graph = get_affinity_matrix(fa if single else (fa, fb), tau=affinity_tau, eps=eps, painted=painting)
        vec = second_smallest_eigenvector(graph).cpu().numpy()
        if callable(sign_rule):
            vec = vec * sign_rule(vec)
        else:
            fg = vec > vec.sum() / len(vec)
            if fg.sum() * 2 > len(vec):
                vec = -vec
        if trace is not None:
            trace.append(vec.copy())
        bip = vec > vec.sum() / len(vec)
        if bip.sum() / len(bip) > max_extent_ratio:
            bip, vec = np.logical_not(bip), -vec
        part = separate_segments(bip, vec, unique_segments, seg_connectivity, mode=separation_mode, graph=seg_graph)
        current = torch.from_numpy(np.isin(ids, list(part))).to(dev)
        iou = len(part & foreground) / len(part)
        if iou > 0.5 or len(part) < min_segment_size:
            continue
        masks.append(np.isin(ids, list(part - foreground)))
        foreground |= part
    return np.stack(masks) if masks else np.zeros((0, S), dtype=bool)


# ------------------------------------------------------------------------------------------- FreeMask-style variant (A22)
def _row_stats(soft, thr, weights=None, seg_min=None, seg_max=None):
    m, s = soft.shape
    dev = soft.device
    count = torch.empty(m, dtype=torch.int32, device=dev)
    soft_sum = torch.empty(m, dtype=torch.float32, device=dev)
    points = torch.empty(m, dtype=torch.int64, device=dev) if weights is not None else None
    bbox = torch.empty((m, 6), dtype=torch.float64, device=dev) if seg_min is not None else None
    check(lib.us3d_freemask_row_stats(soft.data_ptr(), soft.stride(0), m, s, float(thr), Fn._ptr(weights), Fn._ptr(seg_min), Fn._ptr(seg_max),
                                      count.data_ptr(), soft_sum.data_ptr(), Fn._ptr(points), Fn._ptr(bbox), _stream()))
    return count, soft_sum, points, bbox


def freemask(keys_F: torch.Tensor, matching_segment_ids: torch.Tensor, seg_connectivity: torch.Tensor, lr_coords, coords: torch.Tensor,
             hard_mask_threshold: float = 0.35, nms_maskness_threshold: float = 0.6, instance_to_scene_max_ratio: float = 0.8,
             max_instance_num: int = 50, nms_thr: float = 0.5):
    """Segment branch of the scene loop of pseudo_masks/freemask_main.py:203-417 (defaults: pseudo_masks/config/default.yaml:57-63).

    keys_F [N, C] low-resolution point features (CUDA), matching_segment_ids [N] int64, seg_connectivity [E, 2] int64 (directed),
    lr_coords [N, 3] low-resolution point coordinates, coords [N0, 4] (batch, xyz) of the full-resolution cloud.
    Returns (soft_masks [M, N] float32 on the low-resolution points, maskness [M]) on the host like the reference (:405-406), or
    None where the reference skips the scene.

    Device: segment means, soft masks, every per-candidate statistic and the pairwise intersections; host: the blob separation
    (libus3d host function) and the greedy suppression decisions over the intersection matrix — as in the reference, which runs
    both as Python loops.  No [M, N] tensor is formed before the final <= max_instance_num masks."""
    if not keys_F.is_cuda:
        raise RuntimeError("unscene3d_b200 operators run on CUDA tensors only (no CPU fallback)")
    dev, st, thr = keys_F.device, _stream(), float(hard_mask_threshold)
    ids = matching_segment_ids.to(dev).long()
    unique_all, index = torch.unique(ids, return_inverse=True)
    S_all = unique_all.shape[0]
    # ---- :203-222 per-segment mean over the valid rows; segments whose mean is all-zero are dropped
    valid = torch.any(keys_F != 0, dim=-1)
    src, idx = keys_F[valid].float().contiguous(), index[valid].contiguous()
    feats = torch.empty((S_all, keys_F.shape[1]), dtype=torch.float32, device=dev)
    acc = torch.zeros((S_all, keys_F.shape[1]), dtype=torch.float64, device=dev)
    cnt = torch.zeros(S_all, dtype=torch.float32, device=dev)
    check(lib.us3d_segment_mean_f64(src.data_ptr(), idx.data_ptr(), src.shape[0], src.shape[1], S_all, acc.data_ptr(), feats.data_ptr(),
                                    cnt.data_ptr(), st))
    valid_seg = torch.any(feats != 0, dim=-1)
    feats = feats[valid_seg].contiguous()
    unique_segments = unique_all[valid_seg]
    S = feats.shape[0]
    if S == 0:
        return None
    pos_of_all = torch.full((S_all,), -1, dtype=torch.int64, device=dev)
    pos_of_all[valid_seg] = torch.arange(S, device=dev)
    point_pos = pos_of_all[index]                                     # position of every point's segment, -1 = dropped segment
    # ---- :242-279 soft masks, hard threshold, candidates with more than 2 segments
    soft = torch.empty((S, S), dtype=torch.float32, device=dev)
    norm = torch.empty(S, dtype=torch.float32, device=dev)
    check(lib.us3d_freemask_soft_masks(feats.data_ptr(), S, feats.shape[1], norm.data_ptr(), soft.data_ptr(), st))
    count, _, _, _ = _row_stats(soft, thr)
    keep = count > 2
    if not bool(keep.any()):
        return None
    soft = soft[keep].contiguous()
    # ---- :282-349 separation of non-connected blobs (host), candidates with more than 3 segments
    conn = seg_connectivity.to(dev).long()
    a = torch.searchsorted(unique_segments, conn[:, 0].contiguous())
    b = torch.searchsorted(unique_segments, conn[:, 1].contiguous())
    a_c, b_c = a.clamp(max=S - 1), b.clamp(max=S - 1)
    ok = (unique_segments[a_c] == conn[:, 0]) & (unique_segments[b_c] == conn[:, 1])   # edges between kept segments
    a, b = a_c[ok], b_c[ok]
    order = torch.argsort(a, stable=True)
    adj = b[order].to(torch.int32).cpu().numpy()
    adj_ptr = np.zeros(S + 1, dtype=np.int32)
    np.cumsum(torch.bincount(a, minlength=S).cpu().numpy(), out=adj_ptr[1:])
    masks_h = (soft >= thr).to(torch.uint8).cpu().numpy()
    m = masks_h.shape[0]
    members_cap = int(masks_h.sum())
    blobs_cap = members_cap
    blob_query = np.empty(max(blobs_cap, 1), dtype=np.int32)
    blob_ptr = np.empty(max(blobs_cap, 1) + 1, dtype=np.int32)
    blob_members = np.empty(max(members_cap, 1), dtype=np.int32)
    nb = lib.us3d_freemask_separate_h(masks_h.ctypes.data, m, S, adj_ptr.ctypes.data, adj.ctypes.data, blob_query.ctypes.data,
                                      blob_ptr.ctypes.data, blob_members.ctypes.data, blobs_cap, members_cap)
    if nb < 0:
        check(nb)
    if nb == 0:
        return None
    sizes = np.diff(blob_ptr[: nb + 1])
    rows = torch.from_numpy(np.repeat(np.arange(nb), sizes)).to(dev)
    qrows = torch.from_numpy(np.repeat(blob_query[:nb], sizes).astype(np.int64)).to(dev)
    cols = torch.from_numpy(blob_members[: int(blob_ptr[nb])].astype(np.int64)).to(dev)
    sep = torch.zeros((nb, S), dtype=torch.float32, device=dev)
    sep[rows, cols] = soft[qrows, cols]
    count, soft_sum, _, _ = _row_stats(sep, thr)
    keep = count > 3
    sep, count, soft_sum = sep[keep].contiguous(), count[keep], soft_sum[keep]
    if sep.shape[0] == 0:
        return None
    # ---- :353-356 maskness, descending
    maskness = soft_sum / count.to(torch.int64)
    sort_inds = torch.argsort(maskness, descending=True, stable=True)
    maskness, sep = maskness[sort_inds], sep[sort_inds].contiguous()
    segment_level_counts = count.to(torch.int64)                       # the reference's stale `sum_masks` (pre-sort order), see below
    # ---- :359-396 mapped onto the points only in effect: per-segment point counts and boxes stand in for the [M, N] masks
    lr = torch.as_tensor(lr_coords, device=dev).double()
    member = point_pos >= 0
    weights = torch.bincount(point_pos[member], minlength=S).to(torch.int32)
    pp = point_pos[member][:, None].expand(-1, 3)
    seg_min = torch.full((S, 3), float("inf"), dtype=torch.float64, device=dev).scatter_reduce(0, pp, lr[member], "amin")
    seg_max = torch.full((S, 3), float("-inf"), dtype=torch.float64, device=dev).scatter_reduce(0, pp, lr[member], "amax")
    _, _, points, bbox = _row_stats(sep, thr, weights, seg_min.contiguous(), seg_max.contiguous())
    scene_extents = (coords[:, 1:].max(0)[0] - coords[:, 1:].min(0)[0]).cpu().numpy()
    points_h, bbox_h = points.cpu().numpy(), bbox.cpu().numpy()
    lr_is_int = not torch.as_tensor(lr_coords).is_floating_point()
    filtered = []
    for mask_id in range(sep.shape[0]):
        if points_h[mask_id] == 0:
            continue
        inst_extent = bbox_h[mask_id, 3:] - bbox_h[mask_id, :3]
        if lr_is_int:
            inst_extent = inst_extent.astype(np.int64)
        if np.any((inst_extent / scene_extents)[:2] > instance_to_scene_max_ratio):
            continue
        filtered.append(mask_id)
    if filtered:
        sel = torch.tensor(filtered, device=dev)
        sep, maskness = sep[sel].contiguous(), maskness[sel]
        sum_masks = points_h[filtered]
    else:  # the reference keeps every mask and its previous, segment-level `sum_masks` (:391-396)
        sum_masks = segment_level_counts.cpu().numpy()
    # ---- :398 matrix_nms(kernel='mask'): intersections on the device, the greedy decisions on the host
    M = sep.shape[0]
    inter = torch.empty((M, M), dtype=torch.int32, device=dev)
    check(lib.us3d_freemask_weighted_inter(sep.data_ptr(), sep.stride(0), M, S, thr, weights.data_ptr(), inter.data_ptr(), st))
    inter_h = inter.cpu().numpy().astype(np.float32)
    keep_h = np.ones(M, dtype=bool)
    sums_f = sum_masks.astype(np.int64)
    for i in range(M - 1):
        if not keep_h[i]:
            continue
        union = (sums_f[i] + sums_f[i + 1:]).astype(np.float32) - inter_h[i, i + 1:]
        with np.errstate(divide="ignore", invalid="ignore"):
            drop = np.where(union > 0, inter_h[i, i + 1:] / union > np.float32(nms_thr), True)
        keep_h[i + 1:] &= ~drop
    maskness = maskness.clone()
    maskness[torch.from_numpy(~keep_h).to(dev)] = 0.0
    # ---- :400-417 ranking, top max_instance_num, maskness threshold
    sort_inds = torch.argsort(maskness, descending=True, stable=True)[:max_instance_num]
    maskness, sep = maskness[sort_inds], sep[sort_inds]
    keep = maskness > nms_maskness_threshold
    if not bool(keep.any()):
        return None
    sep, maskness = sep[keep], maskness[keep]
    out = torch.zeros((sep.shape[0], ids.shape[0]), dtype=torch.float32, device=dev)
    out[:, member] = sep[:, point_pos[member]]
    return out.cpu(), maskness.cpu()


# ------------------------------------------------------------------------------------------- scene features, 3D branch (A17)
def encode_scene_feats_3d(model, sinput, resolution_scale: int = 2) -> torch.Tensor:
    """3D branch of encode_scene_feats (pseudo_masks/unscene3d_pseudo_main.py:332-348): forward through the pretrained
    multi-resolution backbone (models/res16unet.py:428-505), take `res_<resolution_scale>` and hand every full-resolution voxel
    the features of its nearest low-resolution voxel.

    The reference finds that voxel with a scipy KDTree over the low-resolution coordinates on the host.  The low-resolution
    map is floor(c / s) * s of the full-resolution one, and the parent p of a voxel c = p + d, d in [0, s)^3, is always a
    nearest low-resolution voxel (any other one, p + s e, differs from c by |d_i - s e_i| >= |d_i| on every axis), so the
    lookup is the coordinate manager's parent map composed over the strides — a gather, no tree, no host round trip.  Where
    several low-resolution voxels are equally near (d_i = s / 2 on an axis with a populated neighbour) the KDTree's choice is
    implementation-defined; this function returns the parent.  (`whiten=True` is not on the configured path:
    pseudo_masks/config/default.yaml:76.)"""
    scale = int(resolution_scale)
    if scale < 1 or scale & (scale - 1):
        raise ValueError("resolution_scale must be a power of two")
    _, feature_maps = model(sinput)
    enc = feature_maps[f"res_{scale}"]
    cm = sinput.coordinate_manager
    key = sinput.coordinate_map_key
    parent = None
    while key.tensor_stride[0] < enc.coordinate_map_key.tensor_stride[0]:
        nxt = cm.stride(key, (2, 2, 2))
        step = cm._parents[(key, nxt)].long()
        parent = step if parent is None else step[parent]
        key = nxt
    assert key == enc.coordinate_map_key, "the backbone's feature map does not lie on the input's coordinate pyramid"
    feats = enc.F.detach()
    return feats if parent is None else feats[parent]
