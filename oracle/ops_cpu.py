"""CPU restatements of the non-MinkowskiEngine native operators on the hot path.
TEST INFRASTRUCTURE (see oracle/__init__.py).

  furthest_point_sampling   third_party/pointnet2/_ext_src/src/sampling_gpu.cu:72-176 (+ cuda_utils.h:15-21)
  scatter_mean              torch_scatter.scatter_mean as used at models/mask3d.py:223 (un-vendored dep)
  matcher_cost              models/matcher.py:12-59, 97-160 (the reference functions themselves run on CPU;
                            this restatement is validated against them in tests/test_oracle_reference.py)
"""
import numpy as np
import torch
import torch.nn.functional as F


def opt_n_threads(work_size: int) -> int:
    """cuda_utils.h:15-21: min(2^floor(log2 n), 512), at least 1."""
    pow_2 = int(np.log(float(work_size)) / np.log(2.0))
    return max(min(1 << pow_2, 512), 1)


def furthest_point_sampling(xyz: np.ndarray, m: int) -> np.ndarray:
    """xyz float32 [n, 3] -> int32 [m].  Simulates the kernel's thread slots and its shared-memory
    reduction tree literally (slot s is merged with s+h for h = T/2..1, lower slot kept on ties)."""
    xyz = np.asarray(xyz, dtype=np.float32)
    n = xyz.shape[0]
    T = opt_n_threads(n)
    temp = np.full(n, 1e10, dtype=np.float32)
    idxs = np.zeros(m, dtype=np.int32)
    mag = (xyz[:, 0] * xyz[:, 0] + xyz[:, 1] * xyz[:, 1] + xyz[:, 2] * xyz[:, 2]).astype(np.float32)
    valid = ~(mag.astype(np.float64) <= 1e-3)
    slots = np.arange(n) % T
    old = 0
    for j in range(1, m):
        diff = xyz - xyz[old]
        d = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1] + diff[:, 2] * diff[:, 2]).astype(np.float32)
        d2 = np.minimum(d, temp)
        temp = np.where(valid, d2, temp)
        # per-slot best: first strict maximum in row order among the slot's valid rows
        dists = np.full(T, -1.0, dtype=np.float32)
        dists_i = np.zeros(T, dtype=np.int64)
        cand = np.where(valid, d2, -np.inf)
        order = np.lexsort((np.arange(n), -cand, slots))  # by slot, then larger d2, then lower row
        first_of_slot = np.concatenate([[True], slots[order][1:] != slots[order][:-1]])
        best_rows = order[first_of_slot]
        ok = cand[best_rows] > -1.0
        dists[slots[best_rows][ok]] = cand[best_rows][ok]
        dists_i[slots[best_rows][ok]] = best_rows[ok]
        h = T // 2
        while h >= 1:
            v1, v2 = dists[:h].copy(), dists[h:2 * h].copy()
            i1, i2 = dists_i[:h].copy(), dists_i[h:2 * h].copy()
            dists[:h] = np.maximum(v1, v2)
            dists_i[:h] = np.where(v2 > v1, i2, i1)
            h //= 2
        old = int(dists_i[0])
        idxs[j] = old
    return idxs


def scatter_mean(src: torch.Tensor, index: torch.Tensor, dim: int = 0) -> torch.Tensor:
    assert dim == 0
    s = int(index.max()) + 1 if index.numel() else 0
    out = torch.zeros((s, src.shape[1]), dtype=src.dtype).index_add_(0, index, src)
    cnt = torch.zeros(s, dtype=src.dtype).index_add_(0, index, torch.ones(index.shape[0], dtype=src.dtype))
    return out / cnt.clamp(min=1)[:, None]


def matcher_cost(pred_logits_q: torch.Tensor, pred_mask_sq: torch.Tensor, tgt_mask_ts: torch.Tensor, tgt_labels: torch.Tensor,
                 cost_class: float, cost_mask: float, cost_dice: float) -> torch.Tensor:
    """Cost matrix [Q, T] of one scene (models/matcher.py:107-160, num_points == -1)."""
    out_prob = pred_logits_q.softmax(-1)
    ids = tgt_labels.clone()
    ignore = ids == 253
    ids[ignore] = 0
    c_class = -out_prob[:, ids]
    c_class[:, ignore] = -1.0
    x = pred_mask_sq.T.float()
    t = tgt_mask_ts.float()
    hw = x.shape[1]
    pos = F.binary_cross_entropy_with_logits(x, torch.ones_like(x), reduction="none")
    neg = F.binary_cross_entropy_with_logits(x, torch.zeros_like(x), reduction="none")
    c_mask = (torch.einsum("nc,mc->nm", pos, t) + torch.einsum("nc,mc->nm", neg, 1 - t)) / hw
    s = x.sigmoid()
    c_dice = 1 - (2 * torch.einsum("nc,mc->nm", s, t) + 1) / (s.sum(-1)[:, None] + t.sum(-1)[None, :] + 1)
    return cost_mask * c_mask + cost_class * c_class + cost_dice * c_dice


# --------------------------------------------------------------------------------------------
# registration under the reference's import names (tests / fixture scripts only)
# --------------------------------------------------------------------------------------------
def _fps_tensor(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """pointnet2._ext.furthest_point_sampling(points[B,N,3], nsamples) -> int32 [B, nsamples]."""
    pts = xyz.detach().cpu().float().numpy()
    return torch.from_numpy(np.stack([furthest_point_sampling(pts[b], npoint) for b in range(pts.shape[0])])).to(torch.int32)


def _scatter_minmax(src, index, dim, reduce):
    assert dim == 0
    s = int(index.max()) + 1 if index.numel() else 0
    idx = index.long()[:, None].expand_as(src)
    out = torch.zeros((s, src.shape[1]), dtype=src.dtype).scatter_reduce(0, idx, src, reduce=reduce, include_self=False)
    return out, None


def multihead_cross_attention(mha, query, key, value, attn_mask=None):
    """What the reference's CrossAttentionLayer executes (models/mask3d.py:561-651): nn.MultiheadAttention itself, with the
    boolean memory_mask (True = hidden) in torch's [B*h, Q, K] layout."""
    if hasattr(attn_mask, "torch_layout"):  # the decoder's own [B,K,Q] tensor (unscene3d_b200 models): expand as mask3d.py:358 does
        attn_mask = attn_mask.torch_layout(mha.num_heads)
    return mha(query=query, key=key, value=value, attn_mask=attn_mask, key_padding_mask=None)[0]


def masked_attention_core(q, k, v, mask_bhqk, num_heads):
    """Explicit restatement of the attention core on projected tensors: q [Q,B,E], k / v [K,B,E], mask [B,h,Q,K] bool
    (True = hidden) -> [Q,B,E]; softmax(q k^T / sqrt(head_dim)) v per (scene, head) as torch's
    multi_head_attention_forward computes it.  Differentiable (float64 capable)."""
    Q, B, E = q.shape
    K = k.shape[0]
    hd = E // num_heads
    qh = q.reshape(Q, B, num_heads, hd).permute(1, 2, 0, 3)
    kh = k.reshape(K, B, num_heads, hd).permute(1, 2, 0, 3)
    vh = v.reshape(K, B, num_heads, hd).permute(1, 2, 0, 3)
    s = (qh @ kh.transpose(-1, -2)) * (float(hd) ** -0.5)
    if mask_bhqk is not None:
        s = s.masked_fill(mask_bhqk, float("-inf"))
    o = torch.softmax(s, dim=-1) @ vh  # [B,h,Q,hd]
    return o.permute(2, 0, 1, 3).reshape(Q, B, E)


def dice_loss(inputs, targets, num_masks, weights):
    """models/criterion.py:22-39."""
    inputs = inputs.sigmoid().flatten(1)
    numerator = 2 * (inputs * targets).sum(-1)
    denominator = inputs.sum(-1) + targets.sum(-1)
    loss = weights * (1 - (numerator + 1) / (denominator + 1))
    return loss.sum() / num_masks


def sigmoid_ce_loss(inputs, targets, num_masks, weights):
    """models/criterion.py:47-65."""
    loss = weights.view(-1, 1) * torch.nn.functional.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    return loss.mean(1).sum() / num_masks


def mask_losses(logits_sq, targets_ts, qidx, tidx, weights, n):
    """(loss_mask, loss_dice) of one scene as models/criterion.py:176-216 computes them: matched columns of the [S, Q] logits
    against the matched rows of the [T_all, S] targets."""
    pred = logits_sq[:, qidx].T
    tgt = targets_ts[tidx].float()
    w = torch.ones(pred.shape[0], device=pred.device) if weights is None else weights
    return sigmoid_ce_loss(pred, tgt, n, w), dice_loss(pred, tgt, n, w)


def segment_attention_masks(model, mask_features, output_segments, point2segment, num_pooling_steps):
    """models/mask3d.py:419-446 as the reference runs it: segment logits gathered to the voxels, concatenated into a
    SparseTensor on mask_features' map, average-pooled `num_pooling_steps` times, `sigmoid < 0.5`."""
    ST = type(mask_features)
    output_masks = [seg[p2s] for seg, p2s in zip(output_segments, point2segment)]
    attn_mask = ST(features=torch.cat(output_masks), coordinate_manager=mask_features.coordinate_manager,
                   coordinate_map_key=mask_features.coordinate_map_key)
    for _ in range(num_pooling_steps):
        attn_mask = model.pooling(attn_mask.float())
    return ST(features=(attn_mask.F.detach().sigmoid() < 0.5), coordinate_manager=attn_mask.coordinate_manager,
              coordinate_map_key=attn_mask.coordinate_map_key)


def fourier_posenc(xyz, gauss_b, d_out, lo=None, hi=None):
    """Rows [N, 2 d_out] of the reference's Fourier features, restated operation by operation from
    models/position_embedding.py:12-40 (shift_scale_points onto the unit cube) and :128-160 (get_fourier_embeddings):
    clone, normalise, `xyz *= 2 * np.pi`, `torch.mm(xyz, gauss_B[:, :d_out])`, cat(sin, cos).  The reference returns the
    transpose [1, 2 d_out, N]; every caller permutes it back (models/mask3d.py:195-196)."""
    import numpy as np

    t = xyz.clone().float()
    if lo is not None:
        lo, hi = lo.reshape(1, -1).float(), hi.reshape(1, -1).float()
        src_diff = hi - lo                                    # :30  src_diff = src_range[1][:, None, :] - src_range[0][:, None, :]
        dst_lo, dst_diff = torch.zeros_like(lo), torch.ones_like(lo)  # :18-22 default dst_range = unit cube
        t = ((t - lo) * dst_diff) / src_diff + dst_lo         # :32-35
    t *= 2 * np.pi                                            # :149
    proj = torch.mm(t.view(-1, gauss_b.shape[0]), gauss_b[:, :d_out].float())  # :150-152
    return torch.cat([proj.sin(), proj.cos()], dim=1)         # :153-156 (before the permute)


def as_module_tree():
    """Module objects `torch_scatter`, `pointnet2`, `pointnet2._ext` exporting the CPU restatements."""
    import types

    ts = types.ModuleType("torch_scatter")
    ts.scatter_mean = lambda src, index, dim=-1, out=None, dim_size=None: scatter_mean(src, index, dim)
    ts.scatter_max = lambda src, index, dim=-1, out=None, dim_size=None: _scatter_minmax(src, index, dim, "amax")
    ts.scatter_min = lambda src, index, dim=-1, out=None, dim_size=None: _scatter_minmax(src, index, dim, "amin")
    pn = types.ModuleType("pointnet2")
    pn.__path__ = []
    ext = types.ModuleType("pointnet2._ext")
    ext.furthest_point_sampling = _fps_tensor
    pn._ext = ext
    return {"torch_scatter": ts, "pointnet2": pn, "pointnet2._ext": ext}


def project_voxels_to_planes(coords, pred, tgt, dims):
    """utils/cuda_utils/cuda_utils_kernel.cu:371-433 restated with numpy: sums of predictions / targets and voxel counts per cell of
    the xy, xz and yz planes.  coords int [n, 4] (batch, x, y, z) centred at zero; dims = (x_dim, y_dim, z_dim) = the MAXIMUM
    coordinates (models/noise_robust_loss.py:84): voxels with a coordinate >= its dim are skipped (:392)."""
    import numpy as np

    c = np.asarray(coords)[:, 1:].astype(np.int64)
    p, t = np.asarray(pred, dtype=np.float64), np.asarray(tgt, dtype=np.float64)
    xd, yd, zd = (int(d) for d in dims)
    inst = p.shape[1]
    ok = (c >= 0).all(1) & (c[:, 0] < xd) & (c[:, 1] < yd) & (c[:, 2] < zd)
    c, p, t = c[ok], p[ok], t[ok]
    out = {}
    for name, (a, b), (da, db) in (("xy", (0, 1), (xd, yd)), ("xz", (0, 2), (xd, zd)), ("yz", (1, 2), (yd, zd))):
        cell = c[:, a] * db + c[:, b]
        num = np.bincount(cell, minlength=da * db).reshape(da, db)
        ps, ts = np.zeros((da * db, inst)), np.zeros((da * db, inst))
        np.add.at(ps, cell, p)
        np.add.at(ts, cell, t)
        out[name] = (ps.reshape(da, db, inst), ts.reshape(da, db, inst), num.astype(np.int32))
    return out


def project_voxels_to_planes_bwd(coords, grads, dims, n_inst):
    """cuda_utils_kernel.cu:496-556: per voxel and instance the mean of the non-zero plane gradients at its three cells
    (grads = dict xy / xz / yz of float [da, db, inst]); skipped voxels keep 0."""
    import numpy as np

    c = np.asarray(coords)[:, 1:].astype(np.int64)
    xd, yd, zd = (int(d) for d in dims)
    ok = (c >= 0).all(1) & (c[:, 0] < xd) & (c[:, 1] < yd) & (c[:, 2] < zd)
    out = np.zeros((c.shape[0], n_inst), dtype=np.float32)
    cc = c[ok]
    g = [np.asarray(grads["xy"], dtype=np.float32)[cc[:, 0], cc[:, 1]], np.asarray(grads["xz"], dtype=np.float32)[cc[:, 0], cc[:, 2]],
         np.asarray(grads["yz"], dtype=np.float32)[cc[:, 1], cc[:, 2]]]
    cnt = sum((x != 0).astype(np.int32) for x in g)
    s = (g[0] + g[1]) + g[2]
    out[ok] = np.where(cnt > 0, s / np.maximum(cnt, 1).astype(np.float32), 0.0).astype(np.float32)
    return out


def project_features_2d3d(feats, occ, view_inv, intr, depth_min, depth_max, ray_inc):
    """utils/cuda_utils/project_image_cuda_kernel.cu:24-64, 113-146 restated with numpy float32 (vectorised over the pixels): every
    pixel's ray is marched through occ [B, Z, Y, X] (0 = empty, else voxel index) from depth_min to depth_max in steps of ray_inc;
    returns (hit [B, V, H, W] voxel index of the first occupied cell or 0, counts [n_vox], sums [n_vox, C]).  The CUDA kernels fuse
    a * b + c into FMAs, numpy does not: rays grazing a cell boundary within float round-off may land in the neighbouring cell."""
    import numpy as np

    f32 = np.float32
    feats, occ = np.asarray(feats, dtype=f32), np.asarray(occ)
    view_inv, intr = np.asarray(view_inv, dtype=f32), np.asarray(intr, dtype=f32)
    B, V, H, W, C = feats.shape
    _, Z, Y, X = occ.shape
    n_vox = int(occ.max()) + 1
    hit = np.zeros((B, V, H, W), dtype=np.int64)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    for b in range(B):
        fx, fy, mx, my = intr[b]
        depth = f32(1.0) * (f32(depth_max) - f32(depth_min)) + f32(depth_min)
        cx = (xs.astype(f32) - mx) / fx
        cy = (ys.astype(f32) - my) / fy
        cam = np.stack([depth * cx, depth * cy, np.full_like(cx, depth)], -1).astype(f32)
        cam = cam * (f32(1.0) / np.sqrt((cam * cam).sum(-1, dtype=f32)))[..., None]
        for v in range(V):
            m = view_inv[b, v]
            origin = m[:3, 3]
            d = (cam @ m[:3, :3].T).astype(f32)
            d = d * (f32(1.0) / np.sqrt((d * d).sum(-1, dtype=f32)))[..., None]
            scale = f32(1.0) / cam[..., 2]
            ray = scale * f32(depth_min)
            end = scale * f32(depth_max)
            found = np.zeros((H, W), dtype=np.int64)
            alive = ray < end
            while alive.any():
                w = origin[None, None, :] + ray[..., None] * d
                p = (w + np.sign(w).astype(f32) * f32(0.5)).astype(np.int64)  # truncation toward zero, as int(float)
                inside = alive & (p >= 0).all(-1) & (p[..., 0] < X) & (p[..., 1] < Y) & (p[..., 2] < Z)
                val = np.zeros((H, W), dtype=np.int64)
                val[inside] = occ[b, p[..., 2][inside], p[..., 1][inside], p[..., 0][inside]]
                newly = inside & (val != 0) & (found == 0)
                found[newly] = val[newly]
                alive = alive & ~newly
                ray = (ray + f32(ray_inc)).astype(f32)
                alive = alive & (ray < end)
            hit[b, v] = found
    counts = np.bincount(hit[hit != 0].ravel(), minlength=n_vox)
    sums = np.zeros((n_vox, C), dtype=np.float64)
    np.add.at(sums, hit[hit != 0].ravel(), feats[hit != 0].astype(np.float64))
    return hit, counts, sums
