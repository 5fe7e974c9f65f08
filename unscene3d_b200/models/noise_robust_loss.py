"""Tri-plane projection loss for noise-robust training — same module surface as the reference's
models/noise_robust_loss.py (ProjectionFunction :16-71, ProjectionFunctionWrapper :74-104, ProjectionMaskLoss :107-172), over the
`custom_cuda_utils` projection pair (libus3d: csrc/projection.cu).  Evaluated by SetCriterion.loss_masks only when the weight of
`loss_noise_robust` is non-zero (models/criterion.py:170)."""
import torch
from torch import nn
from torch.autograd import Function

import custom_cuda_utils

EPS = 10e-9


class ProjectionFunction(Function):
    """Mean prediction / target per occupied cell of the xy, xz and yz planes; the gradient of a voxel is the mean of the non-zero
    gradients of its three cells (the division by the cell counts happens on the way in, as in the reference :39-46)."""

    @staticmethod
    def forward(ctx, s_coords, s_predictions, s_targets, dims):
        x_dim, y_dim, z_dim = dims
        dev, inst = s_coords.device, s_predictions.shape[1]
        shapes = ((x_dim, y_dim), (x_dim, z_dim), (y_dim, z_dim))
        preds = [torch.zeros((*s, inst), device=dev) for s in shapes]
        tgts = [torch.zeros((*s, inst), device=dev) for s in shapes]
        nums = [torch.zeros(s, device=dev, dtype=torch.int) for s in shapes]
        custom_cuda_utils.project_sparse_voxels_to_planes(s_coords, s_predictions, s_targets, *preds, *tgts, *nums)
        ctx.save_for_backward(s_coords, *nums)
        ctx.grad_shape = s_predictions.shape
        out = []
        for group in (preds, tgts):
            for plane, num in zip(group, nums):
                plane = plane / (num.unsqueeze(-1) + EPS)
                plane[num == 0] = 0.0
                out.append(plane)
        ctx.mark_non_differentiable(*nums)
        return (*out, *nums)

    @staticmethod
    def backward(ctx, g_xy, g_xz, g_yz, *unused):
        s_coords, n_xy, n_xz, n_yz = ctx.saved_tensors
        s_grads = torch.zeros(ctx.grad_shape, device=s_coords.device)
        custom_cuda_utils.project_sparse_voxels_to_planes_backward(s_coords, s_grads, g_xy.contiguous(), g_xz.contiguous(), g_yz.contiguous(),
                                                                   n_xy, n_xz, n_yz)
        return None, s_grads, None, None


class ProjectionFunctionWrapper(nn.Module):
    def forward(self, s_coords, s_predictions, s_targets):
        centered = s_coords - torch.amin(s_coords, 0)
        x_dim, y_dim, z_dim = (int(v) for v in centered[:, 1:].max(0)[0])  # the reference sizes the planes by the maximum coordinate
        out = ProjectionFunction.apply(centered.int().contiguous(), s_predictions.contiguous(), s_targets.contiguous(), (x_dim, y_dim, z_dim))
        return out, (x_dim, y_dim, z_dim)


class ProjectionMaskLoss(nn.Module):
    def __init__(self, config=None, base_loss="bce", directions="xyz"):
        super().__init__()
        if base_loss != "bce":
            raise NotImplementedError(base_loss)
        self.base_loss, self.eps, self.directions = base_loss, EPS, directions
        self.projection_module = ProjectionFunctionWrapper()
        self.criterion = nn.BCELoss(reduction="none")

    def forward(self, all_mask_preds, all_mask_targets, coords):
        """all_mask_preds [inst, N] logits, all_mask_targets [inst, N], coords [N, 4] -> (summed BCE over the occupied cells of the
        requested views, number of (instance, occupied cell) terms)."""
        inst_num = all_mask_preds.shape[0]
        outs, _ = self.projection_module(coords, torch.sigmoid(all_mask_preds.T), all_mask_targets.T)
        xy_p, xz_p, yz_p, xy_t, xz_t, yz_t, n_xy, n_xz, n_yz = outs
        all_shape = inst_num * (int((n_xy != 0).sum()) + int((n_xz != 0).sum()) + int((n_yz != 0).sum()))
        loss = 0
        for axis, pred, tgt, num in (("x", yz_p, yz_t, n_yz), ("y", xz_p, xz_t, n_xz), ("z", xy_p, xy_t, n_xy)):
            if axis in self.directions:
                term = self.criterion(pred, tgt.detach())
                loss = loss + term[num != 0].sum()
        return loss, all_shape
