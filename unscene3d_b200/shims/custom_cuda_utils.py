"""`custom_cuda_utils` (utils/cuda_utils/cuda_utils.cpp:49-54) on libus3d.

models/noise_robust_loss.py:11 imports this module; its projection pair runs when cost_noise_robust != 0
(models/criterion.py:170).  Same calling convention as the reference extension: every tensor is allocated by the caller and
filled in place, inputs must be contiguous CUDA tensors (CHECK_INPUT, cuda_utils.cpp:4-6 -> RuntimeError), nothing is returned.
The reference launches on the legacy default stream and synchronises; here the kernels run on torch's current stream.
`trilinear_interpolate[_backward]` (utils/cuda_utils/cuda_utils.py — not imported by the training or pseudo-mask path) are not
built and say so when called.
"""
import torch

from unscene3d_b200._lib import check, lib
from unscene3d_b200.engine.coords import _stream


def _check_input(**tensors):
    for name, t in tensors.items():
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous")


def project_sparse_voxels_to_planes(s_coords, s_predictions, s_targets, xy_pred_projections, xz_pred_projections, yz_pred_projections,
                                    xy_target_projections, xz_target_projections, yz_target_projections, xy_projection_nums,
                                    xz_projection_nums, yz_projection_nums):
    _check_input(s_coords=s_coords, s_predictions=s_predictions, s_targets=s_targets, xy_pred_projections=xy_pred_projections,
                 xz_pred_projections=xz_pred_projections, yz_pred_projections=yz_pred_projections,
                 xy_target_projections=xy_target_projections, xz_target_projections=xz_target_projections,
                 yz_target_projections=yz_target_projections, xy_projection_nums=xy_projection_nums,
                 xz_projection_nums=xz_projection_nums, yz_projection_nums=yz_projection_nums)
    if s_coords.dtype != torch.int32 or s_predictions.dtype != torch.float32 or s_targets.dtype != torch.float32:
        raise RuntimeError("project_sparse_voxels_to_planes: expected int32 coordinates and float32 predictions / targets")
    n, inst = s_predictions.shape
    x_dim, y_dim = xy_pred_projections.shape[:2]
    z_dim = xz_pred_projections.shape[1]
    check(lib.us3d_project_voxels_to_planes(s_coords.data_ptr(), s_predictions.data_ptr(), s_targets.data_ptr(), n, inst, x_dim, y_dim, z_dim,
                                            xy_pred_projections.data_ptr(), xz_pred_projections.data_ptr(), yz_pred_projections.data_ptr(),
                                            xy_target_projections.data_ptr(), xz_target_projections.data_ptr(), yz_target_projections.data_ptr(),
                                            xy_projection_nums.data_ptr(), xz_projection_nums.data_ptr(), yz_projection_nums.data_ptr(), _stream()))


def project_sparse_voxels_to_planes_backward(s_coords, s_grads, xy_grads, xz_grads, yz_grads, xy_nums, xz_nums, yz_nums):
    _check_input(s_coords=s_coords, s_grads=s_grads, xy_grads=xy_grads, xz_grads=xz_grads, yz_grads=yz_grads, xy_nums=xy_nums,
                 xz_nums=xz_nums, yz_nums=yz_nums)
    n, inst = s_grads.shape
    x_dim, y_dim = xy_grads.shape[:2]
    z_dim = xz_grads.shape[1]
    check(lib.us3d_project_voxels_to_planes_bwd(s_coords.data_ptr(), n, inst, x_dim, y_dim, z_dim, xy_grads.data_ptr(), xz_grads.data_ptr(),
                                                yz_grads.data_ptr(), s_grads.data_ptr(), _stream()))


def _not_built(name):
    def fn(*a, **k):
        raise NotImplementedError(f"custom_cuda_utils.{name}: utils/cuda_utils/cuda_utils.py's trilinear interpolation is not on the "
                                  "training or pseudo-mask path and is not built")

    fn.__name__ = name
    return fn


trilinear_interpolate = _not_built("trilinear_interpolate")
trilinear_interpolate_backward = _not_built("trilinear_interpolate_backward")
