"""`project_features_cuda` (utils/cuda_utils/project_image_cuda.cpp:30-33) on libus3d: the 2D -> 3D feature lifting the pseudo-mask
path runs per image (utils/cuda_utils/raycast_image.py:18-77).  Same calling convention as the reference extension: every tensor
is allocated by the caller and filled in place, CUDA + contiguous inputs are required, nothing is returned."""
import torch

from unscene3d_b200._lib import check, lib
from unscene3d_b200.engine.coords import _stream


def project_features_cuda(encoded_2d_features, occupancy_3D, viewMatrixInv, intrinsicParams, opts, mapping2dto3d_num, projected_features, pred_mode_t):
    for name, t in (("encoded_2d_features", encoded_2d_features), ("occupancy_3D", occupancy_3D), ("viewMatrixInv", viewMatrixInv),
                    ("intrinsicParams", intrinsicParams), ("mapping2dto3d_num", mapping2dto3d_num), ("projected_features", projected_features)):
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous")
    pred_mode = bool(pred_mode_t[0])
    B, V, H, W, C = encoded_2d_features.shape
    _, Z, Y, X = occupancy_3D.shape
    o = [float(v) for v in opts[:5]]
    if int(o[0] + 0.5) != W or int(o[1] + 0.5) != H:
        raise RuntimeError("project_features_cuda: opts width / height differ from the feature tensor")
    want = torch.int32 if pred_mode else torch.float32
    if encoded_2d_features.dtype != want or projected_features.dtype != want or occupancy_3D.dtype != torch.int64:
        raise RuntimeError(f"project_features_cuda: expected {want} features / output and an int64 occupancy grid")
    hit = torch.empty(B * V * H * W, dtype=torch.int32, device=encoded_2d_features.device)
    check(lib.us3d_project_features_2d3d(encoded_2d_features.data_ptr(), occupancy_3D.data_ptr(), viewMatrixInv.float().contiguous().data_ptr(),
                                         intrinsicParams.float().contiguous().data_ptr(), B, V, H, W, C, Z, Y, X, o[2], o[3], o[4], int(pred_mode),
                                         hit.data_ptr(), mapping2dto3d_num.data_ptr(), projected_features.data_ptr(), _stream()))


def unproject_depth_images(*args, **kwargs):
    raise NotImplementedError("project_features_cuda.unproject_depth_images is not on the pseudo-mask or training path and is not built")
